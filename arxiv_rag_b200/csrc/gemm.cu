// Encoder GEMMs: C = epi(A . B^T) on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
// Replaces the nn.Linear calls of transformers' MPNet (modeling_mpnet.py:145-159 q/k/v,
// :183 o, :225-228 intermediate+GELU, :239-243 output) that sentence-transformers' encode
// reaches from generate_embeddings_parallel.py:146-153.
//
// Epilogue: 8 warps in two groups of 4 (one warp per TMEM lane quarter); group g owns columns
// [g*BN/2, (g+1)*BN/2) of the tile. Per 128-byte-wide column chunk a thread pulls its row from
// TMEM, applies bias / GELU / residual in fp32, writes the converted row into a swizzled staging
// tile in shared memory, and one thread of the group issues a TMA store (coalesced, clipped at
// the M/N edges). The residual chunk is TMA-loaded into the same staging tile beforehand, so
// neither R nor C is ever touched with row-strided global accesses.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "umma_pipe.cuh"

namespace arb {

constexpr int kGemmBN = 256;
constexpr int kGemmStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = 64 + kEpiThreads;
constexpr int kStageTileBytes = kBM * 128;  // one staging tile: 128 rows x 128 bytes
using GemmSmem = PipeSmem<kGemmBN, kGemmStages, 2 * kStageTileBytes>;

// Static persistent schedule: tile t -> (m-block t / num_n, n-block t % num_n) so the CTAs
// running concurrently share A panels (activations) through L2 while B (weights) stays hot.
struct GemmTileIter {
    int tile, step, tiles, num_n, bn;
    __device__ __forceinline__ bool next(int& row_a, int& row_b) {
        if (tile >= tiles) return false;
        row_a = (tile / num_n) * kBM;
        row_b = (tile % num_n) * bn;
        tile += step;
        return true;
    }
};

// Narrow variant for query-time batches (a handful of row blocks): 128 x 128 tiles, so twice as many
// SMs pull weights (a CTA streams at ~0.15 TB/s whatever the tile), 32 KB stages in a 6-deep ring,
// and 1 KB after the staging tiles where the two epilogue groups (64 columns each) combine their row
// statistics into the one partial per 128 columns the consumers expect.
constexpr int kGemmNarrowBN = 128;
constexpr int kGemmNarrowStages = 6;
using GemmNarrowSmem = PipeSmem<kGemmNarrowBN, kGemmNarrowStages, 2 * kStageTileBytes + kBM * 8>;
// ARB_GEMM_NARROW=0 keeps 128 x 256 tiles for every single-CTA launch (A/B)
static const bool g_gemm_narrow = []() {
    const char* e = getenv("ARB_GEMM_NARROW");
    return !(e && e[0] == '0');
}();

// 0 = auto (see launch_gemm_impl), 1 = single CTA 128x256, 2 = CTA pairs, 3 = single CTA 128x128
static int g_gemm_mode = 0;
constexpr int64_t kGemmPairMinRows = 32 * kBM;
void set_gemm_mode(int mode) { g_gemm_mode = mode; }
// ARB_GEMM_COLSMEM=0 keeps the 5-stage kernel with per-row global loads for the o-projection (A/B)
static const bool g_gemm_colsmem = []() {
    const char* e = getenv("ARB_GEMM_COLSMEM");
    return !(e && e[0] == '0');
}();

// CTA-pair variant: 256 x 256 tiles, 32 KB per stage and CTA (A 128 x 64 + half of B): 5 stages
// leave room for double-buffered epilogue staging.
constexpr int kGemm2Stages = 5;
using Gemm2Smem = PipeSmem<kGemmBN, kGemm2Stages, 4 * kStageTileBytes, 2>;  // two staging tiles per group
// Pair variant for the LayerNorm(residual) epilogue of a short-K GEMM (the o-projection, K = 768):
// its tile main loop is only ~6 k cycles, so the epilogue has no slack. One ring stage is given up
// for 8 KB that hold gamma and (bias + beta) of every output column for the whole kernel
// (N <= kColSmemMaxN), read back with broadcast shared-memory loads instead of per-row global loads.
constexpr int kColSmemMaxN = 1024;
constexpr int kColSmemBytes = 2 * kColSmemMaxN * 4;
using Gemm2ColSmem = PipeSmem<kGemmBN, kGemm2Stages - 1, 4 * kStageTileBytes + kColSmemBytes, 2>;
template <bool k2Cta>
constexpr int kGemmStageBufs = k2Cta ? 2 : 1;

// Pair schedule: tile t -> (256-row block t / num_n, n-block t % num_n); row_a is this CTA's half.
struct Gemm2TileIter {
    int tile, step, tiles, num_n, rank;
    __device__ __forceinline__ bool next(int& row_a, int& row_b) {
        if (tile >= tiles) return false;
        row_a = (tile / num_n) * (2 * kBM) + rank * kBM;
        row_b = (tile % num_n) * kGemmBN;
        tile += step;
        return true;
    }
};

// gelu(x) = x * Phi(x) with erf from Abramowitz-Stegun 7.1.28:
//   1 - erf(z) = (1 + a1 z + ... + a6 z^6)^-16, |err| <= 3e-7, z = |x|/sqrt(2) folded into b_i.
// One MUFU (rcp) and ~14 FMA-pipe ops per element; |gelu err| <= 8.2e-7 absolute in fp32.
__device__ __forceinline__ float gelu_erf(float x) {
    constexpr float b1 = 0.0705230784 / 1.4142135623730951;
    constexpr float b2 = 0.0422820123 / 2.0;
    constexpr float b3 = 0.0092705272 / 2.8284271247461903;
    constexpr float b4 = 0.0001520143 / 4.0;
    constexpr float b5 = 0.0002765672 / 5.656854249492381;
    constexpr float b6 = 0.0000430638 / 8.0;
    const float ax = fabsf(x);
    float d = fmaf(ax, b6, b5);
    d = fmaf(ax, d, b4);
    d = fmaf(ax, d, b3);
    d = fmaf(ax, d, b2);
    d = fmaf(ax, d, b1);
    d = fmaf(ax, d, 1.0f);
    d *= d;
    d *= d;
    d *= d;
    d *= d;                                  // d^16 (may overflow to +inf -> t = 0, correct limit)
    float rcp;                               // single MUFU.RCP (1 ulp); rcp(+inf) = 0
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(d));
    const float h = 0.5f * rcp;              // 0.5 * (1 - erf(|x|/sqrt2))
    const float r = x >= 0.f ? 1.0f - h : h; // Phi(x)
    return x * r;
}

// bf16 mode: gelu(x) ~= 0.5 x (1 + tanh(0.8 x + 0.03475 x^3)), the minimax 2-term fit of the
// exact erf form (max abs error 2.8e-4 over all x, i.e. < 1/50 of a bf16 ulp at the magnitudes
// where it peaks) with the hardware tanh (MUFU.TANH, rel. error 2^-11): 6 instructions per
// element instead of 15, which is what keeps the FFN-up epilogue under the tile's MMA time.
// fp16 mode keeps the erf form above (its ulp is 8x finer).
__device__ __forceinline__ float gelu_tanh_fit(float x) {
    const float x2 = x * x;
    const float u = x * fmaf(0.03475f, x2, 0.8f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}

// gelu(x) = x sigmoid(2 u(x)) with u(x) = x (a + b x^2 + c x^4) the minimax fit of atanh(erf(x/sqrt2))
// (|gelu error| <= 2.6e-5 for every x, 1/10 of the 2-term tanh form) evaluated through ex2 and rcp,
// whose approximations are good to 2^-22 — unlike MUFU.TANH (2^-11). x^2 is clamped at 50: the
// quartic term would otherwise turn u around for |x| > 7, where the sigmoid has long saturated.
// 9 instructions (2 MUFU); candidate for the fp16 mode in place of the 15-instruction erf form.
__device__ __forceinline__ float gelu_sigmoid_fit3(float x) {
    constexpr float k = -2.0f * 1.4426950408889634f;  // exp(-2u) = 2^(k u)
    constexpr float a = 0.797507868f * k, b = 0.03700566f * k, c = -0.000351518939f * k;
    const float x2 = fminf(x * x, 50.0f);
    const float w = x * fmaf(x2, fmaf(x2, c, b), a);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(w));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

// Same fit of u(x), finished with the hardware tanh: gelu = 0.5 x (1 + tanh u). 8 instructions, 1 MUFU;
// on top of the fit's 2.6e-5 comes MUFU.TANH's 2^-11 relative error, i.e. about half an fp16 ulp of
// the result.
__device__ __forceinline__ float gelu_tanh_fit3(float x) {
    const float x2 = fminf(x * x, 50.0f);
    const float u = x * fmaf(x2, fmaf(x2, -0.000351518939f, 0.03700566f), 0.797507868f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}

// GELU of the fp16 mode: 0 = erf form, 1 = sigmoid fit, 2 = the bf16 mode's tanh fit, 3 = tanh with the
// 3-term fit (build-time A/B)
#ifndef ARB_GELU_F16
#define ARB_GELU_F16 3  // measured on the 1024 x 384 step: erf 80.2 ms, sigmoid fit 82.3, tanh 2-term 78.7, tanh 3-term 77.8
#endif
__device__ __forceinline__ float gelu_f16_mode(float x) {
#if ARB_GELU_F16 == 0
    return gelu_erf(x);
#elif ARB_GELU_F16 == 1
    return gelu_sigmoid_fit3(x);
#elif ARB_GELU_F16 == 2
    return gelu_tanh_fit(x);
#else
    return gelu_tanh_fit3(x);
#endif
}

// OutT = h16 (16-bit activations in the kF16 format) or float. k2Cta: launched as clusters of two
// CTAs that share one 256 x 256 tile through tcgen05 cta_group::2 (see umma_pipe.cuh).
template <int EPI, bool kF16, typename OutT, bool k2Cta, bool kColSmem = false, int kBN = kGemmBN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r,
              const float* __restrict__ bias, int64_t M, int N, int K, const LnFoldArgs fold, const uint32_t idesc,
              const uint32_t a_box_bytes) {
    constexpr int BN = kBN;
    constexpr bool kNarrow = kBN != kGemmBN;
    static_assert(!kNarrow || (kBN == kGemmNarrowBN && !k2Cta && !kColSmem), "narrow tiles: single-CTA schedule only");
    constexpr int CW = 128 / static_cast<int>(sizeof(OutT));  // columns per staging chunk
    constexpr int CPG = (BN / 2) / CW;                        // chunks per group per tile
    constexpr bool kLnIn = EPI == EPI_LNIN_BIAS || EPI == EPI_LNIN_BIAS_GELU;
    constexpr bool kGelu = EPI == EPI_BIAS_GELU || EPI == EPI_LNIN_BIAS_GELU;
    constexpr bool kRes = EPI == EPI_BIAS_RESIDUAL || EPI == EPI_BIAS_LNRES_STATS || EPI == EPI_BIAS_RES_STATS;
    constexpr bool kLnRes = EPI == EPI_BIAS_LNRES_STATS;
    constexpr bool kStats = EPI == EPI_BIAS_LNRES_STATS || EPI == EPI_BIAS_RES_STATS;
    static_assert(!kColSmem || (k2Cta && EPI == EPI_BIAS_LNRES_STATS), "resident column vectors: pair kernel, LN(residual) epilogue");
    using SM = typename std::conditional<kNarrow, GemmNarrowSmem,
        typename std::conditional<kColSmem, Gemm2ColSmem, typename std::conditional<k2Cta, Gemm2Smem, GemmSmem>::type>::type>::type;
    using Iter = typename std::conditional<k2Cta, Gemm2TileIter, GemmTileIter>::type;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles need a 1024-byte aligned base; align by hand (the launcher adds slack).
    SM sm{smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)};
    const int warp = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
    const int lane = threadIdx.x & 31;
    const int rank = k2Cta ? static_cast<int>(cluster_ctarank()) : 0;

    const int num_n = (N + BN - 1) / BN;
    const int kblocks = (K + kBK - 1) / kBK;
    Iter it;
    if constexpr (k2Cta) {
        const int num_m2 = static_cast<int>((M + 2 * kBM - 1) / (2 * kBM));
        it = Iter{static_cast<int>(blockIdx.x) / 2, static_cast<int>(gridDim.x) / 2, num_m2 * num_n, num_n, rank};
    } else {
        const int num_m = static_cast<int>((M + kBM - 1) / kBM);
        it = Iter{static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), num_m * num_n, num_n, BN};
    }

    uint32_t tmem_base;
    if constexpr (k2Cta) tmem_base = pipe2_setup(sm, warp, &tmap_a, &tmap_b, 2 * kEpiWarps);
    else tmem_base = pipe_setup(sm, warp, &tmap_a, &tmap_b, kEpiThreads);

    // Programmatic dependent launch (small token counts): everything above ran while the previous
    // kernel of the chain was still busy; so does the L2 prefetch of this CTA's first weight tile
    // (B is never written by the chain). Activations, statistics and outputs only after pdl_wait.
    pdl_launch_dependents();
    if (warp == 2) {
        Iter peek = it;
        int row_a, row_b;
        if (peek.next(row_a, row_b)) {
            const int rb = row_b + (k2Cta ? rank * (BN / 2) : 0);
            if (rb < N)
                for (int kb = lane; kb < kblocks; kb += 32) tma_prefetch_2d(&tmap_b, kb * kBK, rb);
        }
    }
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            if constexpr (k2Cta) pipe2_produce(sm, &tmap_a, &tmap_b, it, kblocks, rank, kEvictNormal, kEvictLast);
            else pipe_produce(sm, &tmap_a, &tmap_b, it, kblocks, kEvictNormal, kEvictLast,
                              a_box_bytes ? a_box_bytes + static_cast<uint32_t>(SM::kBBytes) : 0u);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            if constexpr (k2Cta) {
                if (rank == 0) pipe2_mma<SM>(sm, tmem_base, it, kblocks, idesc);
            } else {
                pipe_mma<SM>(sm, tmem_base, it, kblocks, idesc);
            }
        }
    } else {
        const int ew = warp - 2;         // 0..7
        const int grp = ew >> 2;         // column half owned by this warp's group
        const int lane_grp = warp & 3;   // TMEM lanes [32*lane_grp, +32) are the only ones this warp may read
        const int trow = lane_grp * 32 + lane;
        const bool leader = (ew & 3) == 0 && lane == 0;  // issues the group's TMA traffic
        // kBufs staging tiles per group. With two, consecutive live chunks alternate between them: a
        // TMA store is still draining from one while the next chunk is written into the other, and
        // (residual epilogues) the residual of chunk k+1 is requested right after the store of chunk
        // k is issued, so it has a whole chunk of epilogue work to arrive (also across tiles).
        constexpr int kBufs = kGemmStageBufs<k2Cta>;
        uint8_t* stage_base = sm.pre() + grp * kBufs * kStageTileBytes;
        const int sw = trow & 7;
        uint32_t res_phase = 0;  // bit b = parity of the next completion of this group's residual barrier b
        int kc = 0;              // live chunks processed so far by this group
        if (leader) {
            tma_prefetch_desc(&tmap_c);
            if (kRes) tma_prefetch_desc(&tmap_r);
        }
        // kColSmem: gamma and (bias + beta) of all N columns live in shared memory for the whole kernel
        float* col_gamma = reinterpret_cast<float*>(sm.pre() + 4 * kStageTileBytes);
        float* col_bb = col_gamma + kColSmemMaxN;
        if constexpr (kColSmem) {
            for (int j = threadIdx.x - 64; j < N; j += kEpiThreads) {
                col_gamma[j] = __ldg(fold.gamma + j);
                col_bb[j] = __ldg(fold.beta + j) + __ldg(bias + j);
            }
            named_bar_sync(3, kEpiThreads);
        }
        // residual prefetch cursor (leader thread, kBufs == 2): walks the same (tile, chunk) sequence
        // one live chunk ahead of the epilogue
        Iter pf_it = it;
        int pf_ra = 0, pf_rb = 0, pf_c = CPG, pf_k = 0;
        bool pf_valid = true;
        auto prefetch_next_residual = [&]() {
            while (pf_valid) {
                if (++pf_c >= CPG) {
                    pf_valid = pf_it.next(pf_ra, pf_rb);
                    pf_c = 0;
                }
                if (pf_valid && pf_rb + grp * (BN / 2) + pf_c * CW < N) break;
            }
            if (!pf_valid) return;
            const int b = pf_k & 1;
            mbar_arrive_expect_tx(sm.aux(grp * 2 + b), kStageTileBytes);
            tma_load_2d(&tmap_r, sm.aux(grp * 2 + b), stage_base + b * kStageTileBytes,
                        pf_rb + grp * (BN / 2) + pf_c * CW, pf_ra, kEvictFirst);
            ++pf_k;
        };
        if (kRes && kBufs == 2 && leader) prefetch_next_residual();
        // folded LayerNorm: the producer's row partials of the NEXT tile are requested a tile ahead
        constexpr int kMaxParts = 8;
        float2 pre[kMaxParts];
        auto request_stats = [&](int64_t row) {
#pragma unroll
            for (int p = 0; p < kMaxParts; ++p)
                pre[p] = (p < fold.parts_in && row < M) ? __ldg(fold.stats_in + static_cast<int64_t>(p) * M + row)
                                                        : make_float2(0.f, 0.f);
        };
        if constexpr (kLnIn || kLnRes) {
            Iter pk = it;
            int ra2, rb2;
            if (pk.next(ra2, rb2)) request_stats(static_cast<int64_t>(ra2) + trow);
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        int row_a, row_b;
        while (it.next(row_a, row_b)) {
            const int64_t grow = static_cast<int64_t>(row_a) + trow;
            float rstd = 1.f, nmr = 0.f;  // nmr = -mean * rstd
            if constexpr (kLnIn || kLnRes) {
                float su = 0.f, sq = 0.f;
#pragma unroll
                for (int p = 0; p < kMaxParts; ++p) {
                    su += pre[p].x;
                    sq += pre[p].y;
                }
                const float mean = su * fold.inv_width_in;
                rstd = rsqrtf(fmaxf(fmaf(-mean, mean, sq * fold.inv_width_in), 0.f) + fold.eps);
                nmr = -mean * rstd;
                Iter pk = it;  // `it` already points past this tile
                int ra2, rb2;
                if (pk.next(ra2, rb2)) request_stats(static_cast<int64_t>(ra2) + trow);
            }
            float osum = 0.f, osq = 0.f;  // statistics of the rows this tile produces (kStats)
            bool any_live = false;
            mbar_wait(sm.tmem_full(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN + grp * (BN / 2));
#pragma unroll 1
            for (int c = 0; c < CPG; ++c) {
                const int col0 = row_b + grp * (BN / 2) + c * CW;
                const bool live = col0 < N;  // group-uniform
                const int buf = kc & (kBufs - 1);
                uint8_t* stage_tile = stage_base + buf * kStageTileBytes;
                uint8_t* my_row = stage_tile + trow * 128;
                uint64_t* res_bar = sm.aux(grp * 2 + buf);
                if (kRes && kBufs == 1 && live && leader) {
                    tma_store_wait_read<0>();  // the previous store has drained the staging tile
                    mbar_arrive_expect_tx(res_bar, kStageTileBytes);
                    tma_load_2d(&tmap_r, res_bar, stage_tile, col0, row_a, kEvictFirst);
                }
                uint32_t r[32 * (CW / 32)];
#pragma unroll
                for (int q = 0; q < CW / 32; ++q)
                    tmem_ld_32x32(taddr + c * CW + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[q * 32]));
                tmem_ld_wait();
                if (c == CPG - 1) {  // accumulator fully read by this thread: hand the buffer back
                    tc_fence_before();
                    if constexpr (k2Cta) {  // one arrival per warp on the leader CTA's barrier
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote_cta(map_to_cta(smem_u32(sm.tmem_empty(acc)), 0));
                    } else {
                        mbar_arrive(sm.tmem_empty(acc));
                    }
                }
                if (!live) continue;
                any_live = true;
                ++kc;
                // the tile's values as pairs: the fp32 math below is packed (fma.rn.f32x2 & co.), which
                // halves the FMA-pipe issue slots of the epilogue
                float2 v2[CW / 2];
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) v2[j] = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
                if constexpr (kLnIn) {
                    // LN(x) W^T + b = rstd (x W'^T) - mean rstd c + b'   (columns are whole 64-wide chunks here)
                    const float4* bp = reinterpret_cast<const float4*>(bias + col0);
                    const float4* cp = reinterpret_cast<const float4*>(fold.colsum + col0);
#pragma unroll
                    for (int j = 0; j < CW / 4; ++j) {
                        const float4 b4 = __ldg(bp + j), c4 = __ldg(cp + j);
                        v2[2 * j] = __ffma2_rn(rstd2, v2[2 * j], __ffma2_rn(nmr2, make_float2(c4.x, c4.y), make_float2(b4.x, b4.y)));
                        v2[2 * j + 1] = __ffma2_rn(rstd2, v2[2 * j + 1], __ffma2_rn(nmr2, make_float2(c4.z, c4.w), make_float2(b4.z, b4.w)));
                    }
                } else if (!kLnRes && bias != nullptr) {
                    // columns beyond N read zero bias via the clamp; they are clipped by the store
                    const float4* bp = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
                    for (int j = 0; j < CW / 4; ++j) {
                        if (col0 + 4 * j < N) {
                            const float4 b4 = __ldg(bp + j);
                            v2[2 * j] = __fadd2_rn(v2[2 * j], make_float2(b4.x, b4.y));
                            v2[2 * j + 1] = __fadd2_rn(v2[2 * j + 1], make_float2(b4.z, b4.w));
                        }
                    }
                }
                if (kGelu) {
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) {
                        v2[j].x = kF16 ? gelu_f16_mode(v2[j].x) : gelu_tanh_fit(v2[j].x);
                        v2[j].y = kF16 ? gelu_f16_mode(v2[j].y) : gelu_tanh_fit(v2[j].y);
                    }
                }
                if (kRes) {
                    mbar_wait(res_bar, (res_phase >> buf) & 1u);
                    res_phase ^= 1u << buf;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 u = *reinterpret_cast<const uint4*>(my_row + ((j ^ sw) << 4));
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                        if constexpr (kLnRes) {
                            // out = acc + bias + LN(R) = acc + (bias + beta) + gamma (rstd R - mean rstd):
                            // two packed FMAs and one packed add per pair of columns
                            float4 g0, g1, b0, b1;  // gamma / (bias + beta) of the 8 columns of this 16-byte piece
                            if constexpr (kColSmem) {
                                const float4* gs = reinterpret_cast<const float4*>(col_gamma + col0 + 8 * j);
                                const float4* bs = reinterpret_cast<const float4*>(col_bb + col0 + 8 * j);
                                g0 = gs[0]; g1 = gs[1]; b0 = bs[0]; b1 = bs[1];
                            } else {
                                const float4* gp = reinterpret_cast<const float4*>(fold.gamma + col0 + 8 * j);
                                const float4* tp = reinterpret_cast<const float4*>(fold.beta + col0 + 8 * j);
                                const float4* bp = reinterpret_cast<const float4*>(bias + col0 + 8 * j);
                                g0 = __ldg(gp); g1 = __ldg(gp + 1);
                                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), c0 = __ldg(bp), c1 = __ldg(bp + 1);
                                b0 = make_float4(t0.x + c0.x, t0.y + c0.y, t0.z + c0.z, t0.w + c0.w);
                                b1 = make_float4(t1.x + c1.x, t1.y + c1.y, t1.z + c1.z, t1.w + c1.w);
                            }
                            const float2 gg[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
                            const float2 bbv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 t = __ffma2_rn(rstd2, unpack16x2<kF16>(w[q]), nmr2);
                                v2[4 * j + q] = __ffma2_rn(gg[q], t, __fadd2_rn(v2[4 * j + q], bbv[q]));
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) v2[4 * j + q] = __fadd2_rn(v2[4 * j + q], unpack16x2<kF16>(w[q]));
                        }
                    }
                } else {
                    // the store that last read this staging tile has drained
                    if (leader) tma_store_wait_read<kBufs - 1>();
                    named_bar_sync(1 + grp, 128);
                }
                if constexpr (sizeof(OutT) == 2) {
                    float2 osum2 = make_float2(0.f, 0.f), osq2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4 u;
                        u.x = pack16x2<kF16>(v2[4 * j + 0].x, v2[4 * j + 0].y);
                        u.y = pack16x2<kF16>(v2[4 * j + 1].x, v2[4 * j + 1].y);
                        u.z = pack16x2<kF16>(v2[4 * j + 2].x, v2[4 * j + 2].y);
                        u.w = pack16x2<kF16>(v2[4 * j + 3].x, v2[4 * j + 3].y);
                        *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = u;
                        if constexpr (kStats) {
                            // row statistics from the fp32 values: the 16-bit rounding the consumer sees
                            // is zero-mean noise of 2^-9 relative size, far below the LayerNorm's own eps scale
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                osum2 = __fadd2_rn(osum2, v2[4 * j + q]);
                                osq2 = __ffma2_rn(v2[4 * j + q], v2[4 * j + q], osq2);
                            }
                        }
                    }
                    if constexpr (kStats) {
                        osum += osum2.x + osum2.y;
                        osq += osq2.x + osq2.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(my_row + ((j ^ sw) << 4)) =
                            make_float4(v2[2 * j].x, v2[2 * j].y, v2[2 * j + 1].x, v2[2 * j + 1].y);
                }
                fence_proxy_async_smem();
                named_bar_sync(1 + grp, 128);  // every row of the staging tile is written
                if (leader) {
                    tma_store_2d(&tmap_c, stage_tile, col0, row_a);
                    tma_store_commit();
                    if (kRes && kBufs == 2) {
                        // every store but the one just issued has been read: the other tile is free
                        tma_store_wait_read<1>();
                        prefetch_next_residual();
                    }
                }
            }
            if constexpr (kStats && !kNarrow) {
                // this thread covered columns [row_b + 128 grp, +128) of its row: one partial
                if (any_live && grow < M)
                    fold.stats_out[static_cast<int64_t>((row_b >> 7) + grp) * M + grow] = make_float2(osum, osq);
            }
            if constexpr (kStats && kNarrow) {
                // the groups covered 64 columns each: group 1 hands its sums over through shared memory
                // and group 0 writes the partial of the 128 columns (second barrier: the slots are free again)
                float2* comb = reinterpret_cast<float2*>(sm.pre() + 2 * kStageTileBytes);
                if (grp == 1) comb[trow] = make_float2(osum, osq);
                named_bar_sync(4, kEpiThreads);
                if (grp == 0 && any_live && grow < M) {
                    const float2 o = comb[trow];
                    fold.stats_out[static_cast<int64_t>(row_b >> 7) * M + grow] = make_float2(osum + o.x, osq + o.y);
                }
                named_bar_sync(4, kEpiThreads);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (leader) tma_store_wait<0>();
    }
    if constexpr (k2Cta) pipe2_teardown(sm, warp, tmem_base);
    else pipe_teardown(sm, warp, tmem_base);
}

template <int EPI, bool kF16, typename OutT>
static int launch_gemm_impl(const h16* A, int64_t lda, const h16* B, int64_t ldb, OutT* C,
                            int64_t ldc, const float* bias, const h16* R, int64_t ldr, int64_t M,
                            int N, int K, cudaStream_t stream, const LnFoldArgs& fold = LnFoldArgs()) {
    constexpr bool kRes = EPI == EPI_BIAS_RESIDUAL || EPI == EPI_BIAS_LNRES_STATS || EPI == EPI_BIAS_RES_STATS;
    CUtensorMap ta, tb, tc, tr;
    // the TMA element type only matters for OOB fill; both 16-bit formats move as raw 2-byte words
    bool ok = make_tmap_bf16_k64(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBM) &&
              make_tmap_bf16_k64(&tb, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), kGemmBN) &&
              make_tmap_rows128(&tc, C, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldc), sizeof(OutT));
    if (ok) ok = make_tmap_rows128(&tr, kRes ? static_cast<const void*>(R) : static_cast<const void*>(C),
                                   static_cast<uint64_t>(M), static_cast<uint64_t>(N),
                                   static_cast<uint64_t>(kRes ? ldr : ldc),
                                   kRes ? 2 : static_cast<int>(sizeof(OutT)));
    if (!ok) {
        set_error("cuTensorMapEncodeTiled failed (A %p lda %lld, B %p ldb %lld, C %p ldc %lld)", (const void*)A,
                  (long long)lda, (const void*)B, (long long)ldb, (const void*)C, (long long)ldc);
        return ARB_ERR_CUDA;
    }
    // auto: CTA pairs from 32 row blocks up (each SM stages half of B: measured 6-8 % faster per forward
    // than single CTAs at 4 k-16 k tokens, slower at 3 k); below that single CTAs, narrow tiles when few
    const bool pair = g_gemm_mode == 2 || (g_gemm_mode == 0 && M >= kGemmPairMinRows);
    if (pair) {
        // clusters of two CTAs, one 256 x 256 tile per pair; B is staged in halves of 128 rows
        if (!make_tmap_bf16_k64(&tb, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), kGemmBN / 2)) {
            set_error("cuTensorMapEncodeTiled failed (B half tile)");
            return ARB_ERR_CUDA;
        }
        constexpr bool kCanColSmem = EPI == EPI_BIAS_LNRES_STATS && sizeof(OutT) == 2;
        const bool col_smem = kCanColSmem && N <= kColSmemMaxN && K <= 1024 && g_gemm_colsmem;
        auto kern = gemm16_kernel<EPI, kF16, OutT, true>;
        int smem = Gemm2Smem::kExtraOffset + 1024;
        static_assert(Gemm2Smem::kExtraOffset + 1024 <= 232448, "pair GEMM shared memory exceeds 227 KB");
        if constexpr (kCanColSmem) {
            static_assert(Gemm2ColSmem::kExtraOffset + 1024 <= 232448, "pair GEMM (column vectors) shared memory exceeds 227 KB");
            if (col_smem) {
                kern = gemm16_kernel<EPI, kF16, OutT, true, true>;
                smem = Gemm2ColSmem::kExtraOffset + 1024;
            }
        }
        ARB_CHECK_CUDA(set_max_smem_once(kern, smem));
        const int64_t tiles = ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + kGemmBN - 1) / kGemmBN);
        int64_t nclusters = num_sms() / 2;
        if (nclusters > tiles) nclusters = tiles;
        ARB_CHECK_CUDA(launch_kernel(kern, dim3(static_cast<unsigned>(nclusters * 2)), dim3(kGemmThreads), smem, stream, 2, ta, tb,
                                     tc, tr, bias, M, N, K, fold, umma_idesc_16bit(2 * kBM, kGemmBN, kF16), 0u));
        return ARB_OK;
    }
    if constexpr (sizeof(OutT) == 2) {
        // a handful of row blocks (query-time batches): narrow tiles put twice as many SMs on the weights
        const int64_t wide_tiles = ((M + kBM - 1) / kBM) * ((N + kGemmBN - 1) / kGemmBN);
        if (g_gemm_mode == 3 || (g_gemm_mode == 0 && g_gemm_narrow && wide_tiles * 2 <= num_sms())) {
            if (!make_tmap_bf16_k64(&tb, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), kGemmNarrowBN)) {
                set_error("cuTensorMapEncodeTiled failed (B narrow tile)");
                return ARB_ERR_CUDA;
            }
            // fewer than 128 rows in all: load only those (rounded up to a swizzle atom). Half of every
            // stage would otherwise be zero fill, through TMA's slow out-of-bounds path at that.
            uint32_t a_box_bytes = 0;
            if (M < kBM) {
                const uint32_t a_rows = static_cast<uint32_t>((M + 7) / 8 * 8);
                if (a_rows < static_cast<uint32_t>(kBM)) {
                    if (!make_tmap_bf16_k64(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), a_rows)) {
                        set_error("cuTensorMapEncodeTiled failed (A short tile)");
                        return ARB_ERR_CUDA;
                    }
                    a_box_bytes = a_rows * 128u;
                }
            }
            auto nkern = gemm16_kernel<EPI, kF16, OutT, false, false, kGemmNarrowBN>;
            constexpr int nsmem = GemmNarrowSmem::kExtraOffset + 1024;
            static_assert(nsmem <= 232448, "narrow GEMM shared memory exceeds 227 KB");
            ARB_CHECK_CUDA(set_max_smem_once(nkern, nsmem));
            const int64_t ntiles = ((M + kBM - 1) / kBM) * ((N + kGemmNarrowBN - 1) / kGemmNarrowBN);
            const int64_t ngrid = ntiles < num_sms() ? ntiles : num_sms();
            ARB_CHECK_CUDA(launch_kernel(nkern, dim3(static_cast<unsigned>(ngrid)), dim3(kGemmThreads), nsmem, stream, 1, ta, tb, tc, tr,
                                         bias, M, N, K, fold, umma_idesc_16bit(kBM, kGemmNarrowBN, kF16), a_box_bytes));
            return ARB_OK;
        }
    }
    auto kern = gemm16_kernel<EPI, kF16, OutT, false>;
    constexpr int smem = GemmSmem::kExtraOffset + 1024;  // +1024: alignment slack
    static_assert(smem <= 232448, "GEMM shared memory exceeds 227 KB");
    ARB_CHECK_CUDA(set_max_smem_once(kern, smem));
    const int64_t tiles = ((M + kBM - 1) / kBM) * ((N + kGemmBN - 1) / kGemmBN);
    const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(kGemmThreads), smem, stream, 1, ta, tb, tc, tr, bias, M, N, K, fold,
                                 umma_idesc_16bit(kBM, kGemmBN, kF16), 0u));
    return ARB_OK;
}

static int check_gemm_args(const void* A, int64_t lda, const void* B, int64_t ldb, const void* C,
                           int64_t ldc, int64_t M, int N, int K, int c_elt_bytes) {
    ARB_REQUIRE(A && B && C, "gemm: null operand");
    ARB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%lld N=%d K=%d", (long long)M, N, K);
    ARB_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
    ARB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && (ldc * c_elt_bytes) % 16 == 0,
                "gemm: K/lda/ldb must be multiples of 8 and C rows 16-byte multiples");
    ARB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                "gemm: operands must be 16-byte aligned");
    ARB_REQUIRE(M < (1ll << 31), "gemm: M too large");
    return ARB_OK;
}

template <bool kF16>
static int dispatch_gemm16(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                           const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                           int epilogue, cudaStream_t stream) {
    switch (epilogue) {
        case EPI_BIAS:
            return launch_gemm_impl<EPI_BIAS, kF16, h16>(A, lda, B, ldb, C, ldc, bias, nullptr, 0, M, N, K, stream);
        case EPI_BIAS_GELU:
            return launch_gemm_impl<EPI_BIAS_GELU, kF16, h16>(A, lda, B, ldb, C, ldc, bias, nullptr, 0, M, N, K, stream);
        case EPI_BIAS_RESIDUAL:
            ARB_REQUIRE(R != nullptr && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0,
                        "gemm: residual operand missing or misaligned");
            return launch_gemm_impl<EPI_BIAS_RESIDUAL, kF16, h16>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, stream);
        default:
            set_error("gemm: unknown epilogue %d", epilogue);
            return ARB_ERR_INVALID;
    }
}

int launch_gemm16(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                  const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                  int epilogue, bool fp16, cudaStream_t stream) {
    int rc = check_gemm_args(A, lda, B, ldb, C, ldc, M, N, K, 2);
    if (rc) return rc;
    ARB_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                "gemm: bias must be 16-byte aligned");
    return fp16 ? dispatch_gemm16<true>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, epilogue, stream)
                            : dispatch_gemm16<false>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, epilogue, stream);
}

template <bool kF16>
static int dispatch_gemm16_fold(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                                const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                                int epilogue, const LnFoldArgs& f, cudaStream_t stream) {
    switch (epilogue) {
        case EPI_LNIN_BIAS:
            return launch_gemm_impl<EPI_LNIN_BIAS, kF16, h16>(A, lda, B, ldb, C, ldc, bias, nullptr, 0, M, N, K, stream, f);
        case EPI_LNIN_BIAS_GELU:
            return launch_gemm_impl<EPI_LNIN_BIAS_GELU, kF16, h16>(A, lda, B, ldb, C, ldc, bias, nullptr, 0, M, N, K, stream, f);
        case EPI_BIAS_LNRES_STATS:
            return launch_gemm_impl<EPI_BIAS_LNRES_STATS, kF16, h16>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, stream, f);
        default:
            return launch_gemm_impl<EPI_BIAS_RES_STATS, kF16, h16>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, stream, f);
    }
}

int launch_gemm16_fold(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                       const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                       int epilogue, const LnFoldArgs& f, bool fp16, cudaStream_t stream) {
    int rc = check_gemm_args(A, lda, B, ldb, C, ldc, M, N, K, 2);
    if (rc) return rc;
    ARB_REQUIRE(epilogue >= EPI_LNIN_BIAS && epilogue <= EPI_BIAS_RES_STATS, "gemm_fold: unknown epilogue %d", epilogue);
    ARB_REQUIRE(bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm_fold: bias missing or misaligned");
    const bool ln_in = epilogue == EPI_LNIN_BIAS || epilogue == EPI_LNIN_BIAS_GELU;
    const bool stats = epilogue == EPI_BIAS_LNRES_STATS || epilogue == EPI_BIAS_RES_STATS;
    if (ln_in) {
        ARB_REQUIRE(N % 64 == 0, "gemm_fold: N=%d must be a multiple of 64", N);
        ARB_REQUIRE(f.colsum && (reinterpret_cast<uintptr_t>(f.colsum) & 15) == 0, "gemm_fold: colsum missing or misaligned");
    }
    if (stats) {
        ARB_REQUIRE(N % 128 == 0, "gemm_fold: N=%d must be a multiple of 128 for row statistics", N);
        ARB_REQUIRE(f.stats_out != nullptr, "gemm_fold: stats_out missing");
        ARB_REQUIRE(R != nullptr && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0,
                    "gemm_fold: residual operand missing or misaligned");
    }
    if (ln_in || epilogue == EPI_BIAS_LNRES_STATS)
        ARB_REQUIRE(f.stats_in != nullptr && f.parts_in > 0 && f.inv_width_in > 0.f, "gemm_fold: input row statistics missing");
    if (epilogue == EPI_BIAS_LNRES_STATS)
        ARB_REQUIRE(f.gamma && f.beta && (reinterpret_cast<uintptr_t>(f.gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(f.beta) & 15) == 0,
                    "gemm_fold: gamma / beta missing or misaligned");
    return fp16 ? dispatch_gemm16_fold<true>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, epilogue, f, stream)
                            : dispatch_gemm16_fold<false>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, epilogue, f, stream);
}

int launch_gemm16_f32out(const h16* A, int64_t lda, const h16* B, int64_t ldb, float* C,
                         int64_t ldc, int64_t M, int N, int K, bool fp16, cudaStream_t stream) {
    int rc = check_gemm_args(A, lda, B, ldb, C, ldc, M, N, K, 4);
    if (rc) return rc;
    return fp16 ? launch_gemm_impl<EPI_BIAS, true, float>(A, lda, B, ldb, C, ldc, nullptr, nullptr, 0, M, N, K, stream)
                            : launch_gemm_impl<EPI_BIAS, false, float>(A, lda, B, ldb, C, ldc, nullptr, nullptr, 0, M, N, K, stream);
}

}  // namespace arb
