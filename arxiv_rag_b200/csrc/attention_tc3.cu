// tcgen05/TMEM self-attention, 16 softmax warps (experiment; same contract and the same MMA schedule
// as attention_tc.cu: modeling_mpnet.py:162-177, :324-360). attention_tc.cu runs 8 softmax warps —
// two per scheduler — and every pipe sits below 40 %: the per-row chains are latency-bound. Here each
// key half is served by EIGHT warps: the two warps that share a TMEM lane quarter split the half's
// columns, so a thread owns one query row x Kh/2 keys and four warps per scheduler hide each other's
// TMEM / shared-memory / MUFU latencies.
//   * common shift per half: the two column threads of a row exchange their raw maxima through
//     shared memory (one named barrier of the half's 256 threads per tile);
//   * P still aliases S, but each column thread stores its P inside its OWN S columns (a thread's
//     writes trail its reads; the sibling's unread scores are never touched), so P_h is two runs and
//     the P.V K-steps pick their run;
//   * every thread publishes (shift, partial row sum) before it arrives on p_ready; the combining
//     half reads all four after o_full — ordered through the MMA thread's barriers, so no barrier of
//     its own — and the two threads of a row store 32 output dims each.
// TMEM: S[0] cols 0..191, S[1] 192..383, O[0] 384..447, O[1] 448..511. Registers: 576 threads -> 96 each.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "ptx.cuh"

namespace arb {
namespace {

constexpr int kA3Threads = 64 + 512;  // warp 0 TMA, warp 1 MMA, warps 2-9 key half 0, 10-17 key half 1
constexpr int kA3QT = 128;
constexpr int kA3MaxKh = 192;
constexpr uint32_t kA3ColS = 0, kA3ColO = 384;
constexpr float kA3Log2e = 1.4426950408889634f;

__device__ __forceinline__ float a3_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ bool a3_bar_red_and(int id, int n, bool p) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(p)), "r"(id), "r"(n)
        : "memory");
    return out != 0;
}
__device__ __forceinline__ bool a3_bar_red_or(int id, int n, bool p) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(p)), "r"(id), "r"(n)
        : "memory");
    return out != 0;
}
__device__ __forceinline__ void a3_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void a3_wait_backoff(uint64_t* bar, uint32_t parity) {
#ifdef ARB_HANG_GUARD
    uint32_t spins = 0;
#endif
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
#ifdef ARB_HANG_GUARD
        if (++spins > (1u << 24)) {
            printf("arb: attention3 mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x,
                   (void*)bar, parity);
            __trap();
        }
#endif
    }
}

struct A3Layout {
    int kv_bytes;  // one buffer: K (2Kh rows) then V (2Kh rows), 128 B per row
    int q_off, bias_off, mask_off, exch_off, xmax_off, red_off, bar_off, total;
    int nbias;
};
__host__ __device__ inline A3Layout a3_layout(int Kh, int nqt) {
    A3Layout L;
    L.kv_bytes = 4 * Kh * 128;
    L.q_off = 2 * L.kv_bytes;
    L.nbias = nqt * kA3QT + 2 * Kh;
    L.bias_off = L.q_off + kA3QT * 128;
    L.mask_off = L.bias_off + 2 * (L.nbias + 2) * 4;   // two copies (shift 0 / shift 1)
    L.exch_off = L.mask_off;                            // (no per-key mask table: chunk flags + key bitmasks after `red`)
    L.xmax_off = L.exch_off + 2 * 2 * 2 * kA3QT * 8;    // [parity][half][cs][row] (c, partial l)
    L.red_off = L.xmax_off + 2 * 2 * kA3QT * 4;         // [half][cs][row] raw maximum
    L.bar_off = (L.red_off + 64 * 4 + 16 + 64 + 7) & ~7;  // + mask flags [half][8] (bytes) and key bitmasks [half][8]
    L.total = L.bar_off + 16 * 8 + 16;
    return L;
}

// One thread's share of a score half-tile: `nloc` 32-column chunks starting at TMEM address tS.
// Pass 1 (a3_rowmax) returns the raw maximum over the unmasked keys; pass 2 (a3_emit) turns the chunks
// into P in place and returns the partial row sum. 18 warps leave 96 registers per thread (allocation
// is per 4 warps: 576 threads count as 640): one 32-column register buffer, no double buffering —
// the three other warps of the scheduler cover a chunk's TMEM latency (double-buffered 16-column
// pieces, which also fit, measured 1.070 vs 1.058 ms at S = 384).
template <bool kMask>
__device__ __forceinline__ float a3_rowmax(uint32_t tS, int nloc, const float* __restrict__ pm) {
    float mraw = -INFINITY;
    for (int cc = 0; cc < nloc; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tS + cc * 32, v);
        tmem_ld_wait();
        if (!kMask) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(v[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(v[j]) + pm[cc * 32 + j]);
        }
    }
    return mraw;
}

// rel_lo / rel_hi: the relative positions (key - query) of this WARP's 32 rows x the thread's first
// chunk span [rel_lo, rel_hi] (+32 per chunk). A relative-position table is constant outside
// [rneg, rpos] (MPNet's buckets end at |j - i| = 91), so a chunk that lies wholly outside takes ONE
// table read instead of sixteen: the shared-memory load pipe is the busiest unit of this kernel
// (55 % of its wavefront peak, and `mio_throttle` is what the extra warps wait on).
template <bool kF16, bool kMask>
__device__ __forceinline__ float a3_emit(uint32_t tS, int nloc, float scale, float c, const float2* __restrict__ pb2,
                                         const float* __restrict__ pb1, const float* __restrict__ pm, int rel_lo,
                                         int rel_hi, int rneg, int rpos) {
    float l = 0.f;
    for (int cc = 0; cc < nloc; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tS + cc * 32, v);
        const bool flat = rel_lo + cc * 32 >= rpos || rel_hi + cc * 32 <= rneg;  // warp-uniform
        const float bflat = pb1[cc * 32] - c;
        tmem_ld_wait();
        uint32_t pk[16];
        if (flat) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float x0 = fmaf(__uint_as_float(v[j]), scale, bflat);
                float x1 = fmaf(__uint_as_float(v[j + 1]), scale, bflat);
                if (kMask) {
                    x0 += pm[cc * 32 + j];
                    x1 += pm[cc * 32 + j + 1];
                }
                const float p0 = a3_exp2(x0), p1 = a3_exp2(x1);
                l += p0 + p1;
                pk[j >> 1] = pack16x2<kF16>(p0, p1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 bb = pb2[(cc * 32 + j) >> 1];
                float x0 = fmaf(__uint_as_float(v[j]), scale, bb.x) - c;
                float x1 = fmaf(__uint_as_float(v[j + 1]), scale, bb.y) - c;
                if (kMask) {
                    x0 += pm[cc * 32 + j];
                    x1 += pm[cc * 32 + j + 1];
                }
                const float p0 = a3_exp2(x0), p1 = a3_exp2(x1);
                l += p0 + p1;
                pk[j >> 1] = pack16x2<kF16>(p0, p1);
            }
        }
        tmem_st_32x16(tS + cc * 16, pk);  // inside the columns this thread has already read
    }
    return l;
}

// Masked keys: see attention_tc.cu (mask_prepass / mask_postpass) — partly masked chunks are rewritten to
// -inf in TMEM first, the unmasked passes then run over the chunks up to the last one with a valid key,
// and P of the chunks behind it is zeroed afterwards.
__device__ __forceinline__ void a3_mask_prepass(uint32_t tS, int n_eff, const uint32_t* __restrict__ cbits,
                                                const uint8_t* __restrict__ cflag) {
    for (int cc = 0; cc < n_eff; ++cc) {
        if (cflag[cc] == 0) continue;
        const uint32_t w = cbits[cc];
        uint32_t v[32];
        tmem_ld_32x32(tS + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (w >> j) & 1u ? v[j] : 0xff800000u;
        tmem_st_32x16(tS + cc * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_st_32x16(tS + cc * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        tmem_st_wait();  // the main pass reads these columns again
    }
}
__device__ __forceinline__ void a3_mask_postpass(uint32_t tS, int n_eff, int nloc) {
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    for (int cc = n_eff; cc < nloc; ++cc) tmem_st_32x16(tS + cc * 16, z);
}

template <bool kF16>
__global__ void __launch_bounds__(kA3Threads, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const float* __restrict__ rel_bias, int max_rel, const int32_t* __restrict__ mask,
                     h16* __restrict__ ctx, int B, int S, int heads, int Kh, float scale_log2e) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nqt = (S + kA3QT - 1) / kA3QT;
    const A3Layout L = a3_layout(Kh, nqt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_empty = bars + 2;   // [2]
    uint64_t* q_full = bars + 4;
    uint64_t* q_empty = bars + 5;
    uint64_t* s_full = bars + 6;     // [2]
    uint64_t* p_ready = bars + 8;    // [2]
    uint64_t* o_full = bars + 10;    // [2]
    uint64_t* o_free = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
    const int lane = threadIdx.x & 31;
    const int H = heads * 64;
    const int h = static_cast<int>(blockIdx.x) % heads;
    const int b_first = static_cast<int>(blockIdx.x) / heads;
    const int ngroups = static_cast<int>(gridDim.x) / heads;
    const int OFF = nqt * kA3QT;  // table entry e <-> (j - i) = e - OFF
    const int nchunk = Kh / 32;
    const int n0 = (nchunk + 1) / 2;  // chunks of column thread 0; thread 1 takes the rest

    float* T0 = reinterpret_cast<float*>(sm + L.bias_off);
    float* T1 = T0 + L.nbias + 2;
    float* red = reinterpret_cast<float*>(sm + L.red_off);  // [0..31] max, [32..63] min per warp
    int* s_flat = reinterpret_cast<int*>(bars + 14);        // [0] rpos, [1] rneg: the table is constant outside [rneg, rpos]
    if (threadIdx.x == 0) {
        s_flat[0] = 0;
        s_flat[1] = 0;
    }
    __syncthreads();
    {
        float bm = -INFINITY, bn = INFINITY;
        const float* tb = rel_bias + static_cast<int64_t>(h) * (2 * max_rel - 1) + (max_rel - 1);
        for (int e = threadIdx.x; e < L.nbias + 2; e += kA3Threads) {
            const int rel = e - OFF;
            float v = 0.f;
            if (e < L.nbias && rel > -S && rel < S) {
                const float raw = tb[rel];
                v = raw * kA3Log2e;
                bm = fmaxf(bm, v);
                bn = fminf(bn, v);
                if (rel >= 0 && rel + 1 < S && tb[rel + 1] != raw) atomicMax(s_flat + 0, rel + 1);
                if (rel <= 0 && rel - 1 > -S && tb[rel - 1] != raw) atomicMin(s_flat + 1, rel - 1);
            }
            T0[e] = v;
            if (e >= 1) T1[e - 1] = v;
        }
        if (threadIdx.x == 0) T1[L.nbias + 1] = 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffff, bm, o));
            bn = fminf(bn, __shfl_xor_sync(0xffffffff, bn, o));
        }
        if (lane == 0) {
            red[warp] = bm;
            red[32 + warp] = bn;
        }
    }
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(kv_full + i, 1);
            mbar_init(kv_empty + i, 1);
            mbar_init(s_full + i, 1);
            mbar_init(p_ready + i, 256);
            mbar_init(o_full + i, 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        mbar_init(o_free, 256);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();
    float bmax = red[0], bmin = red[32];
#pragma unroll
    for (int w = 1; w < kA3Threads / 32; ++w) {
        bmax = fmaxf(bmax, red[w]);
        bmin = fminf(bmin, red[32 + w]);
    }
    if (kF16) bmax -= fminf(bmax - bmin, 15.f);  // see attention_tc.cu: keeps fp16 P normal for wide bias tables
    const int rpos = s_flat[0], rneg = s_flat[1];

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int n = 0, g = 0;
            for (int b = b_first; b < B; b += ngroups, ++n) {
                const int buf = n & 1;
                uint8_t* K = sm + buf * L.kv_bytes;
                uint8_t* V = K + 2 * Kh * 128;
                a3_wait_backoff(kv_empty + buf, ((n >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full + buf, L.kv_bytes);
                tma_load_3d(&tmap_kv, kv_full + buf, K, H + h * 64, 0, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, K + Kh * 128, H + h * 64, Kh, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, V, 2 * H + h * 64, 0, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, V + Kh * 128, 2 * H + h * 64, Kh, b, kEvictFirst);
                for (int t = 0; t < nqt; ++t, ++g) {
                    a3_wait_backoff(q_empty, (g & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full, kA3QT * 128);
                    tma_load_3d(&tmap_q, q_full, sm + L.q_off, h * 64, t * kA3QT, b, kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc_qk = umma_idesc_16bit(kA3QT, Kh, kF16);
            const uint32_t idesc_pv = umma_idesc_16bit_bmn(kA3QT, 64, kF16);
            const uint64_t dq = umma_desc_sw128(smem_u32(sm + L.q_off));
            const int my_items = B > b_first ? (B - 1 - b_first) / ngroups + 1 : 0;
            const int G = my_items * nqt;
            const uint32_t sm_base = smem_u32(sm);
            auto issue_qk = [&](int tile, int hh) {
                const int n = tile / nqt;
                const uint32_t k_addr = sm_base + (n & 1) * L.kv_bytes;
                const uint64_t dk = umma_desc_sw128(k_addr + hh * Kh * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem + kA3ColS + hh * kA3MaxKh, dq + 2 * k, dk + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                umma_commit(s_full + hh);
            };
            auto issue_pv = [&](int tile, int hh) {
                const int n = tile / nqt;
                const uint32_t v_addr = sm_base + (n & 1) * L.kv_bytes + 2 * Kh * 128;
                const uint64_t dv = umma_desc_sw128(v_addr + hh * Kh * 128);
                for (int kk = 0; kk < Kh / 16; ++kk) {
                    // P of chunk ch sits at the start of its column thread's own S columns
                    const int ch = kk >> 1;
                    const int pcol = ch < n0 ? ch * 16 : n0 * 32 + (ch - n0) * 16;
                    umma_bf16_ts(tmem + kA3ColO + hh * 64, tmem + kA3ColS + hh * kA3MaxKh + pcol + (kk & 1) * 8,
                                 dv + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_pv, kk > 0 ? 1u : 0u);
                }
                umma_commit(o_full + hh);
            };
            auto wait_inputs = [&](int tile) {
                const int n = tile / nqt;
                if (tile % nqt == 0) mbar_wait(kv_full + (n & 1), (n >> 1) & 1);
                mbar_wait(q_full, tile & 1);
                tc_fence_after();
            };
            if (G > 0) {
                wait_inputs(0);
                issue_qk(0, 0);
                issue_qk(0, 1);
                umma_commit(q_empty);
            }
            for (int g = 0; g < G; ++g) {
                const uint32_t ph = g & 1;
                const bool has_next = g + 1 < G;
                mbar_wait(p_ready + 0, ph);
                mbar_wait(o_free, ph ^ 1);  // the combining half has read O of tile g-1
                tc_fence_after();
                issue_pv(g, 0);
                if (has_next) {
                    wait_inputs(g + 1);
                    issue_qk(g + 1, 0);
                }
                mbar_wait(p_ready + 1, ph);
                tc_fence_after();
                issue_pv(g, 1);
                if (has_next) {
                    issue_qk(g + 1, 1);
                    umma_commit(q_empty);
                }
                if (g % nqt == nqt - 1) umma_commit(kv_empty + ((g / nqt) & 1));
            }
        }
    } else {
        // ===================== softmax: 8 warps per key half =====================
        const int hh = (warp - 2) >> 3;             // key half
        const int cs = ((warp - 2) & 7) >> 2;       // column thread of the row (0: chunks [0, n0), 1: the rest)
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;          // query row inside the tile
        const int bar_id = 1 + hh;
        const int cb = cs == 0 ? 0 : n0;            // first chunk of this thread
        const int nloc = cs == 0 ? n0 : nchunk - n0;
        uint8_t* cfl_half = reinterpret_cast<uint8_t*>(sm + L.red_off + 64 * 4) + hh * 8;           // this half's chunk flags
        uint32_t* cbits_half = reinterpret_cast<uint32_t*>(sm + L.red_off + 64 * 4 + 16) + hh * 8; // and key bitmasks
        const uint8_t* cflag = cfl_half + cb;
        const uint32_t* cbits = cbits_half + cb;
        float2* exch = reinterpret_cast<float2*>(sm + L.exch_off);         // [parity][half][cs][row]
        float* xmax = reinterpret_cast<float*>(sm + L.xmax_off);           // [half][cs][row]
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tS = tmem + lane_sel + kA3ColS + hh * kA3MaxKh + cb * 32;
        int g = 0;
        for (int b = b_first; b < B; b += ngroups) {
            a3_bar_sync(bar_id, 256);  // everyone is done with the previous item's table
            bool mine_clear = true, mine_on = false;
            for (int j = cs * 128 + r; j < 2 * Kh; j += 256) {  // a warp covers one 32-key chunk per step
                const bool on = j < S && mask[static_cast<int64_t>(b) * S + j] != 0;
                const uint32_t bal = __ballot_sync(0xffffffff, on);
                if (lane == 0 && j >= hh * Kh && j < (hh + 1) * Kh) {
                    cfl_half[(j - hh * Kh) >> 5] = bal == 0xffffffffu ? 0 : (bal != 0u ? 1 : 2);
                    cbits_half[(j - hh * Kh) >> 5] = bal;
                }
                mine_clear &= on;
                mine_on |= on;
            }
            const bool clear = a3_bar_red_and(bar_id, 256, mine_clear);  // no masked / out-of-range key at all
            const bool any_on = a3_bar_red_or(bar_id, 256, mine_on);    // (the barriers also publish the flags)
            int n_eff = nloc;  // this thread's chunks up to the last one with a valid key
            if (!clear)
                while (n_eff > 0 && cflag[n_eff - 1] == 2) --n_eff;
            for (int t = 0; t < nqt; ++t, ++g) {
                const uint32_t ph = g & 1;
                const int i = t * kA3QT + r;
                const int start = hh * Kh - i + OFF;  // >= 1; bias of key column c of the half is T0[start + c]
                const float2* pb2 = reinterpret_cast<const float2*>((start & 1) ? T1 + (start - 1) : T0 + start) + cb * 16;
                mbar_wait(s_full + hh, ph);
                tc_fence_after();
                float c = 0.f, l = 0.f;
                if (!any_on) {
                    // every key masked -> uniform attention over the S keys (see attention_tc.cu)
                    for (int cc = 0; cc < nloc; ++cc) {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const int key = hh * Kh + (cb + cc) * 32 + j;
                            const float p0 = key < S ? 1.f : 0.f, p1 = key + 1 < S ? 1.f : 0.f;
                            l += p0 + p1;
                            pk[j >> 1] = pack16x2<kF16>(p0, p1);
                        }
                        tmem_st_32x16(tS + cc * 16, pk);
                    }
                } else {
                    if (!clear) a3_mask_prepass(tS, n_eff, cbits, cflag);
                    const float mine = a3_rowmax<false>(tS, n_eff, nullptr);
                    xmax[(hh * 2 + cs) * kA3QT + r] = mine;
                    a3_bar_sync(bar_id, 256);  // the row's other column thread has published its maximum
                    const float mraw = fmaxf(mine, xmax[(hh * 2 + (cs ^ 1)) * kA3QT + r]);
                    c = (mraw == -INFINITY) ? 0.f : fmaf(mraw, scale_log2e, bmax);
                    // relative positions (key - query) of this warp's 32 rows x this thread's first chunk
                    const int key0 = hh * Kh + cb * 32, row0 = t * kA3QT + quarter * 32;
                    const float* pb1 = T0 + start + cb * 32;
                    l = a3_emit<kF16, false>(tS, n_eff, scale_log2e, c, pb2, pb1, nullptr, key0 - (row0 + 31), key0 + 31 - row0, rneg, rpos);
                    if (!clear) a3_mask_postpass(tS, n_eff, nloc);
                }
                exch[((ph * 2 + hh) * 2 + cs) * kA3QT + r] = make_float2(c, l);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_ready + hh);  // release: P and the (c, l) above are ordered before the arrive
                if (hh == 0) continue;
                // ---- key half 1 combines: both halves' P.V done => every (c, l) of the tile is visible
                mbar_wait(o_full + 0, ph);
                mbar_wait(o_full + 1, ph);
                tc_fence_after();
                const float2* e = exch + (ph * 2) * 2 * kA3QT + r;
                const float2 e00 = e[0], e01 = e[kA3QT], e10 = e[2 * kA3QT], e11 = e[3 * kA3QT];
                const float c0 = e00.x, c1 = e10.x;
                const float m = fmaxf(c0, c1);
                const float a0 = a3_exp2(c0 - m), a1 = a3_exp2(c1 - m);
                const float inv = __fdividef(1.f, a0 * (e00.y + e01.y) + a1 * (e10.y + e11.y));
                const float w0 = a0 * inv, w1 = a1 * inv;
                const uint32_t tO = tmem + lane_sel + kA3ColO;
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(tO + cs * 32, o0);        // this column thread stores output dims [32 cs, 32 cs + 32)
                tmem_ld_32x32(tO + 64 + cs * 32, o1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(o_free);  // O is in registers: the next tile's P.V may overwrite it
                if (i < S) {
                    uint4* dst = reinterpret_cast<uint4*>(ctx + (static_cast<int64_t>(b) * S + i) * H + h * 64) + cs * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int ee = 0; ee < 4; ++ee) {
                            const int d = q * 8 + ee * 2;
                            w[ee] = pack16x2<kF16>(w0 * __uint_as_float(o0[d]) + w1 * __uint_as_float(o1[d]),
                                                   w0 * __uint_as_float(o0[d + 1]) + w1 * __uint_as_float(o1[d + 1]));
                        }
                        dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace

bool attention_tc3_supported(int S, int dh) { return dh == 64 && S >= 1 && S <= 2 * kA3MaxKh; }

int launch_attention_tc3(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                         h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream) {
    ARB_REQUIRE(qkv && rel_bias && mask && ctx, "attention_tc3: null pointer");
    ARB_REQUIRE(attention_tc3_supported(S, dh), "attention_tc3: S=%d dh=%d unsupported", S, dh);
    ARB_REQUIRE(B > 0 && S <= max_rel, "attention_tc3: bad shape B=%d S=%d max_rel=%d", B, S, max_rel);
    const int H = heads * dh;
    const int Kh = ((S + 1) / 2 + 31) / 32 * 32;
    const int nqt = (S + kA3QT - 1) / kA3QT;
    const A3Layout L = a3_layout(Kh, nqt);
    const int smem = L.total + 1024;
    ARB_REQUIRE(smem <= 232448, "attention_tc3: shared memory %d exceeds 227 KB", smem);
    CUtensorMap tq, tkv;
    if (!make_tmap_bf16_batched_k64(&tq, qkv, B, S, 3 * H, 3 * H, kA3QT) ||
        !make_tmap_bf16_batched_k64(&tkv, qkv, B, S, 3 * H, 3 * H, Kh)) {
        set_error("attention_tc3: cuTensorMapEncodeTiled failed");
        return ARB_ERR_CUDA;
    }
    auto kern = fp16 ? attention_tc3_kernel<true> : attention_tc3_kernel<false>;
    ARB_CHECK_CUDA(set_max_smem_once(kern, smem));
    int ngroups = num_sms() / heads;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > B) ngroups = B;
    const int grid = ngroups * heads;
    const float scale_log2e = kA3Log2e / sqrtf(static_cast<float>(dh));
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(kA3Threads), smem, stream, 1, tq, tkv, rel_bias, max_rel, mask, ctx, B, S,
                                 heads, Kh, scale_log2e));
    return ARB_OK;
}

}  // namespace arb
