// tcgen05/TMEM self-attention for the MPNet encoder, second schedule (S <= 384, head dim 64):
//   softmax(q.k^T/8 + rel_bias[h][j-i] + mask) . v   per (sequence, head)
// Same contract as attention_tc.cu / attention_mma.cu (modeling_mpnet.py:162-177, :324-360; mask of
// modeling_utils.py:936-947).
//
// What changed against attention_tc.cu, and why (ncu source view, profiles/r2_attention_notes.md).
// There, a softmax group (4 warps, one thread per query row) owns one half of the keys of every
// 128-query tile and works through it in phases: read the scores back from TMEM, take the row
// maximum, add the relative bias (one shared-memory load per pair of scores), exponentiate, store
// P; then it waits for P.V and the next tile's Q.K^T of that half. Only the exp2 phase is bound by
// a pipe (MUFU, 8 cycles per warp instruction); the others are bound by the latency of their own
// loads and waits, and with two warps per scheduler nothing fills those gaps: MUFU 36 % busy.
// Here each half is cut into FOUR sub-blocks of Kb = Kh/4 <= 48 keys that are scored, soft-maxed
// and multiplied independently:
//     S_{h,sb} = Q . K_{h,sb}^T   (M = 128, N = Kb)            -> TMEM S[h][sb]
//     O_h (+)= P_{h,sb} . V_{h,sb}                              -> TMEM O[h]
// and a thread keeps TWO sub-blocks in registers: while it exponentiates sub-block k (the MUFU
// paces it), the same instruction stream reads sub-block k+1 from TMEM, adds scale and bias and
// reduces its maximum in the issue slots the MUFU leaves free. The tensor core turns a sub-block
// around (P.V of tile t, then Q.K^T of tile t+1 into the same TMEM columns) while the group works
// on the other three, so the next scores are waiting when the group comes back. Same TMEM
// (S 2x192 + O 2x64 = 512 columns) and the same shared memory as before.
// Within a half the sub-blocks share one accumulator, so the softmax is online across them:
// running maximum m and row sum l; a sub-block keeps m unless its own maximum exceeds it by more
// than kTau log2 units (then O_h and l are rescaled by the group itself — a rare path). The
// maximum is the exact one of s*scale + bias + mask, so probabilities never exceed 2^kTau and fp16
// mode needs no shift tricks. The two halves are merged as before,
//     O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),  a_h = exp2(m_h - max(m_0, m_1)),
// by the two groups in turn (even tiles: group 1, odd tiles: group 0).
// The scale/bias/shift arithmetic is packed (fma.rn.f32x2 / add.rn.f32x2).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "ptx.cuh"

namespace arb {

// Three warpgroups: warps 0-3 = roles (warp 0 TMA, warp 1 MMA, warps 2-3 idle), warps 4-7 softmax half
// 0, warps 8-11 softmax half 1. The softmax threads keep two 48-key score sub-blocks plus the
// accumulator halves of the merge in registers, more than the 168 a 384-thread CTA gets evenly: the
// role warpgroup hands its registers over with setmaxnreg.
constexpr int kA2Threads = 384;
constexpr int kA2RoleRegs = 80, kA2SoftmaxRegs = 200;
constexpr int kA2QT = 128;            // query rows per tile
constexpr int kA2MaxKh = 192;         // keys per half
constexpr int kA2Sub = 4;             // sub-blocks per half
constexpr uint32_t kA2ColS = 0, kA2ColO = 384;
constexpr float kA2Log2e = 1.4426950408889634f;
constexpr float kA2Tau = 8.f;         // lazy-rescale threshold (log2 units): p <= 2^8 fits fp16

__device__ __forceinline__ float a2_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ bool a2_bar_red_and(int id, int n, bool p) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(p)), "r"(id), "r"(n)
        : "memory");
    return out != 0;
}
__device__ __forceinline__ bool a2_bar_red_or(int id, int n, bool p) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(p)), "r"(id), "r"(n)
        : "memory");
    return out != 0;
}
__device__ __forceinline__ void a2_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// Wait used by the two single-thread roles: back off between polls so the spinning warp does not
// take issue slots from the softmax warps that share its scheduler.
__device__ __forceinline__ void a2_wait_backoff(uint64_t* bar, uint32_t parity) {
#ifdef ARB_HANG_GUARD
    uint32_t spins = 0;
#endif
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
#ifdef ARB_HANG_GUARD
        if (++spins > (1u << 24)) {
            printf("arb: attention2 mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x,
                   (void*)bar, parity);
            __trap();
        }
#endif
    }
}

struct A2Layout {  // byte offsets inside the 1024-aligned dynamic smem
    int kv_bytes;  // one buffer: K (2Kh rows) then V (2Kh rows), 128 B per row
    int q_off, bias_off, mask_off, exch_off, bar_off, total;
    int nbias;     // table entries; entry e <-> (j - i) = e - nqt*128
};
__host__ __device__ inline A2Layout a2_layout(int Kh, int nqt) {
    A2Layout L;
    L.kv_bytes = 4 * Kh * 128;
    L.q_off = 2 * L.kv_bytes;
    L.nbias = nqt * kA2QT + 2 * Kh;                 // even
    L.bias_off = L.q_off + kA2QT * 128;
    L.mask_off = L.bias_off + 2 * (L.nbias + 2) * 4;  // two copies (shift 0 / shift 1), padded
    L.exch_off = L.mask_off + 2 * (2 * Kh) * 4;       // one mask table per softmax group
    L.bar_off = (L.exch_off + 2 * kA2QT * 8 + 7) & ~7;  // exch: [parity][row] (m, l) of the publishing half
    L.total = L.bar_off + 32 * 8 + 16;
    return L;
}

// ---- pieces of the per-row software pipeline. A sub-block is ng <= 3 granules of 16 keys.
// TMEM -> registers (asynchronous; a2_wait_sub makes the registers usable)
// (NG = granules per sub-block is a template parameter: the pipeline body must be one basic block
// for the scheduler to interleave the two streams)
template <int NG>
__device__ __forceinline__ void a2_load_sub(uint32_t tS, uint32_t (&x)[48]) {
    tmem_ld_32x16_nc(tS, *reinterpret_cast<uint32_t(*)[16]>(&x[0]));
    if constexpr (NG > 1) tmem_ld_32x16_nc(tS + 16, *reinterpret_cast<uint32_t(*)[16]>(&x[16]));
    if constexpr (NG > 2) tmem_ld_32x16_nc(tS + 32, *reinterpret_cast<uint32_t(*)[16]>(&x[32]));
}
template <int NG>
__device__ __forceinline__ void a2_wait_sub(uint32_t (&x)[48]) {
    tmem_ld_wait_dep(*reinterpret_cast<uint32_t(*)[16]>(&x[0]));
    if constexpr (NG > 1) tmem_ld_wait_dep(*reinterpret_cast<uint32_t(*)[16]>(&x[16]));
    if constexpr (NG > 2) tmem_ld_wait_dep(*reinterpret_cast<uint32_t(*)[16]>(&x[32]));
}
// x <- s*scale + bias (+ mask); returns the row maximum over the sub-block
template <bool kMask, int NG>
__device__ __forceinline__ float a2_prep(uint32_t (&x)[48], float scale, const float2* __restrict__ pb2,
                                         const float* __restrict__ pm) {
    const float2 scale2 = make_float2(scale, scale);
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = q * 16 + 2 * j;
                float2 v = __ffma2_rn(make_float2(__uint_as_float(x[e]), __uint_as_float(x[e + 1])), scale2, pb2[e >> 1]);
                if (kMask) {
                    v.x += pm[e];
                    v.y += pm[e + 1];
                }
                x[e] = __float_as_uint(v.x);
                x[e + 1] = __float_as_uint(v.y);
                mx = fmaxf(mx, fmaxf(v.x, v.y));
            }
        }
    }
    return mx;
}
// one granule: p = exp2(x - m) -> 16-bit pairs over the S columns already in registers
template <bool kF16>
__device__ __forceinline__ void a2_exp_granule(uint32_t tS, int q, const uint32_t (&x)[48], float2 nm2, float2& l2) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int e = q * 16 + 2 * j;
        const float2 v = __fadd2_rn(make_float2(__uint_as_float(x[e]), __uint_as_float(x[e + 1])), nm2);
        const float2 p = make_float2(a2_exp2(v.x), a2_exp2(v.y));
        l2 = __fadd2_rn(l2, p);
        pk[j] = pack16x2<kF16>(p.x, p.y);
    }
    tmem_st_32x8_nc(tS + q * 8, pk);
}

template <bool kF16, int NG>
__global__ void __launch_bounds__(kA2Threads, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const float* __restrict__ rel_bias, int max_rel, const int32_t* __restrict__ mask,
                     h16* __restrict__ ctx, int B, int S, int heads, int Kh, float scale_log2e) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nqt = (S + kA2QT - 1) / kA2QT;
    constexpr int Kb = 16 * NG;  // keys per sub-block: 16, 32 or 48 (= Kh / kA2Sub)
    const A2Layout L = a2_layout(Kh, nqt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_empty = bars + 2;   // [2]
    uint64_t* q_full = bars + 4;
    uint64_t* q_empty = bars + 5;
    uint64_t* s_full = bars + 6;     // [half][sub-block]
    uint64_t* p_ready = bars + 14;   // [half][sub-block]
    uint64_t* o_prog = bars + 22;    // [half]: a P.V of sub-block 0..2 has retired (only the rescale path waits)
    uint64_t* o_full = bars + 24;    // [half]: P.V of the last sub-block has retired
    uint64_t* o_free = bars + 26;
    uint64_t* ml_ready = bars + 27;  // [2], by tile parity
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
    const int lane = threadIdx.x & 31;
    const int H = heads * 64;
    const int h = static_cast<int>(blockIdx.x) % heads;
    const int b_first = static_cast<int>(blockIdx.x) / heads;
    const int ngroups = static_cast<int>(gridDim.x) / heads;
    const int OFF = nqt * kA2QT;  // table entry e <-> (j - i) = e - OFF

    // ---- once per CTA: the head's bias table (x log2e), two copies shifted by one element
    float* T0 = reinterpret_cast<float*>(sm + L.bias_off);
    float* T1 = T0 + L.nbias + 2;
    for (int e = threadIdx.x; e < L.nbias + 2; e += kA2Threads) {
        const int rel = e - OFF;
        float v = 0.f;
        if (e < L.nbias && rel > -S && rel < S && rel_bias != nullptr)
            v = rel_bias[static_cast<int64_t>(h) * (2 * max_rel - 1) + rel + (max_rel - 1)] * kA2Log2e;
        T0[e] = v;
        if (e >= 1) T1[e - 1] = v;
    }
    if (threadIdx.x == 0) T1[L.nbias + 1] = 0.f;
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(kv_full + i, 1);
            mbar_init(kv_empty + i, 1);
            mbar_init(o_prog + i, 1);
            mbar_init(o_full + i, 1);
            mbar_init(ml_ready + i, 128);
        }
        for (int i = 0; i < 2 * kA2Sub; ++i) {
            mbar_init(s_full + i, 1);
            mbar_init(p_ready + i, 128);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        mbar_init(o_free, 128);
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

    // the two register regimes must be disjoint branches that only meet again at the teardown
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kA2RoleRegs));
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int n = 0, g = 0;
            for (int b = b_first; b < B; b += ngroups, ++n) {
                const int buf = n & 1;
                uint8_t* K = sm + buf * L.kv_bytes;
                uint8_t* V = K + 2 * Kh * 128;
                a2_wait_backoff(kv_empty + buf, ((n >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full + buf, L.kv_bytes);
                tma_load_3d(&tmap_kv, kv_full + buf, K, H + h * 64, 0, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, K + Kh * 128, H + h * 64, Kh, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, V, 2 * H + h * 64, 0, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, V + Kh * 128, 2 * H + h * 64, Kh, b, kEvictFirst);
                for (int t = 0; t < nqt; ++t, ++g) {
                    a2_wait_backoff(q_empty, (g & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full, kA2QT * 128);
                    tma_load_3d(&tmap_q, q_full, sm + L.q_off, h * 64, t * kA2QT, b, kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc_qk = umma_idesc_16bit(kA2QT, Kb, kF16);
            const uint32_t idesc_pv = umma_idesc_16bit_bmn(kA2QT, 64, kF16);
            const uint64_t dq = umma_desc_sw128(smem_u32(sm + L.q_off));
            const int my_items = B > b_first ? (B - 1 - b_first) / ngroups + 1 : 0;
            const int G = my_items * nqt;  // tiles this CTA processes, in order
            const uint32_t sm_base = smem_u32(sm);
            auto issue_qk = [&](int tile, int hh, int sb) {
                const int n = tile / nqt;
                const uint64_t dk = umma_desc_sw128(sm_base + (n & 1) * L.kv_bytes + (hh * Kh + sb * Kb) * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem + kA2ColS + hh * kA2MaxKh + sb * Kb, dq + 2 * k, dk + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                umma_commit(s_full + hh * kA2Sub + sb);
            };
            auto issue_pv = [&](int tile, int hh, int sb) {
                const int n = tile / nqt;
                const uint64_t dv = umma_desc_sw128(sm_base + (n & 1) * L.kv_bytes + 2 * Kh * 128 + (hh * Kh + sb * Kb) * 128);
                const uint32_t tP = tmem + kA2ColS + hh * kA2MaxKh + sb * Kb;
                for (int kk = 0; kk < Kb / 16; ++kk)
                    umma_bf16_ts(tmem + kA2ColO + hh * 64, tP + kk * 8, dv + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_pv,
                                 (sb | kk) != 0 ? 1u : 0u);
                umma_commit(sb < kA2Sub - 1 ? o_prog + hh : o_full + hh);
            };
            auto wait_inputs = [&](int tile) {  // K/V of the tile's item (first tile only) and its Q
                const int n = tile / nqt;
                if (tile % nqt == 0) mbar_wait(kv_full + (n & 1), (n >> 1) & 1);
                mbar_wait(q_full, tile & 1);
                tc_fence_after();
            };
            if (G > 0) {
                wait_inputs(0);
                for (int e = 0; e < 2 * kA2Sub; ++e) issue_qk(0, e & 1, e >> 1);
                umma_commit(q_empty);
            }
            for (int g = 0; g < G; ++g) {
                const uint32_t ph = g & 1;
                const bool has_next = g + 1 < G;
                // serve the sub-blocks in the order the groups finish them: (half 0, sb 0) (half 1, sb 0) (0, 1) ...
                // P.V of a sub-block, then at once the next tile's Q.K^T into the same TMEM columns
                // (tcgen05.mma executes in issue order, so the overwrite cannot pass the read).
                for (int e = 0; e < 2 * kA2Sub; ++e) {
                    const int hh = e & 1, sb = e >> 1;
                    mbar_wait(p_ready + hh * kA2Sub + sb, ph);
                    if (e == 0) mbar_wait(o_free, ph ^ 1);  // the merging group has read O of tile g-1
                    tc_fence_after();
                    issue_pv(g, hh, sb);
                    if (has_next) {
                        if (e == 0) wait_inputs(g + 1);
                        issue_qk(g + 1, hh, sb);
                        if (e == 2 * kA2Sub - 1) umma_commit(q_empty);
                    }
                }
                if (g % nqt == nqt - 1) umma_commit(kv_empty + ((g / nqt) & 1));  // item done: free its K/V buffer
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kA2SoftmaxRegs));
        // ===================== softmax groups =====================
        const int hh = (warp - 4) >> 2;             // key half owned by this group
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;          // query row inside the tile
        const int bar_id = 1 + hh;
        float* msk = reinterpret_cast<float*>(sm + L.mask_off) + hh * (2 * Kh);     // private mask table (0 / -inf)
        float2* exch = reinterpret_cast<float2*>(sm + L.exch_off);                  // [parity][row] (m, l) of the publishing half
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tSh = tmem + lane_sel + kA2ColS + hh * kA2MaxKh;
        const uint32_t tOh = tmem + lane_sel + kA2ColO + hh * 64;
        const int nu = nqt * kA2Sub;  // (tile, sub-block) units per item, even
        int g0 = 0;                   // tile counter of this CTA at the start of the item
        float m = -INFINITY, l = 0.f;

        // end of a tile: publish (m, l) of this half, or merge the halves and store the rows
        auto finish_tile = [&](int g, int b, int i) {
            const uint32_t ph = g & 1;
            const float m_pub = (m == -INFINITY) ? 0.f : m;
            if (hh == (g & 1)) {  // publisher of this tile: go on to the next one at once
                exch[ph * kA2QT + r] = make_float2(m_pub, l);
                mbar_arrive(ml_ready + ph);  // release: the smem write above is ordered before the arrive
                return;
            }
            mbar_wait(ml_ready + ph, (g >> 1) & 1);
            const float2 eo = exch[ph * kA2QT + r];  // (m, l) of the other half
            const float mt = fmaxf(m_pub, eo.x);
            const float a_mine = a2_exp2(m_pub - mt), a_other = a2_exp2(eo.x - mt);
            const float inv = __fdividef(1.f, a_mine * l + a_other * eo.y);
            const float wm = a_mine * inv, wo = a_other * inv;
            const float2 w0 = hh == 0 ? make_float2(wm, wm) : make_float2(wo, wo);  // weight of O[0]
            const float2 w1 = hh == 0 ? make_float2(wo, wo) : make_float2(wm, wm);  // weight of O[1]
            const uint32_t tO = tmem + lane_sel + kA2ColO;
            mbar_wait(o_full + 0, ph);
            mbar_wait(o_full + 1, ph);
            tc_fence_after();
            uint4* dst = reinterpret_cast<uint4*>(ctx + (static_cast<int64_t>(b) * S + i) * H + h * 64);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(tO + half * 32, o0);
                tmem_ld_32x32(tO + 64 + half * 32, o1);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    mbar_arrive(o_free);  // O is in registers: the next tile's P.V may overwrite it
                }
                if (i < S) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int d = q * 8 + e * 2;
                            const float2 x0 = make_float2(__uint_as_float(o0[d]), __uint_as_float(o0[d + 1]));
                            const float2 x1 = make_float2(__uint_as_float(o1[d]), __uint_as_float(o1[d + 1]));
                            const float2 y = __ffma2_rn(w0, x0, __fmul2_rn(w1, x1));
                            w[e] = pack16x2<kF16>(y.x, y.y);
                        }
                        dst[half * 4 + q] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        };

        // One pipeline step: exponentiate unit u (scores already prepared in `cur`), and in the issue
        // slots the MUFU leaves free read and prepare unit u + 1 into `nxt`.
        auto step = [&](auto mask_tag, int u, int b, uint32_t (&cur)[48], float mx_cur, uint32_t (&nxt)[48], float& mx_nxt) {
            constexpr bool kMask = decltype(mask_tag)::value;
            const int t = u >> 2, sb = u & 3;
            const int g = g0 + t;
            const int i = t * kA2QT + r;
            const uint32_t tS = tSh + sb * Kb;
            // online softmax across the sub-blocks of this half
            float alpha = 1.f;
            bool fix = false;
            if (sb == 0) {
                m = mx_cur;
                l = 0.f;
            } else if (mx_cur > m + kA2Tau) {  // also the case m == -inf (nothing unmasked so far)
                if (m != -INFINITY) {
                    alpha = a2_exp2(m - mx_cur);
                    l *= alpha;
                    fix = true;
                }
                m = mx_cur;
            }
            const float mm = (m == -INFINITY) ? 0.f : m;
            const float2 nm2 = make_float2(-mm, -mm);
            float2 l2 = make_float2(0.f, 0.f);
            // The unit after this one; the last unit of an item "prepares" itself again (values unused)
            // so that the body below stays free of branches.
            const int un = u + 1 < nu ? u + 1 : u;
            const int t2 = un >> 2, sb2 = un & 3;
            mbar_wait(s_full + hh * kA2Sub + sb2, (g0 + t2) & 1);
            tc_fence_after();
            const int key0n = hh * Kh + sb2 * Kb;
            const int startn = key0n - (t2 * kA2QT + r) + OFF;  // >= 1; bias of key column c is T0[start + c]
            const float2* pb2 = reinterpret_cast<const float2*>((startn & 1) ? T1 + (startn - 1) : T0 + startn);
            // ---- one basic block: loads of the next unit in flight during the first granule's exp2, then
            // its scale/bias/max arithmetic interleaved with the remaining granules
            a2_load_sub<NG>(tSh + sb2 * Kb, nxt);
            a2_exp_granule<kF16>(tS, 0, cur, nm2, l2);
            a2_wait_sub<NG>(nxt);
            if constexpr (NG > 1) a2_exp_granule<kF16>(tS, 1, cur, nm2, l2);
            if constexpr (NG > 2) a2_exp_granule<kF16>(tS, 2, cur, nm2, l2);
            mx_nxt = a2_prep<kMask, NG>(nxt, scale_log2e, pb2, msk + key0n);
            l += l2.x + l2.y;
            if (sb > 0 && __any_sync(0xffffffff, fix)) {
                // The running maximum moved by more than kTau: bring O_h (written with the old maximum)
                // to the new scale before this sub-block's P.V accumulates onto it.
                mbar_wait(o_prog + hh, static_cast<uint32_t>(3 * g + sb - 1) & 1u);
                tc_fence_after();
                const float a = fix ? alpha : 1.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t o[32];
                    tmem_ld_32x32(tOh + half * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * a);
                    tmem_st_32x16(tOh + half * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
                    tmem_st_32x16(tOh + half * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_ready + hh * kA2Sub + sb);
            if (sb == kA2Sub - 1) finish_tile(g, b, i);
        };

        for (int b = b_first; b < B; b += ngroups, g0 += nqt) {
            // ---- per-item mask table (private to the group: no cross-group synchronisation)
            a2_bar_sync(bar_id, 128);  // everyone is done with the previous item's table
            bool mine_clear = true, mine_on = false;
            for (int j = r; j < 2 * Kh; j += 128) {
                const bool on = j < S && mask[static_cast<int64_t>(b) * S + j] != 0;
                msk[j] = on ? 0.f : -INFINITY;
                mine_clear &= on;
                mine_on |= on;
            }
            const bool clear = a2_bar_red_and(bar_id, 128, mine_clear);  // no masked / out-of-range key at all
            const bool any_on = a2_bar_red_or(bar_id, 128, mine_on);
            if (!any_on) {
                // Every key masked: the reference adds finfo.min to all scores, which absorbs them in
                // fp32 -> uniform attention over the S keys: p = 1 for keys < S, 0 beyond.
                for (int u = 0; u < nu; ++u) {
                    const int t = u >> 2, sb = u & 3, g = g0 + t;
                    const int key0 = hh * Kh + sb * Kb;
                    if (sb == 0) {
                        m = 0.f;
                        l = 0.f;
                    }
                    mbar_wait(s_full + hh * kA2Sub + sb, g & 1);
                    tc_fence_after();
                    for (int q = 0; q < NG; ++q) {
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const int key = key0 + q * 16 + j;
                            const float p0 = key < S ? 1.f : 0.f, p1 = key + 1 < S ? 1.f : 0.f;
                            l += p0 + p1;
                            pk[j >> 1] = pack16x2<kF16>(p0, p1);
                        }
                        tmem_st_32x8_nc(tSh + sb * Kb + q * 8, pk);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(p_ready + hh * kA2Sub + sb);
                    if (sb == kA2Sub - 1) finish_tile(g, b, t * kA2QT + r);
                }
                continue;
            }
            // ---- pipeline prologue: read and prepare unit 0, then two steps per iteration (the two
            // register buffers swap roles without any copy)
            uint32_t xa[48], xb[48];
            float mxa = -INFINITY, mxb = -INFINITY;
            {
                mbar_wait(s_full + hh * kA2Sub + 0, g0 & 1);
                tc_fence_after();
                a2_load_sub<NG>(tSh, xa);
                a2_wait_sub<NG>(xa);
                const int key0 = hh * Kh;
                const int start = key0 - r + OFF;
                const float2* pb2 = reinterpret_cast<const float2*>((start & 1) ? T1 + (start - 1) : T0 + start);
                mxa = clear ? a2_prep<false, NG>(xa, scale_log2e, pb2, msk + key0)
                            : a2_prep<true, NG>(xa, scale_log2e, pb2, msk + key0);
            }
            if (clear) {
#pragma unroll 1
                for (int u = 0; u < nu; u += 2) {
                    step(std::false_type{}, u, b, xa, mxa, xb, mxb);
                    step(std::false_type{}, u + 1, b, xb, mxb, xa, mxa);
                }
            } else {
#pragma unroll 1
                for (int u = 0; u < nu; u += 2) {
                    step(std::true_type{}, u, b, xa, mxa, xb, mxb);
                    step(std::true_type{}, u + 1, b, xb, mxb, xa, mxa);
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

bool attention_tc2_supported(int S, int dh) { return dh == 64 && S >= 1 && S <= 2 * kA2MaxKh; }

int launch_attention_tc2(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                         h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream) {
    ARB_REQUIRE(qkv && mask && ctx, "attention_tc2: null pointer");
    ARB_REQUIRE(attention_tc2_supported(S, dh), "attention_tc2: S=%d dh=%d unsupported", S, dh);
    ARB_REQUIRE(B > 0 && (rel_bias == nullptr || S <= max_rel), "attention_tc2: bad shape B=%d S=%d max_rel=%d", B, S, max_rel);
    const int H = heads * dh;
    const int Kh = ((S + 1) / 2 + 63) / 64 * 64;  // four sub-blocks of whole 16-key granules per half
    const int nqt = (S + kA2QT - 1) / kA2QT;
    const A2Layout L = a2_layout(Kh, nqt);
    const int smem = L.total + 1024;
    ARB_REQUIRE(smem <= 232448, "attention_tc2: shared memory %d exceeds 227 KB", smem);
    CUtensorMap tq, tkv;
    if (!make_tmap_bf16_batched_k64(&tq, qkv, B, S, 3 * H, 3 * H, kA2QT) ||
        !make_tmap_bf16_batched_k64(&tkv, qkv, B, S, 3 * H, 3 * H, Kh)) {
        set_error("attention_tc2: cuTensorMapEncodeTiled failed");
        return ARB_ERR_CUDA;
    }
    const int ng = Kh / kA2Sub / 16;  // 16-key granules per sub-block
    auto kern = ng == 3 ? (fp16 ? attention_tc2_kernel<true, 3> : attention_tc2_kernel<false, 3>)
              : ng == 2 ? (fp16 ? attention_tc2_kernel<true, 2> : attention_tc2_kernel<false, 2>)
                        : (fp16 ? attention_tc2_kernel<true, 1> : attention_tc2_kernel<false, 1>);
    ARB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int ngroups = num_sms() / heads;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > B) ngroups = B;
    const int grid = ngroups * heads;
    const float scale_log2e = kA2Log2e / sqrtf(static_cast<float>(dh));
    kern<<<grid, kA2Threads, smem, stream>>>(tq, tkv, rel_bias, max_rel, mask, ctx, B, S, heads, Kh, scale_log2e);
    ARB_CHECK_CUDA(cudaGetLastError());
    return ARB_OK;
}

}  // namespace arb
