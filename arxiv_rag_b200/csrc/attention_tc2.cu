// tcgen05/TMEM self-attention for the MPNet encoder, second schedule (S <= 384, head dim 64):
//   softmax(q.k^T/8 + rel_bias[h][j-i] + mask) . v   per (sequence, head)
// Same contract as attention_tc.cu / attention_mma.cu (modeling_mpnet.py:162-177, :324-360; mask of
// modeling_utils.py:936-947).
//
// What changed against attention_tc.cu, and why. There, a softmax group (4 warps, one thread per
// query row) owns one half of the keys of every 128-query tile; after writing its probabilities it
// waits for P.V and the next tile's Q.K^T of that half (the turnaround, ~1/3 of a tile's time,
// during which its MUFU/FMA pipes idle: the group's own chain is latency bound, so staggering the
// two groups buys nothing). Here each half is cut into two SUB-BLOCKS of Kb = Kh/2 keys that are
// scored, soft-maxed and multiplied independently:
//     S_{h,sb} = Q . K_{h,sb}^T   (M = 128, N = Kb)            -> TMEM S[h][sb]
//     O_h (+)= P_{h,sb} . V_{h,sb}                              -> TMEM O[h]
// While a group works on sub-block b of tile t, the tensor core turns sub-block a around
// (P.V of tile t, then Q.K^T of tile t+1 into the same TMEM columns): when the group is done with
// b, its next score sub-block is already waiting. Same TMEM (S 2x192 + O 2x64 = 512 columns) and
// the same shared memory as before.
// Within a half the two sub-blocks share one accumulator, so the softmax is online across them:
// running maximum m and row sum l; sub-block b keeps m unless its maximum exceeds it by more than
// kTau log2 units (then O_h and l are rescaled by the group itself, a rare path). The maximum is
// the exact one of s*scale + bias + mask — every score of a sub-block sits in registers between
// the single TMEM read and the exp2 — so probabilities never exceed 2^kTau and fp16 mode needs no
// shift tricks. The two halves are merged as before:
//     O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),  a_h = exp2(m_h - max(m_0, m_1)).
// The scale/bias/shift arithmetic is packed (fma.rn.f32x2 / add.rn.f32x2).
// STATUS: an experiment, compiled only with -DARB_WITH_ATTENTION_TC2 (python -m arxiv_rag_b200.build
// --variant=tc2 -DARB_WITH_ATTENTION_TC2). Parity-green on every shape of the attention tests
// (also S < 64 and the rescale path), but measured SLOWER than attention_tc.cu: 1.23 vs 1.11 ms at
// B 1024 x S 384 (0.66 vs 0.59 at S 256; 1.01 vs 0.57 at B 4096 x S 64), and a finer variant with
// four 48-key sub-blocks per half and two register buffers (prepare sub-block k+1 inside the exp2
// phase of sub-block k; git history) slower again at 1.46 ms: every extra sub-block adds a
// p_ready -> P.V -> Q.K^T -> s_full round trip through the single MMA-issuing thread, and that
// costs more than the turnaround it hides (profiles/r2_attention_notes.md).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "ptx.cuh"

#ifdef ARB_WITH_ATTENTION_TC2
namespace arb {

// Three warpgroups: warps 0-3 = roles (warp 0 TMA, warp 1 MMA, warps 2-3 idle), warps 4-7 softmax half
// 0, warps 8-11 softmax half 1 + combine/store. The softmax threads keep a whole 96-key score
// sub-block in registers, more than the 168 a 384-thread CTA gets evenly: the role warpgroup hands
// its registers over with setmaxnreg (56 for the roles, 224 for the softmax warpgroups).
constexpr int kA2Threads = 384;
constexpr int kA2RoleRegs = 72, kA2SoftmaxRegs = 208;
constexpr int kA2QT = 128;            // query rows per tile
constexpr int kA2MaxKh = 192;         // keys per half
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
    L.total = L.bar_off + 24 * 8 + 16;
    return L;
}

// One sub-block (nch <= 3 chunks of 32 keys) of one query row: S -> P in place; online (m, l).
// Returns the factor by which the half's accumulator O_h and the previous l were to be scaled
// (1 unless the running maximum had to move), with `fix` telling whether O_h needs the rescale.
template <bool kF16, bool kMask>
__device__ __forceinline__ void softmax_sub(uint32_t tS, int nch, float scale, const float2* __restrict__ pb2,
                                            const float* __restrict__ pm, bool first, float& m, float& l,
                                            float& alpha, bool& fix) {
    uint32_t xr[96];
    // every score of the sub-block in one go: three loads in flight, one wait
    tmem_ld_32x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&xr[0]));
    if (nch > 1) tmem_ld_32x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&xr[32]));
    if (nch > 2) tmem_ld_32x32(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&xr[64]));
    tmem_ld_wait();
    const float2 scale2 = make_float2(scale, scale);
    float mx = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        if (ch < nch) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int e = ch * 32 + 2 * j;
                float2 v = __ffma2_rn(make_float2(__uint_as_float(xr[e]), __uint_as_float(xr[e + 1])), scale2, pb2[e >> 1]);
                if (kMask) {
                    v.x += pm[e];
                    v.y += pm[e + 1];
                }
                xr[e] = __float_as_uint(v.x);
                xr[e + 1] = __float_as_uint(v.y);
                mx = fmaxf(mx, fmaxf(v.x, v.y));
            }
        }
    }
    // online softmax across the sub-blocks of this half
    alpha = 1.f;
    fix = false;
    if (first) {
        m = mx;
    } else if (mx > m + kA2Tau) {  // also the case m == -inf (nothing unmasked so far)
        if (m != -INFINITY) {
            alpha = a2_exp2(m - mx);
            l *= alpha;
            fix = true;
        }
        m = mx;
    }
    const float mm = (m == -INFINITY) ? 0.f : m;
    const float2 nm2 = make_float2(-mm, -mm);
    float2 l2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        if (ch < nch) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int e = ch * 32 + 2 * j;
                const float2 v = __fadd2_rn(make_float2(__uint_as_float(xr[e]), __uint_as_float(xr[e + 1])), nm2);
                const float2 p = make_float2(a2_exp2(v.x), a2_exp2(v.y));
                l2 = __fadd2_rn(l2, p);
                pk[j] = pack16x2<kF16>(p.x, p.y);
            }
            tmem_st_32x16(tS + ch * 16, pk);  // P (16-bit pairs) over the S columns already in registers
        }
    }
    l += l2.x + l2.y;
}

template <bool kF16>
__global__ void __launch_bounds__(kA2Threads, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const float* __restrict__ rel_bias, int max_rel, const int32_t* __restrict__ mask,
                     h16* __restrict__ ctx, int B, int S, int heads, int Kh, float scale_log2e) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nqt = (S + kA2QT - 1) / kA2QT;
    const int Kb = Kh / 2;
    const A2Layout L = a2_layout(Kh, nqt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_empty = bars + 2;   // [2]
    uint64_t* q_full = bars + 4;
    uint64_t* q_empty = bars + 5;
    uint64_t* s_full = bars + 6;     // [half][sub-block]
    uint64_t* p_ready = bars + 10;   // [half][sub-block]
    uint64_t* o_part = bars + 14;    // [half]: P.V of sub-block a has retired (only the rescale path waits)
    uint64_t* o_full = bars + 16;    // [half]: P.V of sub-block b has retired
    uint64_t* o_free = bars + 18;
    uint64_t* ml_ready = bars + 19;  // [2], by tile parity
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

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
            mbar_init(o_part + i, 1);
            mbar_init(o_full + i, 1);
            mbar_init(ml_ready + i, 128);
        }
        for (int i = 0; i < 4; ++i) {
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
                umma_commit(s_full + hh * 2 + sb);
            };
            auto issue_pv = [&](int tile, int hh, int sb) {
                const int n = tile / nqt;
                const uint64_t dv = umma_desc_sw128(sm_base + (n & 1) * L.kv_bytes + 2 * Kh * 128 + (hh * Kh + sb * Kb) * 128);
                const uint32_t tP = tmem + kA2ColS + hh * kA2MaxKh + sb * Kb;
                for (int kk = 0; kk < Kb / 16; ++kk)
                    umma_bf16_ts(tmem + kA2ColO + hh * 64, tP + kk * 8, dv + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_pv,
                                 (sb | kk) != 0 ? 1u : 0u);
                umma_commit(sb == 0 ? o_part + hh : o_full + hh);
            };
            auto wait_inputs = [&](int tile) {  // K/V of the tile's item (first tile only) and its Q
                const int n = tile / nqt;
                if (tile % nqt == 0) mbar_wait(kv_full + (n & 1), (n >> 1) & 1);
                mbar_wait(q_full, tile & 1);
                tc_fence_after();
            };
            if (G > 0) {
                wait_inputs(0);
                for (int e = 0; e < 4; ++e) issue_qk(0, e & 1, e >> 1);
                umma_commit(q_empty);
            }
            for (int g = 0; g < G; ++g) {
                const uint32_t ph = g & 1;
                const bool has_next = g + 1 < G;
                // serve the sub-blocks in the order the groups finish them: (half 0, a) (half 1, a) (0, b) (1, b).
                // P.V of a sub-block, then at once the next tile's Q.K^T into the same TMEM columns
                // (tcgen05.mma executes in issue order, so the overwrite cannot pass the read).
                for (int e = 0; e < 4; ++e) {
                    const int hh = e & 1, sb = e >> 1;
                    mbar_wait(p_ready + hh * 2 + sb, ph);
                    if (e == 0) mbar_wait(o_free, ph ^ 1);  // the combining group has read O of tile g-1
                    tc_fence_after();
                    issue_pv(g, hh, sb);
                    if (has_next) {
                        if (e == 0) wait_inputs(g + 1);
                        issue_qk(g + 1, hh, sb);
                        if (e == 3) umma_commit(q_empty);
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
        float2* exch = reinterpret_cast<float2*>(sm + L.exch_off);                  // [parity][row] (m, l) of half 0
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tSh = tmem + lane_sel + kA2ColS + hh * kA2MaxKh;
        const uint32_t tOh = tmem + lane_sel + kA2ColO + hh * 64;
        const int nch = Kb / 32;
        int g = 0;
        for (int b = b_first; b < B; b += ngroups) {
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
            for (int t = 0; t < nqt; ++t, ++g) {
                const uint32_t ph = g & 1;
                const int i = t * kA2QT + r;
                float m = -INFINITY, l = 0.f;
#pragma unroll 1
                for (int sb = 0; sb < 2; ++sb) {
                    const int key0 = hh * Kh + sb * Kb;
                    const int start = key0 - i + OFF;  // >= 1; bias of key column c is T0[start + c]
                    const float2* pb2 = reinterpret_cast<const float2*>((start & 1) ? T1 + (start - 1) : T0 + start);
                    const uint32_t tS = tSh + sb * Kb;
                    mbar_wait(s_full + hh * 2 + sb, ph);
                    tc_fence_after();
                    float alpha = 1.f;
                    bool fix = false;
                    if (!any_on) {
                        // Every key masked: the reference adds finfo.min to all scores, which absorbs them
                        // in fp32 -> uniform attention over the S keys: p = 1 for keys < S, 0 beyond.
                        m = 0.f;
                        for (int cc = 0; cc < nch; ++cc) {
                            uint32_t pk[16];
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const int key = key0 + cc * 32 + j;
                                const float p0 = key < S ? 1.f : 0.f, p1 = key + 1 < S ? 1.f : 0.f;
                                l += p0 + p1;
                                pk[j >> 1] = pack16x2<kF16>(p0, p1);
                            }
                            tmem_st_32x16(tS + cc * 16, pk);
                        }
                    } else if (clear) {
                        softmax_sub<kF16, false>(tS, nch, scale_log2e, pb2, msk + key0, sb == 0, m, l, alpha, fix);
                    } else {
                        softmax_sub<kF16, true>(tS, nch, scale_log2e, pb2, msk + key0, sb == 0, m, l, alpha, fix);
                    }
                    if (sb == 1 && __any_sync(0xffffffff, fix)) {
                        // The running maximum moved by more than kTau: bring O_h (= P_a . V_a, written with the
                        // old maximum) to the new scale before P_b . V_b accumulates onto it.
                        mbar_wait(o_part + hh, ph);
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
                    mbar_arrive(p_ready + hh * 2 + sb);
                }
                const float m_pub = (m == -INFINITY) ? 0.f : m;
                if (hh == 0) {
                    // publish (m, l) of half 0 and go on to the next tile; half 1's group combines
                    exch[ph * kA2QT + r] = make_float2(m_pub, l);
                    mbar_arrive(ml_ready + ph);  // release: the smem write above is ordered before the arrive
                    continue;
                }
                // ---- group 1: merge the halves, normalise, store this row of ctx
                mbar_wait(ml_ready + ph, (g >> 1) & 1);
                const float2 e0 = exch[ph * kA2QT + r];
                const float mt = fmaxf(m_pub, e0.x);
                const float a0 = a2_exp2(e0.x - mt), a1 = a2_exp2(m_pub - mt);
                const float inv = __fdividef(1.f, a0 * e0.y + a1 * l);
                const float2 w0 = make_float2(a0 * inv, a0 * inv), w1 = make_float2(a1 * inv, a1 * inv);
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
    const int Kh = ((S + 1) / 2 + 63) / 64 * 64;  // two sub-blocks of whole 32-column TMEM chunks per half
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
    auto kern = fp16 ? attention_tc2_kernel<true> : attention_tc2_kernel<false>;
    ARB_CHECK_CUDA(set_max_smem_once(kern, smem));
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
#else
#include "kernels.h"
namespace arb {
bool attention_tc2_supported(int, int) { return false; }
int launch_attention_tc2(const h16*, const float*, int, const int32_t*, h16*, int, int, int, int, bool, cudaStream_t) {
    set_error("attention impl 3 (attention_tc2.cu) is not part of this build (-DARB_WITH_ATTENTION_TC2)");
    return ARB_ERR_UNSUPPORTED;
}
}  // namespace arb
#endif
