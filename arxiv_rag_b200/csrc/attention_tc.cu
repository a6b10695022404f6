// tcgen05/TMEM self-attention for the MPNet encoder (64 <= S <= 384, head dim 64):
//   softmax(q.k^T/8 + rel_bias[h][j-i] + mask) . v   per (sequence, head)
// Same contract as attention_mma.cu (modeling_mpnet.py:162-177, :324-360; mask of
// modeling_utils.py:936-947); this is the tensor-core version the encode path uses.
//
// CTA c serves head c % heads for sequences c / heads, + ngroups, ... (one persistent CTA per SM;
// the `heads` CTAs of a group read the same sequence's rows at the same time). The key range is
// cut in two halves of Kh keys; per 128-query tile the MMA warp issues
//     S_h = Q . K_h^T          (tcgen05.mma kind::f16, M=128, N=Kh, 4 K-steps)  -> TMEM S[h]
//     O_h = P_h . V_h          (A = P_h read from TMEM, B = V_h MN-major in smem) -> TMEM O[h]
// and two softmax groups (4 warps each, one thread per query row) turn S_h into P_h in place:
// a max pass over the raw scores, then exp2(s*scale + bias - c) packed to 16-bit and stored back
// over the columns already consumed (P aliases S). Each half keeps its own shift c_h and row sum
// l_h, so no accumulator is ever rescaled; a third group combines
//     O = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1),  a_h = exp2(c_h - max(c_0, c_1))
// and stores the row. The MMA warp interleaves PV_h(t) with QK_h(t+1) so a group's next score
// tile is produced as soon as its P has been consumed.
// The relative-position bias depends only on j-i: the head's table (x log2e) is staged once per
// CTA in two copies shifted by one element, so every thread reads its contiguous run with aligned
// 8-byte loads.
// TMEM: S[0] cols 0..191, S[1] 192..383, O[0] 384..447, O[1] 448..511 (all 512 columns).
// Smem: K and V of the head double-buffered across items (TMA, 128-byte swizzle), one Q tile.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "ptx.cuh"

namespace arb {

// warp 0 TMA, warp 1 MMA, warps 2-5 softmax half 0 + combine/store, 6-9 softmax half 1
constexpr int kAtThreads = 320;
constexpr int kAtQT = 128;            // query rows per tile
constexpr int kAtMaxKh = 192;
constexpr uint32_t kAtColS = 0, kAtColO = 384;
constexpr float kAtLog2e = 1.4426950408889634f;

__device__ __forceinline__ float at_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// barrier + AND / OR reduction of a predicate over the `n` threads of named barrier `id`
__device__ __forceinline__ bool bar_red_and(int id, int n, bool p) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(p)), "r"(id), "r"(n)
        : "memory");
    return out != 0;
}
__device__ __forceinline__ bool bar_red_or(int id, int n, bool p) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(static_cast<uint32_t>(p)), "r"(id), "r"(n)
        : "memory");
    return out != 0;
}
__device__ __forceinline__ void bar_sync_n(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// Wait used by the two single-thread roles: back off between polls so the spinning warp does not
// take issue slots from the softmax warps that share its scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
#ifdef ARB_HANG_GUARD
    uint32_t spins = 0;
#endif
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
#ifdef ARB_HANG_GUARD
        if (++spins > (1u << 24)) {
            printf("arb: attention mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x,
                   (void*)bar, parity);
            __trap();
        }
#endif
    }
}

struct AtLayout {  // byte offsets inside the 1024-aligned dynamic smem
    int kv_bytes;  // one buffer: K (2Kh rows) then V (2Kh rows), 128 B per row
    int q_off, bias_off, mask_off, exch_off, red_off, bar_off, total;
    int nbias;     // table entries; entry e <-> (j - i) = e - nqt*128
};
__host__ __device__ inline AtLayout at_layout(int Kh, int nqt) {
    AtLayout L;
    L.kv_bytes = 4 * Kh * 128;
    L.q_off = 2 * L.kv_bytes;
    L.nbias = nqt * kAtQT + 2 * Kh;                 // even
    L.bias_off = L.q_off + kAtQT * 128;
    L.mask_off = L.bias_off + 2 * (L.nbias + 2) * 4;  // two copies (shift 0 / shift 1), padded
    L.exch_off = L.mask_off;                          // (no per-key mask table: chunk flags + key bitmasks after `red`)
    L.red_off = L.exch_off + 2 * 2 * kAtQT * 8;       // [parity][half][row] (c, l)
    L.bar_off = (L.red_off + 32 * 4 + 16 + 64 + 7) & ~7;  // + mask flags [half][8] (bytes) and key bitmasks [half][8]
    L.total = L.bar_off + 16 * 8 + 16;
    return L;
}

// One 128-row x Kh-column score tile -> P (in place), returns the shift c and the row sum l.
// va/vb double-buffer the TMEM loads so a chunk's latency hides behind the previous chunk's math.
template <bool kF16, bool kMask>
__device__ __forceinline__ void softmax_tile(uint32_t tS, int nchunk, float scale, float bmax,
                                             const float2* __restrict__ pb2, const float* __restrict__ pm,
                                             float& c_out, float& l_out) {
    uint32_t va[32], vb[32];
    float mraw = -INFINITY;
    auto max32 = [&](const uint32_t (&v)[32], int c0) {
        if (!kMask) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(v[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(v[j]) + pm[c0 + j]);
        }
    };
    // ---- pass 1: max of the raw scores over the unmasked keys
    tmem_ld_32x32(tS, va);
    for (int cc = 0; cc < nchunk; cc += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(tS + (cc + 1 < nchunk ? (cc + 1) * 32 : 0), vb);  // next chunk, or chunk 0 for pass 2
        max32(va, cc * 32);
        if (cc + 1 < nchunk) {
            tmem_ld_wait();
            tmem_ld_32x32(tS + (cc + 2 < nchunk ? (cc + 2) * 32 : 0), va);
            max32(vb, (cc + 1) * 32);
        }
    }
    // chunk 0 is in flight again: in va if nchunk is even, in vb if odd.
    // Shift c >= row max of (s*scale + bias + mask): the exact max is not needed, only a bound
    // within a few units so exp2 stays in range.
    const float c = (mraw == -INFINITY) ? 0.f : fmaf(mraw, scale, bmax);
    // ---- pass 2: p = exp2(s*scale + bias + mask - c), packed 16-bit, stored over S (in place)
    float l = 0.f;
    auto emit = [&](const uint32_t (&v)[32], int c0) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float2 bb = pb2[(c0 + j) >> 1];
            float x0 = fmaf(__uint_as_float(v[j]), scale, bb.x) - c;
            float x1 = fmaf(__uint_as_float(v[j + 1]), scale, bb.y) - c;
            if (kMask) {
                x0 += pm[c0 + j];
                x1 += pm[c0 + j + 1];
            }
            const float p0 = at_exp2(x0), p1 = at_exp2(x1);
            l += p0 + p1;
            pk[j >> 1] = pack16x2<kF16>(p0, p1);
        }
        tmem_st_32x16(tS + (c0 >> 1), pk);
    };
    const bool first_in_a = (nchunk & 1) == 0;
    for (int cc = 0; cc < nchunk; ++cc) {
        const bool in_a = first_in_a == ((cc & 1) == 0);
        tmem_ld_wait();
        if (in_a) {
            if (cc + 1 < nchunk) tmem_ld_32x32(tS + (cc + 1) * 32, vb);
            emit(va, cc * 32);
        } else {
            if (cc + 1 < nchunk) tmem_ld_32x32(tS + (cc + 1) * 32, va);
            emit(vb, cc * 32);
        }
    }
    c_out = c;
    l_out = l;
}

// Masked keys (padding inside a batch, keys beyond S). Per 32-key chunk of a half: flag 0 = every key
// valid, 1 = some, 2 = none, plus the chunk's key bitmask (one ballot while the item is set up).
// A sequence with masked keys runs the SAME softmax code as a fully valid one: a pre-pass rewrites
// the scores of the (few) partly masked chunks to -inf in TMEM, the main passes stop after the last
// chunk that has a valid key, and a post-pass zeroes P of the chunks behind it. A length-sorted batch
// of real chunks masks a few trailing keys of almost every row; the former per-key mask reads on whole
// rows cost 25-30 % of the kernel, and keeping a masked variant of the passes inline cost the
// unmasked path 1-8 % through the kernel-wide register allocation.
__device__ __forceinline__ void mask_prepass(uint32_t tS, int n_eff, const uint32_t* __restrict__ cbits,
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
template <bool kF16>
__device__ __forceinline__ void mask_postpass(uint32_t tS, int n_eff, int nchunk) {
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    for (int cc = n_eff; cc < nchunk; ++cc) tmem_st_32x16(tS + cc * 16, z);
}

// kDefer: which softmax group combines the two key halves and stores the tile.
//   false  group 0 does, right after its own softmax — it then waits for group 1's half of the same
//          tile, which keeps the two groups in lockstep (both in the TMEM-read/max pass together,
//          both in the MUFU pass together, both idle during the PV -> QK turnaround);
//   true   group 0 only publishes its (shift, row sum) and moves on to the next tile; group 1
//          combines after its own softmax, inside the wait for its next score tile. The MMA warp
//          serves the groups alternately, so they settle into a stagger: one group's exp2 pass
//          overlaps the other's turnaround and max pass.
template <bool kF16, bool kDefer>
__global__ void __launch_bounds__(kAtThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const float* __restrict__ rel_bias, int max_rel, const int32_t* __restrict__ mask,
                    h16* __restrict__ ctx, int B, int S, int heads, int Kh, float scale_log2e) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nqt = (S + kAtQT - 1) / kAtQT;
    const AtLayout L = at_layout(Kh, nqt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar_off);
    uint64_t* kv_full = bars + 0;    // [2]
    uint64_t* kv_empty = bars + 2;   // [2]
    uint64_t* q_full = bars + 4;
    uint64_t* q_empty = bars + 5;
    uint64_t* s_full = bars + 6;     // [2]
    uint64_t* p_ready = bars + 8;    // [2]
    uint64_t* o_full = bars + 10;    // [2]
    uint64_t* o_free = bars + 12;
    uint64_t* ml_ready = bars + 13;  // [2], by tile parity: a slot is reused two tiles later, after its combine
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
    const int lane = threadIdx.x & 31;
    const int H = heads * 64;
    const int h = static_cast<int>(blockIdx.x) % heads;
    const int b_first = static_cast<int>(blockIdx.x) / heads;
    const int ngroups = static_cast<int>(gridDim.x) / heads;
    const int OFF = nqt * kAtQT;  // table entry e <-> (j - i) = e - OFF

    // ---- once per CTA: the head's bias table (two copies, shift 0 and 1) and its maximum
    float* T0 = reinterpret_cast<float*>(sm + L.bias_off);
    float* T1 = T0 + L.nbias + 2;
    float* red = reinterpret_cast<float*>(sm + L.red_off);
    {
        float bm = -INFINITY, bn = INFINITY;
        for (int e = threadIdx.x; e < L.nbias + 2; e += kAtThreads) {
            const int rel = e - OFF;
            float v = 0.f;
            if (e < L.nbias && rel > -S && rel < S) {
                v = rel_bias[static_cast<int64_t>(h) * (2 * max_rel - 1) + rel + (max_rel - 1)] * kAtLog2e;
                bm = fmaxf(bm, v);
                bn = fminf(bn, v);
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
            red[16 + warp] = bn;
        }
    }
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(kv_full + i, 1);
            mbar_init(kv_empty + i, 1);
            mbar_init(s_full + i, 1);
            mbar_init(p_ready + i, 128);
            mbar_init(o_full + i, 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        mbar_init(o_free, 128);
        mbar_init(ml_ready + 0, 128);
        mbar_init(ml_ready + 1, 128);
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
    // the bias table, barriers and TMEM above depend on nothing an earlier kernel wrote; qkv and mask do
    pdl_launch_dependents();
    pdl_wait();
    float bmax = red[0], bmin = red[16];
#pragma unroll
    for (int w = 1; w < kAtThreads / 32; ++w) {
        bmax = fmaxf(bmax, red[w]);
        bmin = fminf(bmin, red[16 + w]);
    }
    // The softmax shift is the bound c = max(raw) * scale + max(bias) >= the true row maximum; the
    // true maximum sits at most D = max(bias) - min(bias) below it. bf16 probabilities keep the fp32
    // exponent range, but fp16 ones would drift into subnormals once D exceeds 14 (a trained
    // relative-position table spreads over many log2 units). fp16 mode therefore lowers the shift
    // by min(D, 15): p <= 2^15 still fits fp16, the row maximum stays >= 2^-14 for D up to 29, and
    // both key halves use the same constant, so the combine is unchanged.
    if (kF16) bmax -= fminf(bmax - bmin, 15.f);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int n = 0, g = 0;
            for (int b = b_first; b < B; b += ngroups, ++n) {
                const int buf = n & 1;
                uint8_t* K = sm + buf * L.kv_bytes;
                uint8_t* V = K + 2 * Kh * 128;
                mbar_wait_backoff(kv_empty + buf, ((n >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(kv_full + buf, L.kv_bytes);
                tma_load_3d(&tmap_kv, kv_full + buf, K, H + h * 64, 0, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, K + Kh * 128, H + h * 64, Kh, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, V, 2 * H + h * 64, 0, b, kEvictFirst);
                tma_load_3d(&tmap_kv, kv_full + buf, V + Kh * 128, 2 * H + h * 64, Kh, b, kEvictFirst);
                for (int t = 0; t < nqt; ++t, ++g) {
                    mbar_wait_backoff(q_empty, (g & 1) ^ 1);
                    mbar_arrive_expect_tx(q_full, kAtQT * 128);
                    tma_load_3d(&tmap_q, q_full, sm + L.q_off, h * 64, t * kAtQT, b, kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc_qk = umma_idesc_16bit(kAtQT, Kh, kF16);
            const uint32_t idesc_pv = umma_idesc_16bit_bmn(kAtQT, 64, kF16);
            const uint64_t dq = umma_desc_sw128(smem_u32(sm + L.q_off));
            const int my_items = B > b_first ? (B - 1 - b_first) / ngroups + 1 : 0;
            const int G = my_items * nqt;  // tiles this CTA processes, in order
            const uint32_t sm_base = smem_u32(sm);
            // S_h(tile) = Q . K_h^T ; the tile's item selects the K/V buffer
            auto issue_qk = [&](int tile, int hh) {
                const int n = tile / nqt;
                const uint32_t k_addr = sm_base + (n & 1) * L.kv_bytes;
                const uint64_t dk = umma_desc_sw128(k_addr + hh * Kh * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem + kAtColS + hh * kAtMaxKh, dq + 2 * k, dk + 2 * k, idesc_qk, k > 0 ? 1u : 0u);
                umma_commit(s_full + hh);
            };
            auto issue_pv = [&](int tile, int hh) {
                const int n = tile / nqt;
                const uint32_t v_addr = sm_base + (n & 1) * L.kv_bytes + 2 * Kh * 128;
                const uint64_t dv = umma_desc_sw128(v_addr + hh * Kh * 128);
                for (int kk = 0; kk < Kh / 16; ++kk)
                    umma_bf16_ts(tmem + kAtColO + hh * 64, tmem + kAtColS + hh * kAtMaxKh + kk * 8,
                                 dv + static_cast<uint64_t>(kk) * (2048 >> 4), idesc_pv, kk > 0 ? 1u : 0u);
                umma_commit(o_full + hh);
            };
            auto wait_inputs = [&](int tile) {  // K/V of the tile's item (first tile only) and its Q
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
                // PV_0(g), then immediately the next tile's QK_0 (its S aliases the P just consumed;
                // tcgen05.mma executes in issue order, so the overwrite cannot pass the read)
                mbar_wait(p_ready + 0, ph);
                mbar_wait(o_free, ph ^ 1);  // the combine group has read O of tile g-1
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
                if (g % nqt == nqt - 1) umma_commit(kv_empty + ((g / nqt) & 1));  // item done: free its K/V buffer
            }
        }
    } else if (warp < 10) {
        // ===================== softmax groups =====================
        const int hh = (warp - 2) >> 2;             // key half owned by this group
        const int quarter = warp & 3;               // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;          // query row inside the tile
        const int bar_id = 1 + hh;
        uint8_t* cflag = reinterpret_cast<uint8_t*>(sm + L.red_off + 32 * 4) + hh * 8;           // this half's chunk flags
        uint32_t* cbits = reinterpret_cast<uint32_t*>(sm + L.red_off + 32 * 4 + 16) + hh * 8;   // and key bitmasks
        float2* exch = reinterpret_cast<float2*>(sm + L.exch_off);                  // [parity][half][row] (c, l)
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tS = tmem + lane_sel + kAtColS + hh * kAtMaxKh;
        const int nchunk = Kh / 32;
        int g = 0;
        for (int b = b_first; b < B; b += ngroups) {
            // ---- per-item chunk flags and key bitmasks of this half (private to the group: no cross-group synchronisation)
            bar_sync_n(bar_id, 128);  // everyone is done with the previous item's table
            bool mine_clear = true, mine_on = false;
            for (int j = r; j < 2 * Kh; j += 128) {  // a warp covers one 32-key chunk per step
                const bool on = j < S && mask[static_cast<int64_t>(b) * S + j] != 0;
                const uint32_t bal = __ballot_sync(0xffffffff, on);
                if (lane == 0 && j >= hh * Kh && j < (hh + 1) * Kh) {
                    cflag[(j - hh * Kh) >> 5] = bal == 0xffffffffu ? 0 : (bal != 0u ? 1 : 2);
                    cbits[(j - hh * Kh) >> 5] = bal;
                }
                mine_clear &= on;
                mine_on |= on;
            }
            const bool clear = bar_red_and(bar_id, 128, mine_clear);  // no masked / out-of-range key at all
            const bool any_on = bar_red_or(bar_id, 128, mine_on);    // (the barriers also publish the flags)
            int n_eff = nchunk;  // chunks up to the last one of this half with a valid key
            if (!clear)
                while (n_eff > 0 && cflag[n_eff - 1] == 2) --n_eff;
            for (int t = 0; t < nqt; ++t, ++g) {
                const uint32_t ph = g & 1;
                const int i = t * kAtQT + r;
                const int start = hh * Kh - i + OFF;  // >= 1; bias of key column c is T0[start + c]
                const float2* pb2 = reinterpret_cast<const float2*>((start & 1) ? T1 + (start - 1) : T0 + start);
                mbar_wait(s_full + hh, ph);
                tc_fence_after();
                float c, l;
                if (!any_on) {
                    // Every key masked: the reference adds finfo.min to all scores, which absorbs them
                    // in fp32 -> uniform attention over the S keys: p = 1 for keys < S, 0 beyond.
                    l = 0.f;
                    c = 0.f;
                    for (int cc = 0; cc < nchunk; ++cc) {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const int key = hh * Kh + cc * 32 + j;
                            const float p0 = key < S ? 1.f : 0.f, p1 = key + 1 < S ? 1.f : 0.f;
                            l += p0 + p1;
                            pk[j >> 1] = pack16x2<kF16>(p0, p1);
                        }
                        tmem_st_32x16(tS + cc * 16, pk);
                    }
                } else if (clear) {
                    softmax_tile<kF16, false>(tS, nchunk, scale_log2e, bmax, pb2, nullptr, c, l);
                } else {
                    mask_prepass(tS, n_eff, cbits, cflag);
                    c = 0.f;
                    l = 0.f;
                    if (n_eff > 0) softmax_tile<kF16, false>(tS, n_eff, scale_log2e, bmax, pb2, nullptr, c, l);
                    mask_postpass<kF16>(tS, n_eff, nchunk);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(p_ready + hh);
                constexpr int kCombiner = kDefer ? 1 : 0;
                if (hh != kCombiner) {
                    exch[ph * kAtQT + r] = make_float2(c, l);
                    mbar_arrive(ml_ready + ph);  // release: the smem write above is ordered before the arrive
                    continue;
                }
                // ---- combining group: merge the halves, normalise, store this row of ctx
                mbar_wait(ml_ready + ph, (g >> 1) & 1);
                const float2 e1 = exch[ph * kAtQT + r];
                const float m = fmaxf(c, e1.x);
                const float a_mine = at_exp2(c - m), a_other = at_exp2(e1.x - m);
                const float inv = __fdividef(1.f, a_mine * l + a_other * e1.y);
                const float w0 = (kCombiner == 0 ? a_mine : a_other) * inv;  // weight of O[0]
                const float w1 = (kCombiner == 0 ? a_other : a_mine) * inv;  // weight of O[1]
                const uint32_t tO = tmem + lane_sel + kAtColO;
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
                        mbar_arrive(o_free);  // O is in registers: the next tile's PV may overwrite it
                    }
                    if (i < S) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t w[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int d = q * 8 + e * 2;
                                w[e] = pack16x2<kF16>(w0 * __uint_as_float(o0[d]) + w1 * __uint_as_float(o1[d]),
                                                      w0 * __uint_as_float(o0[d + 1]) + w1 * __uint_as_float(o1[d + 1]));
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

bool attention_tc_supported(int S, int dh) { return dh == 64 && S >= 1 && S <= 2 * kAtMaxKh; }

int launch_attention_tc(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                        h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream) {
    ARB_REQUIRE(qkv && rel_bias && mask && ctx, "attention_tc: null pointer");
    ARB_REQUIRE(attention_tc_supported(S, dh), "attention_tc: S=%d dh=%d unsupported", S, dh);
    ARB_REQUIRE(B > 0 && S <= max_rel, "attention_tc: bad shape B=%d S=%d max_rel=%d", B, S, max_rel);
    const int H = heads * dh;
    const int Kh = ((S + 1) / 2 + 31) / 32 * 32;  // multiple of 32: whole 32-column TMEM chunks
    const int nqt = (S + kAtQT - 1) / kAtQT;
    const AtLayout L = at_layout(Kh, nqt);
    const int smem = L.total + 1024;
    ARB_REQUIRE(smem <= 232448, "attention_tc: shared memory %d exceeds 227 KB", smem);
    CUtensorMap tq, tkv;
    if (!make_tmap_bf16_batched_k64(&tq, qkv, B, S, 3 * H, 3 * H, kAtQT) ||
        !make_tmap_bf16_batched_k64(&tkv, qkv, B, S, 3 * H, 3 * H, Kh)) {
        set_error("attention_tc: cuTensorMapEncodeTiled failed");
        return ARB_ERR_CUDA;
    }
    // ARB_ATTN_DEFER=0 selects the lockstep schedule (group 0 combines) for A/B runs
    static const bool defer = []() {
        const char* e = getenv("ARB_ATTN_DEFER");
        return !(e && e[0] == '0');
    }();
    auto kern = fp16 ? (defer ? attention_tc_kernel<true, true> : attention_tc_kernel<true, false>)
                     : (defer ? attention_tc_kernel<false, true> : attention_tc_kernel<false, false>);
    ARB_CHECK_CUDA(set_max_smem_once(kern, smem));
    int ngroups = num_sms() / heads;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > B) ngroups = B;
    const int grid = ngroups * heads;
    const float scale_log2e = kAtLog2e / sqrtf(static_cast<float>(dh));
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(kAtThreads), smem, stream, 1, tq, tkv, rel_bias, max_rel, mask, ctx, B, S,
                                 heads, Kh, scale_log2e));
    return ARB_OK;
}

}  // namespace arb
