#include <cstdlib>
// Fused self-attention, mma.sync version: softmax(q.k^T/sqrt(dh) + rel_bias + mask) . v
// per (sequence, head), never materialising the [S,S] probability matrix in HBM.
// Follows MPNetSelfAttention.forward (modeling_mpnet.py:162-177) with the shared relative
// position bias of MPNetEncoder.compute_position_bias (:324-360) and the additive mask
// (1-m)*finfo.min of get_extended_attention_mask (modeling_utils.py:936-947). With a NULL bias
// table it is BertSelfAttention (the all-MiniLM-L6-v2 encoder of the second encode call site,
// 3-chunks/pipeline/src/processors/text_processor.py:1379-1396). Head dim 64 or 32.
//
// One CTA per (sequence, head): K and V of the head ([S,dh] 16-bit each) are staged once in
// shared memory (XOR-swizzled 16-byte chunks, conflict-free for ldmatrix); each warp then
// runs a flash-style online softmax over 16-query-row blocks with mma.sync tiles and fp32
// statistics/accumulators. The relative-position bias depends only on j-i, so a thread reads it
// as one contiguous run of a padded table (rows g and g+8 share the run shifted by 8 keys).
// Keys after the last unmasked key contribute exactly 0 in the reference's fp32 arithmetic, so
// key blocks beyond it are skipped (unless the whole row is masked: uniform attention).
// The encode path uses the tcgen05 kernel (attention_tc.cu) when 64 <= S <= 384 and dh == 64.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace arb {

// shortest sequence the encoder sends to the tcgen05 attention in auto mode
constexpr int kAttnTcMinSeq = 64;
constexpr int kAttnSplitMinSeq = 192;  // auto: attention_tc3 above this length

constexpr int kAttnThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMaskMin = -3.4028234663852886e38f;  // torch.finfo(float32).min

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                            uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                                  uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
template <bool kF16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                          uint32_t b1) {
    if constexpr (kF16) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
            "{%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
            "{%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {  // single MUFU.EX2; exp2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// byte offset of 16-byte chunk c of key row r: rows are DH*2 bytes; the XOR keeps the 8 rows an
// ldmatrix touches on 8 different bank groups (128-byte rows: r&7; 64-byte rows: (r>>1)&3)
template <int DH>
__device__ __forceinline__ uint32_t kv_off(int r, int c) {
    if constexpr (DH == 64) return r * 128 + ((c ^ (r & 7)) << 4);
    else return r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
}

template <bool kF16, int DH>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const h16* __restrict__ qkv, const float* __restrict__ rel_bias, int max_rel,
                 const int32_t* __restrict__ mask, h16* __restrict__ ctx, int S, int heads,
                 float scale_log2e) {
    constexpr int CPR = DH / 8;    // 16-byte chunks per row
    constexpr int KS = DH / 16;    // k-steps of q.k^T
    constexpr int DB = DH / 8;     // 8-wide output dim blocks
    constexpr int RB = DH * 2;     // row bytes
    extern __shared__ __align__(128) uint8_t smem_attn[];
    const int Spad = (S + 63) & ~63;
    const int OFF = Spad + 16;                                // bias table: entry (j - i) + OFF
    const int nbias = 2 * Spad + 32;
    uint8_t* sK = smem_attn;                                  // [Spad][RB], swizzled
    uint8_t* sV = sK + static_cast<size_t>(Spad) * RB;        // [Spad][RB], swizzled
    float* sBias = reinterpret_cast<float*>(sV + static_cast<size_t>(Spad) * RB);  // [nbias], x log2e
    float* sMask = sBias + nbias;                             // [Spad] additive, log2 domain
    __shared__ int s_last;                                    // index of the last unmasked key, -1 if none
    __shared__ int s_blk_clear[16];                           // per 64-key block: 1 = no masked / out-of-range key

    const int h = blockIdx.x, b = blockIdx.y;
    const int H = heads * DH;
    const int64_t ld = 3 * static_cast<int64_t>(H);
    const h16* base = qkv + static_cast<int64_t>(b) * S * ld;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    pdl_launch_dependents();
    pdl_wait();
    // ---- stage K, V (zero rows beyond S), bias table and mask
    if (tid == 0) s_last = -1;
    if (tid < 16) s_blk_clear[tid] = 1;
    const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
    for (int idx = tid; idx < Spad * 2 * CPR; idx += kAttnThreads) {
        const int r = idx / (2 * CPR), w = idx % (2 * CPR), c = w % CPR, isv = w / CPR;
        const uint32_t off = kv_off<DH>(r, c);
        if (r < S) {
            cp_async16((isv ? sV_u : sK_u) + off, base + static_cast<int64_t>(r) * ld + (isv ? 2 * H : H) + h * DH + c * 8);
        } else {
            *reinterpret_cast<uint4*>((isv ? sV : sK) + off) = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < nbias; i += kAttnThreads) {
        const int rel = i - OFF;  // j - i_query
        sBias[i] = (rel_bias != nullptr && rel > -S && rel < S)
                       ? rel_bias[static_cast<int64_t>(h) * (2 * max_rel - 1) + rel + (max_rel - 1)] * kLog2e
                       : 0.f;
    }
    __syncthreads();  // s_last initialised
    int my_last = -1;
    for (int j = tid; j < Spad; j += kAttnThreads) {
        float m = -INFINITY;  // keys beyond the sequence never contribute
        if (j < S) {
            const bool on = mask[static_cast<int64_t>(b) * S + j] != 0;
            m = on ? 0.f : kMaskMin;
            if (on) my_last = j;
        }
        sMask[j] = m;
        if (m != 0.f) s_blk_clear[j >> 6] = 0;  // benign race: every writer stores 0
    }
    if (my_last >= 0) atomicMax(&s_last, my_last);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // keys in (last, S) are masked: exactly zero probability unless every key is masked
    const int kv_len = s_last >= 0 ? s_last + 1 : S;
    const int kv_end = (kv_len + 63) & ~63;

    const int g = lane >> 2, t = lane & 3;
    const int nqb = (S + 15) / 16;
    for (int qb = warp; qb < nqb; qb += kAttnThreads / 32) {
        const int q0 = qb * 16;
        const int i0 = q0 + g, i1 = q0 + g + 8;
        h16* out0 = ctx + (static_cast<int64_t>(b) * S + i0) * H + h * DH + 2 * t;
        h16* out1 = ctx + (static_cast<int64_t>(b) * S + i1) * H + h * DH + 2 * t;
        if (q0 >= kv_len && s_last >= 0) {
            // query rows past the last real token are never pooled; keep them finite (zeros)
#pragma unroll
            for (int db = 0; db < DB; ++db) {
                if (i0 < S) *reinterpret_cast<uint32_t*>(out0 + db * 8) = 0u;
                if (i1 < S) *reinterpret_cast<uint32_t*>(out1 + db * 8) = 0u;
            }
            continue;
        }
        const int r0 = min(i0, S - 1), r1 = min(i1, S - 1);
        // Q fragments straight from global in the m16n8k16 A layout
        uint32_t qa[KS][4];
        {
            const h16* q_r0 = base + static_cast<int64_t>(r0) * ld + h * DH;
            const h16* q_r1 = base + static_cast<int64_t>(r1) * ld + h * DH;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                qa[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q_r0 + ks * 16 + 2 * t));
                qa[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q_r1 + ks * 16 + 2 * t));
                qa[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q_r0 + ks * 16 + 8 + 2 * t));
                qa[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q_r1 + ks * 16 + 8 + 2 * t));
            }
        }
        float o[DB][4];
#pragma unroll
        for (int i = 0; i < DB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        // bias run of this thread for key block kb: pb[8*n], pb[8*n+1], n = -1..7
        const float* pb0 = sBias + (OFF - i0 + 2 * t);

        for (int kb = 0; kb < kv_end; kb += 64) {
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            // ---- S = Q.K^T
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int key = kb + nb * 8 + (lane & 7);
#pragma unroll
                for (int kp = 0; kp < KS / 2; ++kp) {
                    uint32_t b0, b1, b2, b3;
                    const int chunk = 4 * kp + (lane >> 3);
                    ldmatrix_x4(sK_u + kv_off<DH>(key, chunk), b0, b1, b2, b3);
                    mma_16816<kF16>(s[nb], qa[2 * kp], b0, b1);
                    mma_16816<kF16>(s[nb], qa[2 * kp + 1], b2, b3);
                }
            }
            // ---- scale + relative-position bias (+ mask) in the log2 domain, block row max
            const float* pb = pb0 + kb;
            float e0 = pb[-8], e1 = pb[-7];  // bias pair of the previous 8-key group (row g+8)
            float bm0 = -INFINITY, bm1 = -INFINITY;
            const bool masked_blk = s_blk_clear[kb >> 6] == 0;  // CTA-uniform
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const float c0 = pb[nb * 8], c1 = pb[nb * 8 + 1];
                s[nb][0] = fmaf(s[nb][0], scale_log2e, c0);
                s[nb][1] = fmaf(s[nb][1], scale_log2e, c1);
                s[nb][2] = fmaf(s[nb][2], scale_log2e, e0);
                s[nb][3] = fmaf(s[nb][3], scale_log2e, e1);
                e0 = c0;
                e1 = c1;
                if (masked_blk) {
                    const float2 mk = *reinterpret_cast<const float2*>(sMask + kb + nb * 8 + 2 * t);
                    s[nb][0] += mk.x;
                    s[nb][1] += mk.y;
                    s[nb][2] += mk.x;
                    s[nb][3] += mk.y;
                }
                bm0 = fmaxf(bm0, fmaxf(s[nb][0], s[nb][1]));
                bm1 = fmaxf(bm1, fmaxf(s[nb][2], s[nb][3]));
            }
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffff, bm0, 1));
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffff, bm0, 2));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffff, bm1, 1));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffff, bm1, 2));
            const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
            const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);
            m0 = mn0;
            m1 = mn1;
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                s[nb][0] = fast_exp2(s[nb][0] - mn0);
                s[nb][1] = fast_exp2(s[nb][1] - mn0);
                s[nb][2] = fast_exp2(s[nb][2] - mn1);
                s[nb][3] = fast_exp2(s[nb][3] - mn1);
                rs0 += s[nb][0] + s[nb][1];
                rs1 += s[nb][2] + s[nb][3];
            }
            l0 = l0 * c0 + rs0;
            l1 = l1 * c1 + rs1;
#pragma unroll
            for (int db = 0; db < DB; ++db) {
                o[db][0] *= c0;
                o[db][1] *= c0;
                o[db][2] *= c1;
                o[db][3] *= c1;
            }
            // ---- O += P.V
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t pa[4];
                pa[0] = pack16x2<kF16>(s[2 * kk][0], s[2 * kk][1]);
                pa[1] = pack16x2<kF16>(s[2 * kk][2], s[2 * kk][3]);
                pa[2] = pack16x2<kF16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                pa[3] = pack16x2<kF16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
                const int key = kb + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                for (int dp = 0; dp < DB / 2; ++dp) {
                    uint32_t b0, b1, b2, b3;
                    const int chunk = 2 * dp + (lane >> 4);
                    ldmatrix_x4_trans(sV_u + kv_off<DH>(key, chunk), b0, b1, b2, b3);
                    mma_16816<kF16>(o[2 * dp], pa, b0, b1);
                    mma_16816<kF16>(o[2 * dp + 1], pa, b2, b3);
                }
            }
        }
        // ---- finalise: row sums across the 4 lanes of a quad, normalise, store
        l0 += __shfl_xor_sync(0xffffffff, l0, 1);
        l0 += __shfl_xor_sync(0xffffffff, l0, 2);
        l1 += __shfl_xor_sync(0xffffffff, l1, 1);
        l1 += __shfl_xor_sync(0xffffffff, l1, 2);
        const float inv0 = __fdividef(1.f, l0), inv1 = __fdividef(1.f, l1);
#pragma unroll
        for (int db = 0; db < DB; ++db) {
            if (i0 < S) *reinterpret_cast<uint32_t*>(out0 + db * 8) = pack16x2<kF16>(o[db][0] * inv0, o[db][1] * inv0);
            if (i1 < S) *reinterpret_cast<uint32_t*>(out1 + db * 8) = pack16x2<kF16>(o[db][2] * inv1, o[db][3] * inv1);
        }
    }
}

int launch_attention_mma(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                         h16* ctx, int B, int S, int heads, int dh, bool fp16, cudaStream_t stream) {
    ARB_REQUIRE(qkv && mask && ctx, "attention: null pointer");
    ARB_REQUIRE(dh == 64 || dh == 32, "attention: head dim %d unsupported (64 or 32)", dh);
    ARB_REQUIRE(B > 0 && S > 0 && S <= 768 && (rel_bias == nullptr || S <= max_rel) && ((S + 63) / 64) <= 16,
                "attention: bad shape B=%d S=%d max_rel=%d", B, S, max_rel);
    ARB_REQUIRE(B <= 65535, "attention: batch %d exceeds grid.y", B);
    const int Spad = (S + 63) & ~63;
    const size_t smem = static_cast<size_t>(Spad) * 4 * dh + (2 * Spad + 32 + Spad) * sizeof(float);
    void (*kern)(const h16*, const float*, int, const int32_t*, h16*, int, int, float);
    if (dh == 64) kern = fp16 ? attention_kernel<true, 64> : attention_kernel<false, 64>;
    else kern = fp16 ? attention_kernel<true, 32> : attention_kernel<false, 32>;
    ARB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const float scale_log2e = kLog2e / sqrtf(static_cast<float>(dh));
    ARB_CHECK_CUDA(launch_kernel(kern, dim3(heads, B), dim3(kAttnThreads), smem, stream, 1, qkv, rel_bias, max_rel, mask, ctx, S,
                                 heads, scale_log2e));
    return ARB_OK;
}

// impl: 0 = auto, 1 = mma.sync, 2 = tcgen05, 3 = tcgen05 with sub-block pipelining (experiment, only in
// builds with -DARB_WITH_ATTENTION_TC2), 4 = tcgen05 with 16 softmax warps (attention_tc3.cu). impl 2 and 4
// cover head dim 64, 1 <= S <= 384; auto picks 2 for 64 <= S <= 192, 4 for 192 < S <= 384, else mma.sync;
// ARB_ATTN_IMPL overrides auto for A/B runs.
int launch_attention(const h16* qkv, const float* rel_bias, int max_rel, const int32_t* mask,
                     h16* ctx, int B, int S, int heads, int dh, bool fp16, int impl, cudaStream_t stream) {
    ARB_REQUIRE(impl >= 0 && impl <= 4, "attention: impl %d must be 0 (auto), 1 (mma.sync), 2, 3 or 4 (tcgen05)", impl);
    if (impl == 0) {
        static const int forced = []() {
            const char* e = getenv("ARB_ATTN_IMPL");
            return e ? atoi(e) : 0;
        }();
        // 16 softmax warps (attention_tc3.cu) above S = 192: 5-9 % faster than the 8-warp kernel at
        // S = 272..384, 1-4 % at 224..256 (4 % on ragged batches, +0.7 % on an all-valid S = 256 one),
        // slower from 192 down (fewer columns per thread than latency to hide)
        impl = forced >= 1 && forced <= 4 ? forced : (S > kAttnSplitMinSeq ? 4 : 2);
        if (impl == 4 && !(rel_bias != nullptr && attention_tc3_supported(S, dh) && S >= kAttnTcMinSeq)) impl = 1;
        if (impl == 3 && !(rel_bias != nullptr && attention_tc2_supported(S, dh) && S >= kAttnTcMinSeq)) impl = 1;
        // the tcgen05 kernel also runs S < 64 (parity-tested), but there a forward takes the same time with
        // either kernel (443 vs 440 us at Q=1, S=16): auto keeps mma.sync below one 64-key block
        if (impl == 2 && !(rel_bias != nullptr && attention_tc_supported(S, dh) && S >= kAttnTcMinSeq)) impl = 1;
    }
    if (impl == 4) return launch_attention_tc3(qkv, rel_bias, max_rel, mask, ctx, B, S, heads, dh, fp16, stream);
    if (impl == 3) return launch_attention_tc2(qkv, rel_bias, max_rel, mask, ctx, B, S, heads, dh, fp16, stream);
    if (impl == 2) return launch_attention_tc(qkv, rel_bias, max_rel, mask, ctx, B, S, heads, dh, fp16, stream);
    return launch_attention_mma(qkv, rel_bias, max_rel, mask, ctx, B, S, heads, dh, fp16, stream);
}

}  // namespace arb
