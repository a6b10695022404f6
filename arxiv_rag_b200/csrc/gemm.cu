// Encoder GEMMs: C = epi(A . B^T) on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
// Replaces the nn.Linear calls of transformers' MPNet (modeling_mpnet.py:145-159 q/k/v,
// :183 o, :225-228 intermediate+GELU, :239-243 output) that sentence-transformers' encode
// reaches from generate_embeddings_parallel.py:146-153.
#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"
#include "umma_pipe.cuh"

namespace arb {

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

// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), far below the 16-bit output rounding.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    const float e = exp2f(-z * z * 1.4426950408889634f);
    const float erf_abs = fmaf(-p, e, 1.0f);
    const float erf_v = copysignf(erf_abs, x);
    return 0.5f * x * (1.0f + erf_v);
}

// OutT = h16 (16-bit activations in the kF16 format) or float.
template <int BN, int STAGES, int EPI, bool kF16, typename OutT>
__global__ void __launch_bounds__(kPipeThreads, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              OutT* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
              const h16* __restrict__ R, int64_t ldr, int64_t M, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles need a 1024-byte aligned base; align by hand (the launcher adds slack).
    PipeSmem<BN, STAGES> sm{smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)};
    const int warp = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
    const int lane = threadIdx.x & 31;

    const int num_m = static_cast<int>((M + kBM - 1) / kBM);
    const int num_n = (N + BN - 1) / BN;
    const int kblocks = (K + kBK - 1) / kBK;
    GemmTileIter it{static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), num_m * num_n,
                    num_n, BN};

    const uint32_t tmem_base = pipe_setup(sm, warp, &tmap_a, &tmap_b);

    if (warp == 0) {
        if (elect_one()) pipe_produce(sm, &tmap_a, &tmap_b, it, kblocks, kEvictNormal, kEvictLast);
    } else if (warp == 1) {
        if (elect_one()) pipe_mma<BN, STAGES, kF16>(sm, tmem_base, it, kblocks);
    } else {
        // Epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32).
        const int lane_grp = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        int row_a, row_b;
        while (it.next(row_a, row_b)) {
            mbar_wait(sm.tmem_full(acc), acc_phase);
            tc_fence_after();
            const int64_t row = static_cast<int64_t>(row_a) + lane_grp * 32 + lane;
            const bool row_ok = row < M;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c0, r);
                tmem_ld_wait();
                const int col = row_b + c0;
                if (col < N) {  // N % 32 == 0 is enforced by the launcher
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(bp + j);
                            v[4 * j + 0] += b4.x;
                            v[4 * j + 1] += b4.y;
                            v[4 * j + 2] += b4.z;
                            v[4 * j + 3] += b4.w;
                        }
                    }
                    if (EPI == EPI_BIAS_GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (row_ok) {
                        if (EPI == EPI_BIAS_RESIDUAL) {
                            const uint4* rp = reinterpret_cast<const uint4*>(R + row * ldr + col);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 u = __ldg(rp + j);
                                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 f = unpack16x2<kF16>(w[q]);
                                    v[8 * j + 2 * q] += f.x;
                                    v[8 * j + 2 * q + 1] += f.y;
                                }
                            }
                        }
                        if constexpr (sizeof(OutT) == 2) {
                            uint4* cp = reinterpret_cast<uint4*>(C + row * ldc + col);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 u;
                                u.x = pack16x2<kF16>(v[8 * j + 0], v[8 * j + 1]);
                                u.y = pack16x2<kF16>(v[8 * j + 2], v[8 * j + 3]);
                                u.z = pack16x2<kF16>(v[8 * j + 4], v[8 * j + 5]);
                                u.w = pack16x2<kF16>(v[8 * j + 6], v[8 * j + 7]);
                                cp[j] = u;
                            }
                        } else {
                            float4* cp = reinterpret_cast<float4*>(C + row * ldc + col);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                cp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2],
                                                    v[4 * j + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(sm.tmem_empty(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }
    pipe_teardown(sm, warp, tmem_base);
}

template <int BN, int STAGES, int EPI, bool kF16, typename OutT>
static int launch_gemm_impl(const h16* A, int64_t lda, const h16* B, int64_t ldb, OutT* C,
                            int64_t ldc, const float* bias, const h16* R, int64_t ldr, int64_t M,
                            int N, int K, cudaStream_t stream) {
    CUtensorMap ta, tb;
    // the TMA element type only matters for OOB fill; both 16-bit formats move as raw 2-byte words
    if (!make_tmap_bf16_k64(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K),
                            static_cast<uint64_t>(lda), kBM) ||
        !make_tmap_bf16_k64(&tb, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                            static_cast<uint64_t>(ldb), BN)) {
        set_error("cuTensorMapEncodeTiled failed (A %p lda %lld, B %p ldb %lld)", (const void*)A,
                  (long long)lda, (const void*)B, (long long)ldb);
        return ARB_ERR_CUDA;
    }
    auto kern = gemm16_kernel<BN, STAGES, EPI, kF16, OutT>;
    constexpr int smem = PipeSmem<BN, STAGES>::kExtraOffset + 1024;  // +1024: alignment slack
    ARB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t tiles = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN);
    const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    kern<<<grid, kPipeThreads, smem, stream>>>(ta, tb, C, ldc, bias, R, ldr, M, N, K);
    ARB_CHECK_CUDA(cudaGetLastError());
    return ARB_OK;
}

static int check_gemm_args(const void* A, int64_t lda, const void* B, int64_t ldb, const void* C,
                           int64_t ldc, int64_t M, int N, int K) {
    ARB_REQUIRE(A && B && C, "gemm: null operand");
    ARB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%lld N=%d K=%d", (long long)M, N, K);
    ARB_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
    ARB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0,
                "gemm: K/lda/ldb/ldc must be multiples of 8 (16-byte rows)");
    ARB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
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
            return launch_gemm_impl<256, 4, EPI_BIAS, kF16, h16>(A, lda, B, ldb, C, ldc, bias, nullptr,
                                                                 0, M, N, K, stream);
        case EPI_BIAS_GELU:
            return launch_gemm_impl<256, 4, EPI_BIAS_GELU, kF16, h16>(A, lda, B, ldb, C, ldc, bias,
                                                                      nullptr, 0, M, N, K, stream);
        case EPI_BIAS_RESIDUAL:
            ARB_REQUIRE(R != nullptr && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0,
                        "gemm: residual operand missing or misaligned");
            return launch_gemm_impl<256, 4, EPI_BIAS_RESIDUAL, kF16, h16>(A, lda, B, ldb, C, ldc, bias,
                                                                          R, ldr, M, N, K, stream);
        default:
            set_error("gemm: unknown epilogue %d", epilogue);
            return ARB_ERR_INVALID;
    }
}

int launch_gemm16(const h16* A, int64_t lda, const h16* B, int64_t ldb, h16* C, int64_t ldc,
                  const float* bias, const h16* R, int64_t ldr, int64_t M, int N, int K,
                  int epilogue, bool fp16, cudaStream_t stream) {
    int rc = check_gemm_args(A, lda, B, ldb, C, ldc, M, N, K);
    if (rc) return rc;
    ARB_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                "gemm: bias must be 16-byte aligned");
    return fp16 ? dispatch_gemm16<true>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, epilogue, stream)
                : dispatch_gemm16<false>(A, lda, B, ldb, C, ldc, bias, R, ldr, M, N, K, epilogue, stream);
}

int launch_gemm16_f32out(const h16* A, int64_t lda, const h16* B, int64_t ldb, float* C,
                         int64_t ldc, int64_t M, int N, int K, bool fp16, cudaStream_t stream) {
    int rc = check_gemm_args(A, lda, B, ldb, C, ldc, M, N, K);
    if (rc) return rc;
    return fp16 ? launch_gemm_impl<256, 4, EPI_BIAS, true, float>(A, lda, B, ldb, C, ldc, nullptr,
                                                                  nullptr, 0, M, N, K, stream)
                : launch_gemm_impl<256, 4, EPI_BIAS, false, float>(A, lda, B, ldb, C, ldc, nullptr,
                                                                   nullptr, 0, M, N, K, stream);
}

}  // namespace arb
