// MUFU rate check: ex2.approx.ftz.f32 vs ex2.approx.f16x2 (two results per instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ex2_rate tools/microbench/ex2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void k_f32(float* out, int iters) {
    float a = threadIdx.x * 1e-3f - 0.5f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    for (int i = 0; i < iters; ++i) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
        a -= 1.0f; b -= 1.0f; c -= 1.0f; d -= 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void k_f16x2(float* out, int iters) {
    uint32_t a = 0x34003800u + threadIdx.x, b = a + 7, c = a + 11, d = a + 13;
    for (int i = 0; i < iters; ++i) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(c));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(d));
        a ^= 0x00010001u; b ^= 0x00010001u; c ^= 0x00010001u; d ^= 0x00010001u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(a ^ b ^ c ^ d);
}
int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_f32<<<148 * 4, 512>>>(out, iters);
            else k_f16x2<<<148 * 4, 512>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double instr = 148.0 * 4 * 512 * iters * 4;
            if (rep == 1)
                printf("%s: %.3f ms, %.1f G MUFU instr/s per lane-op, %.2f results/clk/SM at 1.9 GHz\n", which ? "ex2.f16x2" : "ex2.f32  ", ms,
                       instr / ms / 1e6, instr * (which ? 2 : 1) / (ms * 1e-3) / 148 / 1.9e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
