// Error plumbing shared by every translation unit behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/arxiv_rag_b200.h"  // ARB_OK / ARB_ERR_* codes

namespace arb {

void set_error(const char* fmt, ...);
const char* get_error();

#define ARB_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            arb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                         \
            return ARB_ERR_CUDA;                                                         \
        }                                                                                     \
    } while (0)

#define ARB_REQUIRE(cond, ...)               \
    do {                                     \
        if (!(cond)) {                       \
            arb::set_error(__VA_ARGS__);     \
            return ARB_ERR_INVALID;     \
        }                                    \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel, device and size instead of on
// every launch (the attribute is sticky; the encoder launches 63 kernels per forward).
template <class Kern>
inline cudaError_t set_max_smem_once(Kern kern, int bytes) {
    struct Slot { const void* fn; int dev; int bytes; };
    static thread_local Slot cache[64];
    static thread_local int used = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kern);
    for (int i = 0; i < used; ++i)
        if (cache[i].fn == fn && cache[i].dev == dev && cache[i].bytes >= bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) {
        if (used == 64) used = 0;
        cache[used++] = Slot{fn, dev, bytes};
    }
    return e;
}

inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace arb
