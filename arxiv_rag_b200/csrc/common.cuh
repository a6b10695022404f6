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
