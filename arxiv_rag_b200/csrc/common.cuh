// Error plumbing shared by every translation unit behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

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
// every launch (the attribute is sticky; the encoder launches 62 kernels per forward).
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

// Programmatic dependent launch for the encoder's kernel chain. While `pdl_scope` is alive on this
// thread, `launch_kernel` adds cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel
// may start (set-up, TMEM allocation, weight prefetch into L2) while its predecessor is still running,
// and blocks in `griddepcontrol.wait` (ptx.cuh: pdl_wait) before it touches anything the predecessor
// wrote. Every kernel launched this way executes pdl_wait on all threads, so completion stays
// transitive along the chain. Used for small token counts (query-time latency), where the forward is
// 62 short kernels and launch + set-up are a third of each.
inline bool& pdl_flag() {
    static thread_local bool on = false;
    return on;
}
// ARB_PDL=0 never, 2 always, default 1: where the launcher says the call is latency-bound.
int& pdl_mode_ref();  // capi.cu: process-wide, initialised from ARB_PDL, set by arb_set_pdl_mode
inline int pdl_env_mode() { return pdl_mode_ref(); }
struct pdl_scope {
    bool prev;
    explicit pdl_scope(bool latency_bound) : prev(pdl_flag()) {
        pdl_flag() = pdl_env_mode() == 2 || (pdl_env_mode() == 1 && latency_bound);
    }
    ~pdl_scope() { pdl_flag() = prev; }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_flag()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
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
