// Shared helpers for the ces_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include "../../include/ces_b200.h"   // status codes CES_OK / CES_ERR_*

namespace ces {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launches;   // kernels launched by this library (ces_launch_count); host threads may race

// Ordinal of the calling thread's current device, clamped to the size of the per-device flag tables below.
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per DEVICE, not per process: every
// "configured once" flag in this library is indexed by this slot.
constexpr int kMaxDevices = 64;
inline int device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
    return code;
}

#define CES_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            snprintf(::ces::g_last_error, sizeof(::ces::g_last_error), "%s:%d %s -> %s", \
                     __FILE__, __LINE__, #expr, cudaGetErrorString(_e));                 \
            return CES_ERR_CUDA;                                                  \
        }                                                                                \
    } while (0)

// After a kernel launch: count it and surface launch-configuration errors.
#define CES_LAUNCHED(n)                  \
    do {                                 \
        ::ces::g_launches.fetch_add((n), std::memory_order_relaxed); \
        CES_CUDA(cudaGetLastError());    \
    } while (0)

#define CES_TRY(expr)               \
    do {                            \
        int _s = (expr);            \
        if (_s != 0) return _s;     \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// Leading dimension used for every internal matrix: rows start 128 B aligned.
inline int64_t padded_ld(int64_t cols) { return round_up(cols < 1 ? 1 : cols, 16); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum, result valid in thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? scratch[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? scratch[lane] : 0.0;
        v = warp_max(v);
    }
    return v;
}

}  // namespace ces
