// Shared device/host helpers for the spherehand_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define SH_EXPORT extern "C" __attribute__((visibility("default")))

#define SH_OK 0
#define SH_ERR_INVALID 1
#define SH_ERR_CUDA 2
#define SH_ERR_UNSUPPORTED 3

#define SH_NUM_SMS 148

// Last error text (thread-unsafe by design: the reference path is single-threaded).
extern char g_sh_last_error[512];
// Number of kernel launches issued through this library (every launch site ends in SH_CHECK_LAUNCH).
extern unsigned long long g_sh_launches;

#define SH_REQUIRE(cond, ...)                                                   \
    do {                                                                        \
        if (!(cond)) {                                                          \
            snprintf(g_sh_last_error, sizeof(g_sh_last_error), __VA_ARGS__);    \
            return SH_ERR_INVALID;                                              \
        }                                                                       \
    } while (0)

#define SH_CHECK_LAUNCH(name)                                                                       \
    do {                                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                       \
        ++g_sh_launches;                                                                            \
        if (e__ != cudaSuccess) {                                                                   \
            snprintf(g_sh_last_error, sizeof(g_sh_last_error), "%s: %s", name, cudaGetErrorString(e__)); \
            return SH_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

#define SH_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            snprintf(g_sh_last_error, sizeof(g_sh_last_error), "%s: %s", #call, cudaGetErrorString(e__)); \
            return SH_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

static inline int sh_div_up(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The reference pixel grid (mesh/render.py:31-32): ((u - W/2) * 300) / W, each op rounded on its own.
__device__ __forceinline__ float sh_grid_mm(int u, float half, float size) {
    return __fdiv_rn(__fmul_rn(__fsub_rn((float)u, half), 300.0f), size);
}

#define SH_BACKGROUND 100.0f
#define SH_S_MIN 0.01f
