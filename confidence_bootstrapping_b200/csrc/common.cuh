// Shared helpers for the cb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CB_OK 0
#define CB_ERR_ARG 1
#define CB_ERR_CUDA 2

void cb_set_error(const char* fmt, ...);

#define CB_CHECK_ARG(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            cb_set_error(__VA_ARGS__);     \
            return CB_ERR_ARG;             \
        }                                  \
    } while (0)

#define CB_CHECK_LAUNCH(name)                                                  \
    do {                                                                       \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess) {                                              \
            cb_set_error("%s: %s", name, cudaGetErrorString(e__));             \
            return CB_ERR_CUDA;                                                \
        }                                                                      \
    } while (0)

static inline int cb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// number of SMs on a B200; grids of persistent / grid-stride kernels are sized in multiples of it
#define CB_NUM_SMS 148

__device__ __forceinline__ float cb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
