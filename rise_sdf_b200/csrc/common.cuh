// Shared helpers for the sm_100a kernels of librsdf_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rsdf_b200.h"

#define RSDF_EBADARG (-1)
#define RSDF_ECAPACITY (-2)

#define RSDF_LAUNCH_CHECK()                         \
    do {                                            \
        cudaError_t e__ = cudaGetLastError();       \
        if (e__ != cudaSuccess) return (int)e__;    \
    } while (0)

static inline int rsdf_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// B200: 148 SMs.  Grid-stride kernels size their grids as a multiple of this.
#define RSDF_NUM_SMS 148

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// full-precision sigmoid (parity-critical paths)
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
