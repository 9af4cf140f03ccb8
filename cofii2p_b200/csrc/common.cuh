// common.cuh -- shared helpers for libcofi_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cofi_b200.h"

namespace cofi {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return COFI_ECUDA;
    }
    count_launch();
    return COFI_OK;
}

#define COFI_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            cofi::set_error(__VA_ARGS__);  \
            return COFI_EINVAL;            \
        }                                  \
    } while (0)

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case COFI_ACT_RELU: return fmaxf(v, 0.0f);
        case COFI_ACT_LRELU01: return v > 0.0f ? v : v * 0.1f;
        case COFI_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace cofi
