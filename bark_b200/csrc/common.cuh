// Shared helpers for the bark_b200 CUDA sources (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bark_b200.h"

namespace bark {

void set_last_error(const char* fmt, ...);

#define BARK_CHECK_ARG(cond, msg)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            bark::set_last_error("%s: %s", __func__, msg);          \
            return BARK_E_INVALID;                                  \
        }                                                           \
    } while (0)

#define BARK_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            bark::set_last_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__));  \
            return BARK_E_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define BARK_LAUNCH_CHECK() BARK_CUDA(cudaGetLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- feature types (FeatureTypeEnum, src/bark/forest.py:22-25) ----
constexpr int FEAT_CAT = 0;
constexpr int FEAT_INT = 1;
constexpr int FEAT_CONT = 2;

// Split test of src/bark/forest.py:37-41.  Numeric: left iff x <= threshold compared in f64 (the f32
// threshold is promoted).  Categorical: left iff bit int(x) of int(threshold) is set; int() truncates.
__device__ __forceinline__ bool goes_left(double x, float threshold, int feat_type) {
    if (feat_type == FEAT_CAT) {
        long long cat = (long long)x;
        long long mask = (long long)threshold;
        // numba: (1 << int(x)) & int(threshold) on int64
        unsigned long long bit = (cat >= 0 && cat < 64) ? (1ull << cat) : 0ull;
        return (bit & (unsigned long long)mask) != 0ull;
    }
    return x <= (double)threshold;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// two independent sums, interleaved step by step (the shuffle chains overlap instead of running back to back);
// each result is bit-identical to warp_sum of the same value
__device__ __forceinline__ void warp_sum2(double& a, double& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o);
        const double tb = __shfl_xor_sync(0xffffffffu, b, o);
        a += ta;
        b += tb;
    }
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed reduction tree).  `scratch` holds >= 32 doubles.  All threads get the sum.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect scratch from a previous use
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

}  // namespace bark
