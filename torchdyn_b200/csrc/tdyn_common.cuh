// Shared helpers for libtdyn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/tdyn_b200.h"

namespace tdyn {

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr int kMaxReduceBlocks = 148 * 8;

void set_error(const char* fmt, ...);

#define TDYN_CHECK_CUDA(expr)                                                          \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            tdyn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                            __FILE__, __LINE__);                                       \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

#define TDYN_REQUIRE(cond, ...)                                                        \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            tdyn::set_error(__VA_ARGS__);                                              \
            return 2;                                                                  \
        }                                                                              \
    } while (0)

struct KPtrs { const float* p[TDYN_MAX_STAGES]; };
struct Coefs { float c[TDYN_MAX_STAGES]; };

// One IEEE rounding per reference ATen op; never contracted into FMA.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// y = combine(x, k, coef) for one element; `da` = dt*coef (CHAIN) precomputed per thread.
template <int NK, int MODE>
__device__ __forceinline__ float combine1(float x, bool has_x, const float (&kv)[TDYN_MAX_STAGES],
                                          const Coefs& a, const Coefs& da, float dt) {
    if (MODE == TDYN_MODE_CHAIN) {
        float y = x;
#pragma unroll
        for (int j = 0; j < NK; ++j) y = add_rn(y, mul_rn(da.c[j], kv[j]));
        return y;
    } else {
        float s = mul_rn(a.c[0], kv[0]);
#pragma unroll
        for (int j = 1; j < NK; ++j) s = add_rn(s, mul_rn(a.c[j], kv[j]));
        s = mul_rn(dt, s);
        return has_x ? add_rn(x, s) : s;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-device kernel attributes (opt-in shared memory sizes) live in each device's context: true exactly once per
// (calling thread, current device) for the `seen` mask of one call site.
inline bool first_use_on_device(unsigned long long& seen) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}

}  // namespace tdyn
