// Shared helpers for the glam_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/glam_b200.h"

namespace glam {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GLAM_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            glam::set_error(__VA_ARGS__);       \
            return -1;                          \
        }                                       \
    } while (0)

#define GLAM_CHECK_LAUNCH()                                                     \
    do {                                                                        \
        cudaError_t e__ = cudaGetLastError();                                   \
        if (e__ != cudaSuccess) {                                               \
            glam::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return (int)e__;                                                    \
        }                                                                       \
        glam::count_launch();                                                   \
    } while (0)

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
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// CELU(alpha = 1) = max(0,x) + min(0, exp(x) - 1).  Branch-free expm1 for x <= 0: degree-7 Taylor polynomial on
// [-0.25, 0] (truncation error < 4e-10), exp(x) - 1 below (cancellation error <= 1 ulp of 1 * 2^-2 relative there).
__device__ __forceinline__ float expm1_neg(float x) {
    const float p = x * (1.f + x * (0.5f + x * (1.f / 6 + x * (1.f / 24 + x * (1.f / 120 + x * (1.f / 720 + x * (1.f / 5040)))))));
    return x > -0.25f ? p : expf(x) - 1.f;
}
__device__ __forceinline__ float celu1(float x) { return x > 0.f ? x : expm1_neg(x); }

// activation codes shared with the host side (glam_b200/functional.py)
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_CELU = 3 };
__device__ __forceinline__ float act_fwd(float x, int act, float p) {
    switch (act) {
        case ACT_RELU: return x > 0.f ? x : 0.f;
        case ACT_LEAKY: return x > 0.f ? x : p * x;
        case ACT_CELU: return celu1(x);
        default: return x;
    }
}
// derivative expressed with the activation OUTPUT y (all four are sign-preserving)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float p) {
    switch (act) {
        case ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case ACT_LEAKY: return y > 0.f ? 1.f : p;
        case ACT_CELU: return y > 0.f ? 1.f : y + 1.f;
        default: return 1.f;
    }
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// out = sum over S partials (each `stride` floats, partial s at partial + s*stride) in a fixed order: block = 32 outputs x 8
// slices; a thread adds its slices' partials in ascending s (4 loads in flight), the 8 slices are combined in order.
// Element i < rows*cols goes to out[r*ldo + c] (or transposed); element i >= rows*cols goes to out2[i - rows*cols].
__global__ void __launch_bounds__(256)
reduce_partials_fixed_kernel(const float* __restrict__ partial, int S, int rows, int cols, int extra, float* __restrict__ out,
                             int64_t ldo, int transpose_out, float* __restrict__ out2);
// Raise a kernel's dynamic shared-memory limit.  The attribute is per (function, device) and sticky, so it is only ever
// raised: lowering it for a later, smaller launch would break tools that re-launch the nodes of an instantiated CUDA
// graph on their own (ncu) — they pick up the function's CURRENT limit, not the one the node was captured with.
cudaError_t ensure_dyn_smem(const void* fn, size_t bytes);
int launch_reduce_partials(const float* partial, int S, int rows, int cols, int extra, float* out, int64_t ldo, int transpose_out,
                           float* out2, cudaStream_t stream);

}  // namespace glam
