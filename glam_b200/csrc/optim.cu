// Flat Adam step for the data-parallel training step (reference: torch.optim.Adam(model.parameters(), lr=args.lr),
// src_1gp/trainer.py:49-50; defaults betas=(0.9, 0.999), eps=1e-8, no weight decay, no amsgrad).
//
// All parameters of a GLAM model are ~0.1-0.2 M floats in ~20 tensors.  torch's multi-tensor Adam is one
// launch but ~70 us on that shape; with parameters, gradients and both moments each in ONE flat fp32 buffer (the
// gradient buffer is the bucket the NCCL all-reduce already runs on) the step is a single streaming pass.  The step
// count and learning rate live in device memory so the launch is CUDA-graph replayable and an LR scheduler
// (ReduceLROnPlateau, trainer.py:55) can change lr without re-capturing.
#include "common.cuh"

namespace glam {

// state[0] = step (as float), state[1] = 1 - beta1^step, state[2] = 1 - beta2^step
__global__ void adam_tick_kernel(float* __restrict__ state, float beta1, float beta2) {
    const float t = state[0] + 1.f;
    state[0] = t;
    state[1] = 1.f - powf(beta1, t);
    state[2] = 1.f - powf(beta2, t);
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 const float* __restrict__ lr, const float* __restrict__ state, float beta1, float beta2, float eps,
                 float weight_decay, float grad_scale) {
    const float step_size = lr[0] / state[1];
    const float inv_sqrt_bc2 = rsqrtf(state[2]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const float denom = fmaf(sqrtf(vi), inv_sqrt_bc2, eps);
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace glam

using namespace glam;

extern "C" int glam_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const float* lr, float* state, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, void* stream_) {
    GLAM_REQUIRE(n >= 0 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "glam_adam_step: bad hyper-parameters");
    GLAM_REQUIRE(lr && state, "glam_adam_step: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    adam_tick_kernel<<<1, 1, 0, stream>>>(state, beta1, beta2);
    GLAM_CHECK_LAUNCH();
    if (n == 0) return 0;
    GLAM_REQUIRE(param && grad && exp_avg && exp_avg_sq, "glam_adam_step: null pointer");
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    adam_flat_kernel<<<(unsigned)blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, state, beta1, beta2, eps,
                                                          weight_decay, grad_scale);
    GLAM_CHECK_LAUNCH();
    return 0;
}
