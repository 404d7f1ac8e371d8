// Per-graph attention pooling shared by GlobalLAPool (PyG GlobalAttention) and Set2Set
// (see include/glam_b200.h (5)).  One warp per graph; a graph's node rows are contiguous (PyG batches are
// block-diagonal with sorted `batch`), so every pass streams x[n0:n1, :] coalesced; the softmax is the PyG
// form exp(e - max) / (sum + 1e-16); all sums run in node order (deterministic).
#include "common.cuh"

namespace glam {

constexpr int kPoolWarps = 8;

__global__ void __launch_bounds__(kPoolWarps * 32)
seg_attn_pool_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ q, int64_t q_stride,
                         const float* __restrict__ q_bias, const int32_t* __restrict__ gptr, int64_t B, int C,
                         float* __restrict__ a, float* __restrict__ r, int64_t ldr, float* __restrict__ asum) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float qb = q_bias ? q_bias[0] : 0.f;
    for (int64_t g = warp0; g < B; g += nwarps) {
        const int n0 = gptr[g], n1 = gptr[g + 1];
        const float* qg = q + g * q_stride;
        // e[n] = <x[n], q> + b
        float mx = -INFINITY;
        for (int n = n0; n < n1; ++n) {
            const float* xn = x + (int64_t)n * ldx;
            float s = 0.f;
            for (int k = lane; k < C; k += 32) s = fmaf(xn[k], qg[k], s);
            s = warp_sum(s) + qb;
            mx = fmaxf(mx, s);
            if (lane == 0) a[n] = s;
        }
        __syncwarp();
        float sum = 0.f;
        for (int n = n0 + lane; n < n1; n += 32) {
            float e = expf(a[n] - mx);
            a[n] = e;
            sum += e;
        }
        sum = warp_sum(sum) + 1e-16f;
        float tot = 0.f;
        for (int n = n0 + lane; n < n1; n += 32) {
            float v = a[n] / sum;
            a[n] = v;
            tot += v;
        }
        tot = warp_sum(tot);
        __syncwarp();
        for (int k = lane; k < C; k += 32) {
            float acc = 0.f;
            for (int n = n0; n < n1; ++n) acc = fmaf(a[n], x[(int64_t)n * ldx + k], acc);
            r[g * ldr + k] = acc;
        }
        if (lane == 0 && asum) asum[g] = tot;
    }
}

__global__ void __launch_bounds__(kPoolWarps * 32)
seg_attn_pool_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ q, int64_t q_stride,
                         const float* __restrict__ a, const float* __restrict__ g_r, int64_t ldgr,
                         const float* __restrict__ g_asum, const int32_t* __restrict__ gptr, int64_t B, int C,
                         int accumulate, float* __restrict__ g_x, int64_t ldgx, float* __restrict__ g_q,
                         float* __restrict__ g_e) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp0; g < B; g += nwarps) {
        const int n0 = gptr[g], n1 = gptr[g + 1];
        const float* qg = q + g * q_stride;
        const float* gr = g_r + g * ldgr;
        const float gas = g_asum ? g_asum[g] : 0.f;
        float dot = 0.f;
        for (int n = n0; n < n1; ++n) {
            const float* xn = x + (int64_t)n * ldx;
            float s = 0.f;
            for (int k = lane; k < C; k += 32) s = fmaf(xn[k], gr[k], s);
            s = warp_sum(s) + gas;
            dot = fmaf(a[n], s, dot);
            if (lane == 0) g_e[n] = s;
        }
        __syncwarp();
        for (int n = n0 + lane; n < n1; n += 32) g_e[n] = a[n] * (g_e[n] - dot);
        __syncwarp();
        for (int k = lane; k < C; k += 32) {
            const float grk = gr[k], qk = qg[k];
            float gq = 0.f;
            for (int n = n0; n < n1; ++n) {
                const float ge = g_e[n];
                float v = fmaf(a[n], grk, ge * qk);
                float* dst = g_x + (int64_t)n * ldgx + k;
                *dst = accumulate ? *dst + v : v;
                gq = fmaf(ge, x[(int64_t)n * ldx + k], gq);
            }
            g_q[g * C + k] = gq;
        }
    }
}

static int pool_grid(int64_t B) {
    int64_t g = (B + kPoolWarps - 1) / kPoolWarps;
    int64_t cap = (int64_t)kNumSMs * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace glam

using namespace glam;

extern "C" int glam_seg_attn_pool_fwd(const float* x, int64_t ldx, const float* q, int64_t q_stride, const float* q_bias,
                                      const int32_t* graph_ptr, int64_t B, int C, float* a, float* r, int64_t ldr,
                                      float* asum, void* stream_) {
    GLAM_REQUIRE(B >= 0 && C > 0 && ldx >= C && ldr >= C, "glam_seg_attn_pool_fwd: bad shape");
    if (B == 0) return 0;
    GLAM_REQUIRE(x && q && graph_ptr && a && r, "glam_seg_attn_pool_fwd: null pointer");
    seg_attn_pool_fwd_kernel<<<pool_grid(B), kPoolWarps * 32, 0, (cudaStream_t)stream_>>>(x, ldx, q, q_stride, q_bias, graph_ptr, B, C,
                                                                                        a, r, ldr, asum);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_seg_attn_pool_bwd(const float* x, int64_t ldx, const float* q, int64_t q_stride, const float* a,
                                      const float* g_r, int64_t ldgr, const float* g_asum, const int32_t* graph_ptr,
                                      int64_t B, int C, int accumulate, float* g_x, int64_t ldgx, float* g_q, float* g_e,
                                      void* stream_) {
    GLAM_REQUIRE(B >= 0 && C > 0 && ldx >= C && ldgr >= C && ldgx >= C, "glam_seg_attn_pool_bwd: bad shape");
    if (B == 0) return 0;
    GLAM_REQUIRE(x && q && a && g_r && graph_ptr && g_x && g_q && g_e, "glam_seg_attn_pool_bwd: null pointer");
    seg_attn_pool_bwd_kernel<<<pool_grid(B), kPoolWarps * 32, 0, (cudaStream_t)stream_>>>(x, ldx, q, q_stride, a, g_r, ldgr, g_asum,
                                                                                        graph_ptr, B, C, accumulate, g_x, ldgx,
                                                                                        g_q, g_e);
    GLAM_CHECK_LAUNCH();
    return 0;
}
