// PairNorm per graph (PyG norm/pair_norm.py @1.7.2, scale = 1, scale_individually = False; the reference's default
// `graph_norm`, src_1gp/run.py:28, applied to the block input, src_1gp/layer.py:255):
//     xc = x - mean_n(x)            (per graph, per channel)
//     y  = xc / sqrt(eps + mean_n(sum_c xc^2))
// The torch formulation (index_add_ on CUDA = floating-point atomics, scatter order) is not reproducible run to run;
// here one warp owns a graph and every reduction is a fixed-order loop + shuffle tree.  Two passes over the graph's rows
// (they are L1/L2 resident: ~25 rows x C floats), no intermediate tensor.
//     backward:  g_xc = s g_y - (s^3 / n) xc <g_y, xc>_graph ;   g_x = g_xc - mean_n(g_xc),   s = 1 / sqrt(eps + q / n)
#include "common.cuh"

namespace glam {
namespace {

constexpr int kPnWarps = 8;
constexpr int kPnMaxCPL = 4;          // channels per lane: C <= 128

template <bool BWD>
__global__ void __launch_bounds__(kPnWarps * 32)
pair_norm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g_y, int64_t ldg, const int32_t* __restrict__ gptr,
                 int64_t B, int C, float eps, float* __restrict__ out, int64_t ldo, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= B) return;
    const int n0 = gptr[g], n1 = gptr[g + 1], n = n1 - n0;
    if (n <= 0) return;
    const float inv_n = 1.f / (float)n;
    float mean[kPnMaxCPL];
#pragma unroll
    for (int k = 0; k < kPnMaxCPL; ++k) {
        const int c = lane + 32 * k;
        float s = 0.f;
        if (c < C)
            for (int i = n0; i < n1; ++i) s += x[(int64_t)i * ldx + c];
        mean[k] = s * inv_n;
    }
    float q = 0.f;                                        // sum over rows and this lane's channels of xc^2
    for (int i = n0; i < n1; ++i)
#pragma unroll
        for (int k = 0; k < kPnMaxCPL; ++k) {
            const int c = lane + 32 * k;
            if (c < C) { const float d = x[(int64_t)i * ldx + c] - mean[k]; q = fmaf(d, d, q); }
        }
    q = warp_sum(q);
    const float s = 1.f / sqrtf(eps + q * inv_n);
    if (!BWD) {
        for (int i = n0; i < n1; ++i)
#pragma unroll
            for (int k = 0; k < kPnMaxCPL; ++k) {
                const int c = lane + 32 * k;
                if (c < C) out[(int64_t)i * ldo + c] = (x[(int64_t)i * ldx + c] - mean[k]) * s;
            }
        return;
    }
    float dot = 0.f, gsum[kPnMaxCPL];                     // <g_y, xc> over the graph; column sums of g_y
#pragma unroll
    for (int k = 0; k < kPnMaxCPL; ++k) gsum[k] = 0.f;
    for (int i = n0; i < n1; ++i)
#pragma unroll
        for (int k = 0; k < kPnMaxCPL; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                const float gy = g_y[(int64_t)i * ldg + c];
                dot = fmaf(gy, x[(int64_t)i * ldx + c] - mean[k], dot);
                gsum[k] += gy;
            }
        }
    dot = warp_sum(dot);
    const float coef = s * s * s * inv_n * dot;
    // mean_n(g_xc)[c] = s * mean_n(g_y)[c] - coef * mean_n(xc)[c] = s * gsum[c] / n   (xc has zero column means)
    for (int i = n0; i < n1; ++i)
#pragma unroll
        for (int k = 0; k < kPnMaxCPL; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                const float xc = x[(int64_t)i * ldx + c] - mean[k];
                const float v = s * g_y[(int64_t)i * ldg + c] - coef * xc - s * gsum[k] * inv_n;
                float* o = out + (int64_t)i * ldo + c;
                *o = accumulate ? *o + v : v;
            }
        }
}

}  // namespace
}  // namespace glam

using namespace glam;

extern "C" int glam_pair_norm_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels, float eps,
                                  float* y, int64_t ldy, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && channels > 0 && channels <= 32 * kPnMaxCPL, "glam_pair_norm_fwd: channels must be in [1, 128]");
    if (num_graphs == 0) return 0;
    GLAM_REQUIRE(x && graph_ptr && y && ldx >= channels && ldy >= channels, "glam_pair_norm_fwd: bad arguments");
    pair_norm_kernel<false><<<(unsigned)((num_graphs * 32 + kPnWarps * 32 - 1) / (kPnWarps * 32)), kPnWarps * 32, 0, (cudaStream_t)stream_>>>(
        x, ldx, nullptr, 0, graph_ptr, num_graphs, channels, eps, y, ldy, 0);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_pair_norm_bwd(const float* x, int64_t ldx, const float* g_y, int64_t ldg, const int32_t* graph_ptr, int64_t num_graphs,
                                  int channels, float eps, float* g_x, int64_t ldgx, int accumulate, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && channels > 0 && channels <= 32 * kPnMaxCPL, "glam_pair_norm_bwd: channels must be in [1, 128]");
    if (num_graphs == 0) return 0;
    GLAM_REQUIRE(x && g_y && graph_ptr && g_x && ldx >= channels && ldg >= channels && ldgx >= channels, "glam_pair_norm_bwd: bad arguments");
    pair_norm_kernel<true><<<(unsigned)((num_graphs * 32 + kPnWarps * 32 - 1) / (kPnWarps * 32)), kPnWarps * 32, 0, (cudaStream_t)stream_>>>(
        x, ldx, g_y, ldg, graph_ptr, num_graphs, channels, eps, g_x, ldgx, accumulate);
    GLAM_CHECK_LAUNCH();
    return 0;
}
