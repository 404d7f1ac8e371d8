// GlobalPool5 readout (src_1gp/layer.py:197-203) — the reference's default `mol_readout` (src_1gp/run.py:25):
//   out[g] = [ mean_n x | sum_n x | sort-pool(k=3) ]       [B, 5C]
// PyG global_sort_pool(x, batch, k=3) @1.7.2: nodes of each graph sorted by their LAST channel, descending, first k rows
// kept in that order, missing rows zero.  Ties keep node order (lowest index first).
// One warp per graph; the three winners are found by three warp arg-max passes over the graph's last-channel column.
#include "common.cuh"

namespace glam {

__global__ void __launch_bounds__(256)
pool5_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ gptr, int64_t B, int C,
                 float* __restrict__ out, int32_t* __restrict__ top_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp0; g < B; g += nwarps) {
        const int n0 = gptr[g], n = gptr[g + 1] - n0;
        float* o = out + g * 5 * C;
        // mean | sum: lanes over channels, nodes in order
        const float inv = n > 0 ? 1.f / (float)n : 0.f;
        for (int c = lane; c < C; c += 32) {
            float s = 0.f;
            for (int i = 0; i < n; ++i) s += x[(int64_t)(n0 + i) * ldx + c];
            o[c] = s * inv;
            o[C + c] = s;
        }
        // top-3 by last channel
        int chosen[3] = {-1, -1, -1};
        for (int r = 0; r < 3; ++r) {
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = lane; i < n; i += 32) {
                if (i == chosen[0] || i == chosen[1]) continue;
                const float k = x[(int64_t)(n0 + i) * ldx + (C - 1)];
                if (bi == 0x7fffffff || k > best || (k == best && i < bi)) { best = k; bi = i; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
            }
            chosen[r] = bi == 0x7fffffff ? -1 : bi;
            float* t = o + (2 + r) * C;
            for (int c = lane; c < C; c += 32) t[c] = chosen[r] >= 0 ? x[(int64_t)(n0 + chosen[r]) * ldx + c] : 0.f;
            if (lane == 0) top_idx[g * 3 + r] = chosen[r] >= 0 ? n0 + chosen[r] : -1;
            if (chosen[r] < 0) {                                 // fewer than k nodes: the rest stays zero
                for (int rr = r + 1; rr < 3; ++rr) {
                    float* tt = o + (2 + rr) * C;
                    for (int c = lane; c < C; c += 32) tt[c] = 0.f;
                    if (lane == 0) top_idx[g * 3 + rr] = -1;
                }
                break;
            }
        }
    }
}

// g_x[n,:] = g_mean[g]/cnt + g_sum[g] + sum_r [n == top_r(g)] g_top[g,r,:]
__global__ void __launch_bounds__(256)
pool5_bwd_kernel(const float* __restrict__ g_out, const int32_t* __restrict__ gptr, const int32_t* __restrict__ top_idx, int64_t B,
                 int C, float* __restrict__ g_x, int64_t ldgx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp0; g < B; g += nwarps) {
        const int n0 = gptr[g], n = gptr[g + 1] - n0;
        const float inv = n > 0 ? 1.f / (float)n : 0.f;
        const float* go = g_out + g * 5 * C;
        const int t0 = top_idx[g * 3], t1 = top_idx[g * 3 + 1], t2 = top_idx[g * 3 + 2];
        for (int c = lane; c < C; c += 32) {
            const float base = go[c] * inv + go[C + c];
            for (int i = 0; i < n; ++i) {
                const int node = n0 + i;
                float v = base;
                if (node == t0) v += go[2 * C + c];
                if (node == t1) v += go[3 * C + c];
                if (node == t2) v += go[4 * C + c];
                g_x[(int64_t)node * ldgx + c] = v;
            }
        }
    }
}

}  // namespace glam

using namespace glam;

extern "C" int glam_pool5_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels, float* out,
                              int32_t* top_idx, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && channels > 0 && ldx >= channels, "glam_pool5_fwd: bad shape");
    if (num_graphs == 0) return 0;
    GLAM_REQUIRE(x && graph_ptr && out && top_idx, "glam_pool5_fwd: null pointer");
    int64_t blocks = (num_graphs + 7) / 8;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    pool5_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, ldx, graph_ptr, num_graphs, channels, out, top_idx);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_pool5_bwd(const float* g_out, const int32_t* graph_ptr, const int32_t* top_idx, int64_t num_graphs, int channels,
                              float* g_x, int64_t ldgx, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && channels > 0 && ldgx >= channels, "glam_pool5_bwd: bad shape");
    if (num_graphs == 0) return 0;
    GLAM_REQUIRE(g_out && graph_ptr && top_idx && g_x, "glam_pool5_bwd: null pointer");
    int64_t blocks = (num_graphs + 7) / 8;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    pool5_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(g_out, graph_ptr, top_idx, num_graphs, channels, g_x, ldgx);
    GLAM_CHECK_LAUNCH();
    return 0;
}
