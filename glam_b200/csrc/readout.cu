// Per-graph attention pooling shared by GlobalLAPool (PyG GlobalAttention) and Set2Set
// (see include/glam_b200.h (5)).  A graph's node rows are contiguous (PyG batches are block-diagonal with sorted
// `batch`): one warp owns one graph, copies its [n, C] block into a private shared-memory tile with coalesced
// 16-byte loads, and runs logits (lane = node), the PyG softmax exp(e - max) / (sum + 1e-16), and the weighted sum
// (lane = channel) from there.  Graphs larger than the tile (protein contact maps) stream from global memory.
// All sums run in node order: deterministic.
#include "common.cuh"

namespace glam {

constexpr int kPoolWarps = 4;
constexpr int kPoolMaxTileFloats = 4800;      // per-warp tile (19 KB): e.g. 128 nodes x (36+1) channels

__device__ __forceinline__ int pool_tile_nodes(int C) { return kPoolMaxTileFloats / (C + 1); }

// copy rows [n0, n1) x C into tile[n][C+1]
__device__ __forceinline__ void pool_load_tile(float* tile, const float* __restrict__ x, int64_t ldx, int n0, int n1, int C, int lane) {
    const int n = n1 - n0, ld = C + 1;
    if (ldx == C && (C & 3) == 0 && (((uintptr_t)(x + (int64_t)n0 * ldx)) & 15) == 0) {
        const float4* src = reinterpret_cast<const float4*>(x + (int64_t)n0 * ldx);
        const int total4 = n * C / 4, cq = C >> 2;
        for (int i = lane; i < total4; i += 32) {
            const float4 v = src[i];
            const int r = i / cq, c = (i - r * cq) << 2;
            float* d = tile + r * ld + c;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int r = 0; r < n; ++r)
            for (int c = lane; c < C; c += 32) tile[r * ld + c] = x[(int64_t)(n0 + r) * ldx + c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kPoolWarps * 32)
seg_attn_pool_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ q, int64_t q_stride,
                         const float* __restrict__ q_bias, const int32_t* __restrict__ gptr, int64_t B, int C,
                         float* __restrict__ a, float* __restrict__ r, int64_t ldr, float* __restrict__ asum) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* tile = smem + wid * kPoolMaxTileFloats;
    const int cap = pool_tile_nodes(C), ld = C + 1;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float qb = q_bias ? q_bias[0] : 0.f;
    for (int64_t g = warp0; g < B; g += nwarps) {
        const int n0 = gptr[g], n1 = gptr[g + 1], n = n1 - n0;
        const float* qg = q + g * q_stride;
        if (n <= cap) {
            pool_load_tile(tile, x, ldx, n0, n1, C, lane);
            // logits: lane = node
            float mx = -INFINITY;
            for (int nb = 0; nb < n; nb += 32) {
                const int i = nb + lane;
                float s = 0.f;
                if (i < n) {
                    const float* row = tile + i * ld;
                    for (int c = 0; c < C; ++c) s = fmaf(row[c], qg[c], s);
                    s += qb;
                    tile[i * ld + C] = s;                        // spare column holds the logit
                    mx = fmaxf(mx, s);
                }
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int i = lane; i < n; i += 32) {
                const float e = expf(tile[i * ld + C] - mx);
                tile[i * ld + C] = e;
                sum += e;
            }
            sum = warp_sum(sum) + 1e-16f;
            float tot = 0.f;
            for (int i = lane; i < n; i += 32) {
                const float v = tile[i * ld + C] / sum;
                tile[i * ld + C] = v;
                a[n0 + i] = v;
                tot += v;
            }
            tot = warp_sum(tot);
            __syncwarp();
            // weighted sum: lane = channel
            for (int c = lane; c < C; c += 32) {
                float acc = 0.f;
                for (int i = 0; i < n; ++i) acc = fmaf(tile[i * ld + C], tile[i * ld + c], acc);
                r[g * ldr + c] = acc;
            }
            if (lane == 0 && asum) asum[g] = tot;
            __syncwarp();
            continue;
        }
        // ---- large graph: stream from global memory
        float mx = -INFINITY;
        for (int nn = n0; nn < n1; ++nn) {
            const float* xn = x + (int64_t)nn * ldx;
            float s = 0.f;
            for (int k = lane; k < C; k += 32) s = fmaf(xn[k], qg[k], s);
            s = warp_sum(s) + qb;
            mx = fmaxf(mx, s);
            if (lane == 0) a[nn] = s;
        }
        __syncwarp();
        float sum = 0.f;
        for (int nn = n0 + lane; nn < n1; nn += 32) {
            float e = expf(a[nn] - mx);
            a[nn] = e;
            sum += e;
        }
        sum = warp_sum(sum) + 1e-16f;
        float tot = 0.f;
        for (int nn = n0 + lane; nn < n1; nn += 32) {
            float v = a[nn] / sum;
            a[nn] = v;
            tot += v;
        }
        tot = warp_sum(tot);
        __syncwarp();
        for (int k = lane; k < C; k += 32) {
            float acc = 0.f;
            for (int nn = n0; nn < n1; ++nn) acc = fmaf(a[nn], x[(int64_t)nn * ldx + k], acc);
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
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* tile = smem + wid * kPoolMaxTileFloats;
    const int cap = pool_tile_nodes(C), ld = C + 1;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t g = warp0; g < B; g += nwarps) {
        const int n0 = gptr[g], n1 = gptr[g + 1], n = n1 - n0;
        const float* qg = q + g * q_stride;
        const float* gr = g_r + g * ldgr;
        const float gas = g_asum ? g_asum[g] : 0.f;
        if (n <= cap) {
            pool_load_tile(tile, x, ldx, n0, n1, C, lane);
            // g_a[n] = <g_r, x[n]> + g_asum ; dot = sum a g_a      (lane = node)
            float dot = 0.f;
            for (int nb = 0; nb < n; nb += 32) {
                const int i = nb + lane;
                if (i < n) {
                    const float* row = tile + i * ld;
                    float s = 0.f;
                    for (int c = 0; c < C; ++c) s = fmaf(row[c], gr[c], s);
                    s += gas;
                    tile[i * ld + C] = s;
                    dot = fmaf(a[n0 + i], s, dot);
                }
            }
            dot = warp_sum(dot);
            for (int i = lane; i < n; i += 32) {
                const float ge = a[n0 + i] * (tile[i * ld + C] - dot);
                tile[i * ld + C] = ge;
                g_e[n0 + i] = ge;
            }
            __syncwarp();
            // g_x[n,c] = a[n] g_r[c] + g_e[n] q[c] ; g_q[c] = sum_n g_e[n] x[n,c]      (lane = channel)
            for (int c = lane; c < C; c += 32) {
                const float grc = gr[c], qc = qg[c];
                float gq = 0.f;
                for (int i = 0; i < n; ++i) {
                    const float ge = tile[i * ld + C];
                    const float v = fmaf(a[n0 + i], grc, ge * qc);
                    float* dst = g_x + (int64_t)(n0 + i) * ldgx + c;
                    *dst = accumulate ? *dst + v : v;
                    gq = fmaf(ge, tile[i * ld + c], gq);
                }
                g_q[g * C + c] = gq;
            }
            __syncwarp();
            continue;
        }
        // ---- large graph: stream from global memory
        float dot = 0.f;
        for (int nn = n0; nn < n1; ++nn) {
            const float* xn = x + (int64_t)nn * ldx;
            float s = 0.f;
            for (int k = lane; k < C; k += 32) s = fmaf(xn[k], gr[k], s);
            s = warp_sum(s) + gas;
            dot = fmaf(a[nn], s, dot);
            if (lane == 0) g_e[nn] = s;
        }
        __syncwarp();
        for (int nn = n0 + lane; nn < n1; nn += 32) g_e[nn] = a[nn] * (g_e[nn] - dot);
        __syncwarp();
        for (int k = lane; k < C; k += 32) {
            const float grk = gr[k], qk = qg[k];
            float gq = 0.f;
            for (int nn = n0; nn < n1; ++nn) {
                const float ge = g_e[nn];
                float v = fmaf(a[nn], grk, ge * qk);
                float* dst = g_x + (int64_t)nn * ldgx + k;
                *dst = accumulate ? *dst + v : v;
                gq = fmaf(ge, x[(int64_t)nn * ldx + k], gq);
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
constexpr size_t kPoolSmem = sizeof(float) * kPoolWarps * kPoolMaxTileFloats;   // 75 KB

}  // namespace glam

using namespace glam;

extern "C" int glam_seg_attn_pool_fwd(const float* x, int64_t ldx, const float* q, int64_t q_stride, const float* q_bias,
                                      const int32_t* graph_ptr, int64_t B, int C, float* a, float* r, int64_t ldr,
                                      float* asum, void* stream_) {
    GLAM_REQUIRE(B >= 0 && C > 0 && C <= 1024 && ldx >= C && ldr >= C, "glam_seg_attn_pool_fwd: bad shape");
    if (B == 0) return 0;
    GLAM_REQUIRE(x && q && graph_ptr && a && r, "glam_seg_attn_pool_fwd: null pointer");
    // (the attribute is per DEVICE: set on every call — a process-wide "configured" flag broke the second GPU of a process)
    ensure_dyn_smem((const void*)seg_attn_pool_fwd_kernel, (size_t)((int)kPoolSmem));
    seg_attn_pool_fwd_kernel<<<pool_grid(B), kPoolWarps * 32, kPoolSmem, (cudaStream_t)stream_>>>(x, ldx, q, q_stride, q_bias, graph_ptr,
                                                                                                B, C, a, r, ldr, asum);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_seg_attn_pool_bwd(const float* x, int64_t ldx, const float* q, int64_t q_stride, const float* a,
                                      const float* g_r, int64_t ldgr, const float* g_asum, const int32_t* graph_ptr,
                                      int64_t B, int C, int accumulate, float* g_x, int64_t ldgx, float* g_q, float* g_e,
                                      void* stream_) {
    GLAM_REQUIRE(B >= 0 && C > 0 && C <= 1024 && ldx >= C && ldgr >= C && ldgx >= C, "glam_seg_attn_pool_bwd: bad shape");
    if (B == 0) return 0;
    GLAM_REQUIRE(x && q && a && g_r && graph_ptr && g_x && g_q && g_e, "glam_seg_attn_pool_bwd: null pointer");
    ensure_dyn_smem((const void*)seg_attn_pool_bwd_kernel, (size_t)((int)kPoolSmem));
    seg_attn_pool_bwd_kernel<<<pool_grid(B), kPoolWarps * 32, kPoolSmem, (cudaStream_t)stream_>>>(x, ldx, q, q_stride, a, g_r, ldgr, g_asum,
                                                                                                graph_ptr, B, C, accumulate, g_x,
                                                                                                ldgx, g_q, g_e);
    GLAM_CHECK_LAUNCH();
    return 0;
}
