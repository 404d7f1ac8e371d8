// Library-level entry points: version, error string, launch counter.
#include <map>
#include <mutex>
#include <utility>
#include <vector>
#include <cstdint>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace glam {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace glam

namespace glam {
__global__ void __launch_bounds__(256)
reduce_partials_fixed_kernel(const float* __restrict__ partial, int S, int rows, int cols, int extra, float* __restrict__ out,
                             int64_t ldo, int transpose_out, float* __restrict__ out2) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int total = rows * cols, stride = total + extra;
    const int i = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (i < stride) {
        const int per = (S + 7) / 8, k0 = ty * per, k1 = min(S, k0 + per);
        int k = k0;
        for (; k + 4 <= k1; k += 4) {
            const float a = partial[(int64_t)k * stride + i], b = partial[(int64_t)(k + 1) * stride + i];
            const float c = partial[(int64_t)(k + 2) * stride + i], d = partial[(int64_t)(k + 3) * stride + i];
            s = (((s + a) + b) + c) + d;
        }
        for (; k < k1; ++k) s += partial[(int64_t)k * stride + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < stride) {
        float v = red[0][tx];
#pragma unroll
        for (int q = 1; q < 8; ++q) v += red[q][tx];
        if (i >= total) { out2[i - total] = v; return; }
        const int r = i / cols, c = i - r * cols;
        if (transpose_out) out[(int64_t)c * ldo + r] = v; else out[(int64_t)r * ldo + c] = v;
    }
}
cudaError_t ensure_dyn_smem(const void* fn, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> have;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = have[{fn, dev}];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}
int launch_reduce_partials(const float* partial, int S, int rows, int cols, int extra, float* out, int64_t ldo, int transpose_out,
                           float* out2, cudaStream_t stream) {
    const int n = rows * cols + extra;
    reduce_partials_fixed_kernel<<<(n + 31) / 32, 256, 0, stream>>>(partial, S, rows, cols, extra, out, ldo, transpose_out, out2);
    GLAM_CHECK_LAUNCH();
    return 0;
}
}  // namespace glam

// Host-side (no GPU involved): order of a batch's graphs in which consecutive graphs fill the fused kernels' tiles — first-fit
// decreasing bin packing with the sizes bucketed (<= cap_nodes buckets, a bitmap of the non-empty ones), O(B) — see
// glam_b200/synth.py::tile_order, whose numpy form this reproduces exactly (same tie-breaking: within a size the highest
// graph id goes first; empty graphs first, graphs over the node cap last).
extern "C" int glam_tile_order(const int64_t* nodes, const int64_t* edges, int64_t B, int cap_nodes, int cap_edges, int64_t* perm) {
    GLAM_REQUIRE(B >= 0 && cap_nodes >= 1 && cap_nodes <= 1024 && cap_edges >= 0, "glam_tile_order: bad arguments");
    if (B == 0) return 0;
    GLAM_REQUIRE(nodes && perm, "glam_tile_order: null pointer");
    const int S = cap_nodes + 1;
    std::vector<int64_t> start(S + 1, 0), top(S, 0), ids((size_t)B);
    int64_t n_big = 0;
    for (int64_t g = 0; g < B; ++g) {
        GLAM_REQUIRE(nodes[g] >= 0, "glam_tile_order: negative size");
        if (nodes[g] > cap_nodes || (edges && edges[g] > cap_edges)) ++n_big; else ++start[nodes[g] + 1];
    }
    for (int s = 0; s < S; ++s) start[s + 1] += start[s];
    for (int s = 0; s < S; ++s) top[s] = start[s];
    std::vector<int64_t> bigs; bigs.reserve((size_t)n_big);
    for (int64_t g = 0; g < B; ++g) {                     // increasing graph id inside a size: the stack pops the highest first
        if (nodes[g] > cap_nodes || (edges && edges[g] > cap_edges)) bigs.push_back(g); else ids[(size_t)top[nodes[g]]++] = g;
    }
    int64_t out = 0;
    for (int64_t i = start[0]; i < top[0]; ++i) perm[out++] = ids[(size_t)i];      // empty graphs take no rows
    top[0] = start[0];
    std::vector<uint64_t> bits((size_t)(S + 63) / 64, 0);
    int64_t remaining = 0;
    for (int s = 1; s < S; ++s) if (top[s] > start[s]) { bits[s >> 6] |= 1ull << (s & 63); remaining += top[s] - start[s]; }
    auto highest_at_most = [&](int s) -> int {            // largest non-empty size <= s, or 0
        int w = s >> 6;
        uint64_t m = bits[w] & (((s & 63) == 63) ? ~0ull : ((1ull << ((s & 63) + 1)) - 1));
        while (true) {
            if (m) return (w << 6) + 63 - __builtin_clzll(m);
            if (--w < 0) return 0;
            m = bits[w];
        }
    };
    while (remaining) {
        int room = cap_nodes;
        int64_t eroom = cap_edges;
        while (true) {
            int s = highest_at_most(room);
            bool placed = false;
            while (s > 0) {
                const int64_t g = ids[(size_t)top[s] - 1];
                const int64_t eg = edges ? edges[g] : 0;
                if (eg <= eroom) {
                    --top[s]; --remaining;
                    if (top[s] == start[s]) bits[s >> 6] &= ~(1ull << (s & 63));
                    perm[out++] = g; room -= s; eroom -= eg; placed = true;
                    break;
                }
                s = s > 1 ? highest_at_most(s - 1) : 0;
            }
            if (!placed) break;
        }
    }
    for (int64_t g : bigs) perm[out++] = g;
    return 0;
}

extern "C" int glam_abi_version(void) { return GLAM_B200_ABI_VERSION; }
extern "C" const char* glam_last_error(void) { return glam::g_err; }
extern "C" int64_t glam_launch_count(void) { return glam::g_launches.load(std::memory_order_relaxed); }
