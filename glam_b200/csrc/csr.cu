// Destination-/source-sorted CSR builder (see include/glam_b200.h (1)).
//
// Stable counting sort without floating point and without order-dependent results:
//   1. histogram of keys (integer atomics: result independent of arrival order)
//   2. exclusive scan -> rowptr
//   3. bucket fill through per-node cursors (order inside a bucket is arbitrary here)
//   4. each bucket is rank-sorted by ORIGINAL edge id by one warp -> exactly the order of a stable sort,
//      i.e. bit-identical to torch.argsort(key, stable=True) and run-to-run deterministic.
// blockIdx.y selects the key row: 0 = destinations (edge_index[1]), 1 = sources (edge_index[0]).
#include "common.cuh"

namespace glam {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanBlock * kScanItems;

struct CsrWs {
    int32_t* counts;   // [2][N+1]   histogram, then exclusive scan
    int32_t* cursor;   // [2][N+1]
    int32_t* bsum;     // [2][nblk]
    int32_t* tmp;      // [2][E]     unsorted buckets
    int32_t* inv_dst;  // [E]        original edge id -> position in dst order
    int32_t* src_perm; // [E]
};

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

static size_t carve(CsrWs& w, char* base, int64_t N, int64_t E) {
    size_t off = 0;
    int64_t nblk = (N + 1 + kScanTile - 1) / kScanTile;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
    w.counts = (int32_t*)take(sizeof(int32_t) * 2 * (N + 1));
    w.cursor = (int32_t*)take(sizeof(int32_t) * 2 * (N + 1));
    w.bsum = (int32_t*)take(sizeof(int32_t) * 2 * nblk);
    w.tmp = (int32_t*)take(sizeof(int32_t) * 2 * (E > 0 ? E : 1));
    w.inv_dst = (int32_t*)take(sizeof(int32_t) * (E > 0 ? E : 1));
    w.src_perm = (int32_t*)take(sizeof(int32_t) * (E > 0 ? E : 1));
    return off;
}

__global__ void csr_hist_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N, int32_t* counts) {
    const int which = blockIdx.y;                          // 0: dst row (row 1), 1: src row (row 0)
    const int64_t* key = edge_index + (which == 0 ? E : 0);
    int32_t* c = counts + (int64_t)which * (N + 1);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = key[e];
        if (k >= 0 && k < N) atomicAdd(&c[k], 1);
    }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    // 256-thread exclusive scan via warp shuffles
    __shared__ int wsum[kScanBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = lane < kScanBlock / 32 ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < kScanBlock / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < kScanBlock / 32) wsum[lane] = s;
    }
    __syncthreads();
    int base = wid > 0 ? wsum[wid - 1] : 0;
    *total = wsum[kScanBlock / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void csr_scan_blocksum_kernel(const int32_t* __restrict__ counts, int64_t n, int64_t nblk, int32_t* bsum) {
    const int which = blockIdx.y;
    const int32_t* c = counts + (int64_t)which * n;
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + threadIdx.x * kScanItems + i;
        if (idx < n) s += c[idx];
    }
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) bsum[(int64_t)which * nblk + blockIdx.x] = total;
}

__global__ void csr_scan_top_kernel(int32_t* bsum, int64_t nblk) {
    int32_t* b = bsum + (int64_t)blockIdx.y * nblk;
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblk; base += kScanBlock) {
        int64_t idx = base + threadIdx.x;
        int v = idx < nblk ? b[idx] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        int carry = carry_s;
        if (idx < nblk) b[idx] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void csr_scan_apply_kernel(int32_t* counts, int64_t n, int64_t nblk, const int32_t* __restrict__ bsum,
                                      int32_t* out_dst, int32_t* out_src) {
    const int which = blockIdx.y;
    int32_t* c = counts + (int64_t)which * n;
    int32_t* out = which == 0 ? out_dst : out_src;
    int64_t base = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? c[base + i] : 0;
        s += v[i];
    }
    int total;
    int ex = block_exclusive_scan(s, &total) + bsum[(int64_t)which * nblk + blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) { c[base + i] = ex; out[base + i] = ex; }
        ex += v[i];
    }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N,
                                const int32_t* __restrict__ rowptr2, int32_t* cursor, int32_t* tmp) {
    const int which = blockIdx.y;
    const int64_t* key = edge_index + (which == 0 ? E : 0);
    const int32_t* rp = rowptr2 + (int64_t)which * (N + 1);
    int32_t* cur = cursor + (int64_t)which * (N + 1);
    int32_t* t = tmp + (int64_t)which * E;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = key[e];
        if (k >= 0 && k < N) {
            int pos = atomicAdd(&cur[k], 1);
            t[rp[k] + pos] = (int32_t)e;
        }
    }
}

// Rank-sort every node's bucket by edge id (ids are unique) and emit the final arrays.  Molecular graphs have degree <= 4-6:
// one THREAD per node sorts up to 8 ids in registers (a warp per node left 30 lanes idle: 29 us per build at the bench
// shape); buckets up to 64 are ranked serially by their thread, larger ones (hub nodes) by the whole warp afterwards.
__global__ void csr_sort_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N,
                                const int32_t* __restrict__ rowptr2, const int32_t* __restrict__ tmp,
                                int32_t* dst_perm, int32_t* dst_src, int32_t* dst_dst, int32_t* inv_dst, int32_t* src_perm) {
    const int which = blockIdx.y;
    const int32_t* rp = rowptr2 + (int64_t)which * (N + 1);
    const int32_t* t = tmp + (int64_t)which * E;
    const int lane = threadIdx.x & 31;
    auto emit = [&](int64_t node, int p, int v) {
        if (which == 0) { dst_perm[p] = v; dst_src[p] = (int32_t)edge_index[v]; inv_dst[v] = p; if (dst_dst) dst_dst[p] = (int32_t)node; }
        else src_perm[p] = v;
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x - lane; i0 < N; i0 += stride) {
        const int64_t i = i0 + lane;
        int beg = 0, deg = 0;
        if (i < N) { beg = rp[i]; deg = rp[i + 1] - beg; }
        if (deg > 0 && deg <= 8) {
            int v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = k < deg ? t[beg + k] : 0x7fffffff;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < deg) {
                    int rank = 0;
#pragma unroll
                    for (int m = 0; m < 8; ++m) rank += (v[m] < v[k]);
                    emit(i, beg + rank, v[k]);
                }
            }
        } else if (deg > 8 && deg <= 64) {
            for (int a = 0; a < deg; ++a) {
                const int va = t[beg + a];
                int rank = 0;
                for (int k = 0; k < deg; ++k) rank += (t[beg + k] < va);
                emit(i, beg + rank, va);
            }
        }
        unsigned heavy = __ballot_sync(0xffffffffu, deg > 64);
        while (heavy) {                                                     // hub nodes: the warp ranks one bucket together
            const int src_lane = __ffs(heavy) - 1;
            heavy &= heavy - 1;
            const int hb = __shfl_sync(0xffffffffu, beg, src_lane), hd = __shfl_sync(0xffffffffu, deg, src_lane);
            for (int a = lane; a < hd; a += 32) {
                const int va = t[hb + a];
                int rank = 0;
                for (int k = 0; k < hd; ++k) rank += (t[hb + k] < va);
                emit(i0 + src_lane, hb + rank, va);
            }
        }
    }
}

__global__ void csr_src_finish_kernel(const int64_t* __restrict__ edge_index, int64_t E,
                                      const int32_t* __restrict__ src_perm, const int32_t* __restrict__ inv_dst,
                                      int32_t* src_pos, int32_t* src_dst) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < E; k += (int64_t)gridDim.x * blockDim.x) {
        int e = src_perm[k];
        src_pos[k] = inv_dst[e];
        src_dst[k] = (int32_t)edge_index[E + e];
    }
}

__global__ void graph_ptr_kernel(const int64_t* __restrict__ batch, int64_t N, int64_t B, int32_t* gptr) {
    // batch is non-decreasing: graph g starts where batch[n-1] < g <= batch[n]; empty graphs get empty ranges.
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n <= N; n += (int64_t)gridDim.x * blockDim.x) {
        int64_t prev = n == 0 ? -1 : batch[n - 1];
        int64_t cur = n == N ? B : batch[n];
        if (cur > B) cur = B;
        for (int64_t g = prev + 1; g <= cur; ++g) gptr[g] = (int32_t)n;
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ perm, int64_t rows,
                                   int64_t cols, float* __restrict__ out) {
    int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cols, c = i - r * cols;
        out[i] = in[(int64_t)perm[r] * cols + c];
    }
}

static int grid_for(int64_t n, int block, int per_sm = 8) {
    int64_t g = (n + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace glam

using namespace glam;

extern "C" size_t glam_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
    CsrWs w;
    return carve(w, nullptr, num_nodes, num_edges);
}

extern "C" int glam_build_csr(const int64_t* edge_index, int64_t E, int64_t N, int32_t* dst_rowptr, int32_t* dst_src,
                              int32_t* dst_perm, int32_t* dst_dst, int32_t* src_rowptr, int32_t* src_pos, int32_t* src_dst,
                              void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLAM_REQUIRE(N >= 0 && E >= 0, "glam_build_csr: negative sizes");
    GLAM_REQUIRE(N < (int64_t)1 << 31 && E < (int64_t)1 << 31, "glam_build_csr: N and E must fit int32");
    GLAM_REQUIRE(dst_rowptr && src_rowptr, "glam_build_csr: null rowptr");
    GLAM_REQUIRE(E == 0 || (edge_index && dst_src && dst_perm && src_pos && src_dst), "glam_build_csr: null pointer");
    CsrWs w;
    size_t need = carve(w, (char*)workspace, N, E);
    GLAM_REQUIRE(workspace && workspace_bytes >= need, "glam_build_csr: workspace too small (%zu < %zu)", workspace_bytes, need);
    const int64_t n1 = N + 1;
    const int64_t nblk = (n1 + kScanTile - 1) / kScanTile;
    cudaMemsetAsync(w.counts, 0, sizeof(int32_t) * 2 * n1, stream);
    cudaMemsetAsync(w.cursor, 0, sizeof(int32_t) * 2 * n1, stream);
    if (E > 0) {
        csr_hist_kernel<<<dim3(grid_for(E, 256), 2), 256, 0, stream>>>(edge_index, E, N, w.counts);
        GLAM_CHECK_LAUNCH();
    }
    csr_scan_blocksum_kernel<<<dim3((unsigned)nblk, 2), kScanBlock, 0, stream>>>(w.counts, n1, nblk, w.bsum);
    GLAM_CHECK_LAUNCH();
    csr_scan_top_kernel<<<dim3(1, 2), kScanBlock, 0, stream>>>(w.bsum, nblk);
    GLAM_CHECK_LAUNCH();
    csr_scan_apply_kernel<<<dim3((unsigned)nblk, 2), kScanBlock, 0, stream>>>(w.counts, n1, nblk, w.bsum, dst_rowptr, src_rowptr);
    GLAM_CHECK_LAUNCH();
    if (E > 0) {
        csr_fill_kernel<<<dim3(grid_for(E, 256), 2), 256, 0, stream>>>(edge_index, E, N, w.counts, w.cursor, w.tmp);
        GLAM_CHECK_LAUNCH();
        csr_sort_kernel<<<dim3(grid_for(N, 128), 2), 128, 0, stream>>>(edge_index, E, N, w.counts, w.tmp, dst_perm, dst_src,
                                                                          dst_dst, w.inv_dst, w.src_perm);
        GLAM_CHECK_LAUNCH();
        csr_src_finish_kernel<<<grid_for(E, 256), 256, 0, stream>>>(edge_index, E, w.src_perm, w.inv_dst, src_pos, src_dst);
        GLAM_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int glam_graph_ptr(const int64_t* batch, int64_t N, int64_t B, int32_t* graph_ptr, void* stream_) {
    GLAM_REQUIRE(N >= 0 && B >= 0 && graph_ptr && (N == 0 || batch), "glam_graph_ptr: bad arguments");
    GLAM_REQUIRE(N < (int64_t)1 << 31, "glam_graph_ptr: N must fit int32");
    graph_ptr_kernel<<<grid_for(N + 1, 256), 256, 0, (cudaStream_t)stream_>>>(batch, N, B, graph_ptr);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_gather_rows(const float* in, const int32_t* perm, int64_t rows, int64_t cols, float* out, void* stream_) {
    GLAM_REQUIRE(rows >= 0 && cols >= 0, "glam_gather_rows: negative sizes");
    if (rows * cols == 0) return 0;
    GLAM_REQUIRE(in && perm && out, "glam_gather_rows: null pointer");
    gather_rows_kernel<<<grid_for(rows * cols, 256), 256, 0, (cudaStream_t)stream_>>>(in, perm, rows, cols, out);
    GLAM_CHECK_LAUNCH();
    return 0;
}
