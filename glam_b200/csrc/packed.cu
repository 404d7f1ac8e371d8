// Packed graph store -> device index: the screening input format.
//
// The reference ships every batch as fp32 features + int64 edge_index + fp32 one-hot edge_attr + int64 batch
// (src_1gp/dataset.py:60-97 -> Batch.from_data_list): 2.6 KB per 25-atom molecule, and the index (dst-sorted CSR, bond
// types, graph offsets) is re-derived from it on the device for every batch (0.57 ms of a 4.4 ms 64k-graph screening batch).
// A molecule's graph never changes, so the packed store keeps, per graph, what the kernels actually read, in the narrowest
// integer types that hold it, with the dst-sorted order fixed at pack time (glam_b200/packed.py):
//   n_g   uint8  [B]   atoms of graph g                 e_g  uint16 [B]   directed bonds (in-edges) of graph g
//   deg   uint8  [N]   in-degree of every atom          nbr  uint8  [E]   source atom of every in-edge, LOCAL to its graph
//   etype uint8  [E]   bond type of every in-edge       xq   uint8  [N,F] atom features (small non-negative integers)
// (~360 B per molecule).  Unpacking = two short scans + one warp per graph; bit-identical to glam_build_csr on the raw batch.
#include "common.cuh"

namespace glam {
namespace {

// exclusive scans of n_g and e_g -> gptr, eptr (B+1 entries each).  One CTA; each thread a contiguous run of graphs.
__global__ void __launch_bounds__(1024)
packed_scan_kernel(const uint8_t* __restrict__ n_g, const uint16_t* __restrict__ e_g, int64_t B, int32_t* __restrict__ gptr,
                   int32_t* __restrict__ eptr) {
    __shared__ int pn[1024], pe[1024];
    const int t = threadIdx.x;
    const int64_t per = (B + 1023) / 1024, b0 = min(B, t * per), b1 = min(B, b0 + per);
    int sn = 0, se = 0;
    for (int64_t b = b0; b < b1; ++b) { sn += n_g[b]; se += e_g[b]; }
    pn[t] = sn; pe[t] = se;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int vn = t >= o ? pn[t - o] : 0, ve = t >= o ? pe[t - o] : 0;
        __syncthreads();
        pn[t] += vn; pe[t] += ve;
        __syncthreads();
    }
    int rn = pn[t] - sn, re = pe[t] - se;
    for (int64_t b = b0; b < b1; ++b) { gptr[b] = rn; eptr[b] = re; rn += n_g[b]; re += e_g[b]; }
    if (t == 1023) { gptr[B] = pn[1023]; eptr[B] = pe[1023]; }
}

// a warp per graph: rowptr of its atoms (prefix of deg), global source ids of its in-edges, fp32 features
__global__ void __launch_bounds__(256)
packed_unpack_kernel(const int32_t* __restrict__ gptr, const int32_t* __restrict__ eptr, const uint8_t* __restrict__ deg,
                     const uint8_t* __restrict__ nbr, const uint8_t* __restrict__ xq, int64_t B, int F, int32_t* __restrict__ rowptr,
                     int32_t* __restrict__ dst_src, float* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= B) return;
    const int n0 = gptr[g], n1 = gptr[g + 1], e0 = eptr[g], e1 = eptr[g + 1];
    int carry = e0;
    for (int base = n0; base < n1; base += 32) {
        const int i = base + lane;
        const int d = i < n1 ? (int)deg[i] : 0;
        int incl = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (i < n1) rowptr[i] = carry + incl - d;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (g == B - 1 && lane == 0) rowptr[n1] = e1;
    for (int p = e0 + lane; p < e1; p += 32) dst_src[p] = n0 + (int)nbr[p];
    const int64_t f0 = (int64_t)n0 * F, f1 = (int64_t)n1 * F;
    for (int64_t k = f0 + lane; k < f1; k += 32) x[k] = (float)xq[k];
}

}  // namespace
}  // namespace glam

using namespace glam;

extern "C" int glam_unpack_graphs(const uint8_t* n_g, const uint16_t* e_g, const uint8_t* deg, const uint8_t* nbr, const uint8_t* xq,
                                  int64_t num_graphs, int64_t num_nodes, int64_t num_edges, int node_dim, int32_t* graph_ptr,
                                  int32_t* edge_ptr, int32_t* dst_rowptr, int32_t* dst_src, float* x, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && num_edges >= 0 && node_dim > 0, "glam_unpack_graphs: bad sizes");
    GLAM_REQUIRE(graph_ptr && edge_ptr && dst_rowptr, "glam_unpack_graphs: null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (num_graphs == 0) {
        cudaMemsetAsync(graph_ptr, 0, sizeof(int32_t), stream);
        cudaMemsetAsync(edge_ptr, 0, sizeof(int32_t), stream);
        cudaMemsetAsync(dst_rowptr, 0, sizeof(int32_t), stream);
        return 0;
    }
    GLAM_REQUIRE(n_g && e_g && deg && xq && x && (num_edges == 0 || (nbr && dst_src)), "glam_unpack_graphs: null input");
    GLAM_REQUIRE(num_nodes < ((int64_t)1 << 31) && num_edges < ((int64_t)1 << 31), "glam_unpack_graphs: too large");
    packed_scan_kernel<<<1, 1024, 0, stream>>>(n_g, e_g, num_graphs, graph_ptr, edge_ptr);
    GLAM_CHECK_LAUNCH();
    packed_unpack_kernel<<<(unsigned)((num_graphs * 32 + 255) / 256), 256, 0, stream>>>(graph_ptr, edge_ptr, deg, nbr, xq, num_graphs, node_dim,
                                                                                       dst_rowptr, dst_src, x);
    GLAM_CHECK_LAUNCH();
    return 0;
}
