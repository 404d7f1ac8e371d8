// SURVEY.md §8(f) "next" rows beside the triplet path: the GCN tower the reference uses by default for proteins in the
// drug-target model (`_GCNConv` -> PyG GCNConv(in,out) @1.7.2, src_2gi_dti_scr/layer.py wrappers, run.py:19), and a
// generic deterministic CSR row aggregation it is built on.
//
//   GCNConv: edges = non-self edges + exactly one self loop per node (add_remaining_self_loops), all weights 1;
//            deg_i = 1 + #{non-self in-edges of i};  out = D^-1/2 (A + I) D^-1/2 (x W) + b.
//
// The aggregation runs over the destination-sorted CSR forward and over the source-sorted CSR backward (the transpose of
// the normalised adjacency), so there are no atomics and every sum has a fixed order.
#include "common.cuh"

namespace glam {

// dinv[i] = (1 + non-self in-degree)^-1/2 ; w[p] = 0 for self edges (replaced by the single unit self loop), else filled later
__global__ void gcn_dinv_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t N, float* __restrict__ dinv) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int deg = 1;
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) deg += col[p] != i;
        dinv[i] = rsqrtf((float)deg);
    }
}
// per-edge weight in the order of the given CSR (rows = `row of p`, col[p] = other endpoint): dinv[row] * dinv[col], 0 on self edges
__global__ void gcn_edge_weight_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ dinv,
                                       int64_t N, float* __restrict__ w) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float di = dinv[i];
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) w[p] = col[p] != i ? di * dinv[col[p]] : 0.f;
    }
}

// out[i,:] = (accumulate ? out[i,:] : 0) + row_scale[i] * ( sum_p w[p] * Y[col[p],:] + self_w[i] * Y[i,:] ) + bias
// one sub-warp group of G lanes per row, lanes over 16-byte chunks (F % 4 == 0) — or scalar lanes (VEC = false)
template <bool VEC>
__global__ void __launch_bounds__(256)
csr_aggregate_kernel(const float* __restrict__ Y, int64_t ldy, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const float* __restrict__ w, const float* __restrict__ self_w, const float* __restrict__ row_scale,
                     const float* __restrict__ bias, int64_t N, int F, int G, float* __restrict__ out, int64_t ldo, int accumulate) {
    const int lane_in_group = threadIdx.x % G;
    const int64_t group0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int nchunk = VEC ? F >> 2 : F;
    for (int64_t i = group0; i < N; i += ngroups) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        const float sw = self_w ? self_w[i] : 0.f, rs = row_scale ? row_scale[i] : 1.f;
        for (int q = lane_in_group; q < nchunk; q += G) {
            if (VEC) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = beg; p < end; ++p) {
                    const float wp = w ? w[p] : 1.f;
                    const float4 y = *reinterpret_cast<const float4*>(Y + (int64_t)col[p] * ldy + 4 * q);
                    acc.x = fmaf(wp, y.x, acc.x); acc.y = fmaf(wp, y.y, acc.y); acc.z = fmaf(wp, y.z, acc.z); acc.w = fmaf(wp, y.w, acc.w);
                }
                if (self_w) {
                    const float4 y = *reinterpret_cast<const float4*>(Y + i * ldy + 4 * q);
                    acc.x = fmaf(sw, y.x, acc.x); acc.y = fmaf(sw, y.y, acc.y); acc.z = fmaf(sw, y.z, acc.z); acc.w = fmaf(sw, y.w, acc.w);
                }
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias) b = *reinterpret_cast<const float4*>(bias + 4 * q);
                float4* o = reinterpret_cast<float4*>(out + i * ldo + 4 * q);
                float4 r = make_float4(fmaf(rs, acc.x, b.x), fmaf(rs, acc.y, b.y), fmaf(rs, acc.z, b.z), fmaf(rs, acc.w, b.w));
                if (accumulate) { const float4 old = *o; r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
                *o = r;
            } else {
                float acc = 0.f;
                for (int p = beg; p < end; ++p) acc = fmaf(w ? w[p] : 1.f, Y[(int64_t)col[p] * ldy + q], acc);
                if (self_w) acc = fmaf(sw, Y[i * ldy + q], acc);
                float r = fmaf(rs, acc, bias ? bias[q] : 0.f);
                if (accumulate) r += out[i * ldo + q];
                out[i * ldo + q] = r;
            }
        }
    }
}

}  // namespace glam

using namespace glam;

static int ew_blocks(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int glam_gcn_norm(const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* src_rowptr, const int32_t* src_dst,
                             int64_t num_nodes, float* dinv, float* w_dst, float* w_src, void* stream_) {
    GLAM_REQUIRE(num_nodes >= 0, "glam_gcn_norm: bad arguments");
    if (num_nodes == 0) return 0;
    GLAM_REQUIRE(dst_rowptr && dst_src && dinv && w_dst, "glam_gcn_norm: null pointer");
    cudaStream_t st = (cudaStream_t)stream_;
    gcn_dinv_kernel<<<ew_blocks(num_nodes), 256, 0, st>>>(dst_rowptr, dst_src, num_nodes, dinv);
    GLAM_CHECK_LAUNCH();
    gcn_edge_weight_kernel<<<ew_blocks(num_nodes), 256, 0, st>>>(dst_rowptr, dst_src, dinv, num_nodes, w_dst);
    GLAM_CHECK_LAUNCH();
    if (w_src) {
        GLAM_REQUIRE(src_rowptr && src_dst, "glam_gcn_norm: source CSR missing");
        gcn_edge_weight_kernel<<<ew_blocks(num_nodes), 256, 0, st>>>(src_rowptr, src_dst, dinv, num_nodes, w_src);
        GLAM_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int glam_csr_aggregate(const float* Y, int64_t ldy, const int32_t* rowptr, const int32_t* col, const float* edge_w,
                                  const float* self_w, const float* row_scale, const float* bias, int64_t num_rows, int features,
                                  float* out, int64_t ldo, int accumulate, void* stream_) {
    GLAM_REQUIRE(num_rows >= 0 && features > 0 && ldy >= features && ldo >= features, "glam_csr_aggregate: bad shape");
    if (num_rows == 0) return 0;
    GLAM_REQUIRE(Y && rowptr && col && out, "glam_csr_aggregate: null pointer");
    const bool vec = (features & 3) == 0 && (ldy & 3) == 0 && (ldo & 3) == 0 && (((uintptr_t)Y | (uintptr_t)out | (uintptr_t)bias) & 15) == 0;
    const int nchunk = vec ? features >> 2 : features;
    int G = 1;
    while (G < nchunk && G < 32) G <<= 1;
    const int64_t threads = num_rows * G;
    int64_t blocks = (threads + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    if (vec)
        csr_aggregate_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(Y, ldy, rowptr, col, edge_w, self_w, row_scale, bias,
                                                                                     num_rows, features, G, out, ldo, accumulate);
    else
        csr_aggregate_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(Y, ldy, rowptr, col, edge_w, self_w, row_scale, bias,
                                                                                      num_rows, features, G, out, ldo, accumulate);
    GLAM_CHECK_LAUNCH();
    return 0;
}
