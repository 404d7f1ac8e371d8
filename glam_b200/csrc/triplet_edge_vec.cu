// Triplet attention edge phase, vectorised path (channels % 4 == 0): the layout every reference configuration
// with hid_dim_alpha in {4} (C = 36, 60) hits.  Same math as triplet_edge.cu, split so that
// every kernel has abundant memory-level parallelism:
//
//   attention:  one THREAD per destination — logits from s_i[i] + s_j[src] + edge_attr . att_edge, leaky_relu, PyG
//               softmax exp(a - max) / (sum + 1e-16) over the in-edges  -> alpha [E,H]      (tiny traffic, 10^5 threads)
//   aggregate:  one WARP per destination, lanes over the 16-byte channel chunks of the H*C message row — the
//               gather of xp[src, :] is one coalesced 16-byte load per lane per edge; e_ij is rebuilt from
//               edge_attr and weight_edge (shared memory) so no [E,HC] tensor exists
//   backward:   the same three shapes (per-edge dot products by warp, softmax/leaky backward by thread, scatter by
//               SOURCE through the src-sorted CSR by warp); g_weight_edge accumulates in registers per warp and is
//               reduced in fixed order.  No atomics anywhere.
#include "common.cuh"

namespace glam {

constexpr int kVecWarps = 8;
constexpr int kVecThreads = kVecWarps * 32;
constexpr int kVecMaxDe = 8;

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4fma(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// edge_attr row summary: molecular bond features are one-hot, so e_ij = edge_attr . weight_edge is a single scaled row of
// weight_edge; general rows (protein contact features) take the full De-term sum.
struct EaRow { int nz, ty; float val; };
__device__ __forceinline__ EaRow scan_ea(const float* __restrict__ row, int De) {
    EaRow r{0, 0, 0.f};
    for (int d = 0; d < De; ++d) {
        const float e = row[d];
        if (e != 0.f) { ++r.nz; r.ty = d; r.val = e; }
    }
    return r;
}
__device__ __forceinline__ float4 ep_chunk(const EaRow& r, const float* __restrict__ row, int De, const float4* We4, int nq, int q) {
    if (r.nz == 1) {
        const float4 w = We4[r.ty * nq + q];
        return make_float4(r.val * w.x, r.val * w.y, r.val * w.z, r.val * w.w);
    }
    float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 0; d < De; ++d) ep = f4fma(row[d], We4[d * nq + q], ep);
    return ep;
}

template <int H>
__device__ __forceinline__ float pickh(const float (&a)[H], int h) {
    float v = a[0];
#pragma unroll
    for (int q = 1; q < H; ++q) v = (h == q) ? a[q] : v;
    return v;
}

// ------------------------------------------------------------------------------------------------ attention (fwd)
template <int H>
__global__ void __launch_bounds__(256)
edge_alpha_fwd_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ att_edge,
                      const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs, int64_t N, int HC, int De,
                      float slope, float* __restrict__ alpha) {
    __shared__ float Ae[kVecMaxDe * H];
    for (int i = threadIdx.x; i < De * H; i += blockDim.x) Ae[i] = att_edge[i];
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        if (beg == end) continue;
        float si[H], mx[H], sum[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { si[h] = xpe[i * ld + HC + h]; mx[h] = -INFINITY; sum[h] = 0.f; }
        for (int p = beg; p < end; ++p) {
            const int64_t j = srcs[p];
            float l[H];
#pragma unroll
            for (int h = 0; h < H; ++h) l[h] = si[h] + xpe[j * ld + HC + H + h];
            for (int d = 0; d < De; ++d) {
                const float e = ea[(int64_t)p * De + d];
#pragma unroll
                for (int h = 0; h < H; ++h) l[h] = fmaf(e, Ae[d * H + h], l[h]);
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                l[h] = l[h] > 0.f ? l[h] : slope * l[h];
                mx[h] = fmaxf(mx[h], l[h]);
                alpha[(int64_t)p * H + h] = l[h];
            }
        }
        for (int p = beg; p < end; ++p) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float e = expf(alpha[(int64_t)p * H + h] - mx[h]);
                alpha[(int64_t)p * H + h] = e;
                sum[h] += e;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) sum[h] += 1e-16f;
        for (int p = beg; p < end; ++p) {
#pragma unroll
            for (int h = 0; h < H; ++h) alpha[(int64_t)p * H + h] = alpha[(int64_t)p * H + h] / sum[h];
        }
    }
}

// ------------------------------------------------------------------------------------------------ aggregate (fwd)
// G lanes per destination (32/G destinations per warp), each lane owns CPL 16-byte chunks (q = gl + G*t): the per-edge
// scalar work (index, alpha, edge_attr loads) is amortised over CPL chunks instead of being repeated by every lane.
template <int H, int G, int CPL, bool USE_EP>
__global__ void __launch_bounds__(kVecThreads)
edge_aggregate_fwd_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ w_edge,
                          const float* __restrict__ alpha, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs,
                          int64_t N, int C, int De, float* __restrict__ agg) {
    extern __shared__ float4 We4[];                    // [De][HC/4]
    const int HC = H * C, nq = HC >> 2;
    if (USE_EP)
        for (int i = threadIdx.x; i < De * nq; i += blockDim.x) We4[i] = ldg4(w_edge + 4 * i);
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
    constexpr int kPerWarp = 32 / G;
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (gl + G * t)) / C;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp0 * kPerWarp; base < N; base += nwarps * kPerWarp) {
        const int64_t i = base + sub;
        if (i >= N) continue;
        const int beg = rowptr[i], end = rowptr[i + 1];
        float4 acc[CPL];
#pragma unroll
        for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        // two edges per iteration: both gathers (2 x CPL 16-byte loads per lane) are in flight before any math
        for (int p = beg; p < end; p += 2) {
            const bool two = p + 1 < end;
            const int p1 = two ? p + 1 : p;
            const float* xj0 = xpe + (int64_t)srcs[p] * ld;
            const float* xj1 = xpe + (int64_t)srcs[p1] * ld;
            float a0[H], a1[H];
#pragma unroll
            for (int h = 0; h < H; ++h) { a0[h] = alpha[(int64_t)p * H + h]; a1[h] = two ? alpha[(int64_t)p1 * H + h] : 0.f; }
            float4 m0[CPL], m1[CPL];
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = gl + G * t;
                if (q < nq) { m0[t] = ldg4(xj0 + 4 * q); m1[t] = ldg4(xj1 + 4 * q); }
            }
            const float* ea0 = ea + (int64_t)p * De;
            const float* ea1 = ea + (int64_t)p1 * De;
            EaRow e0{0, 0, 0.f}, e1{0, 0, 0.f};
            if (USE_EP) { e0 = scan_ea(ea0, De); e1 = scan_ea(ea1, De); }
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = gl + G * t;
                if (q < nq) {
                    float4 v0 = m0[t], v1 = m1[t];
                    if (USE_EP) { v0 = f4mul(v0, ep_chunk(e0, ea0, De, We4, nq, q)); v1 = f4mul(v1, ep_chunk(e1, ea1, De, We4, nq, q)); }
                    acc[t] = f4fma(pickh<H>(a0, hq[t]), v0, acc[t]);
                    acc[t] = f4fma(pickh<H>(a1, hq[t]), v1, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = gl + G * t;
            if (q < nq) *reinterpret_cast<float4*>(agg + i * HC + 4 * q) = acc[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------ fused tile forward
// One CTA owns a tile of 64 consecutive destinations.  The per-edge scalar work is done ONCE per edge by one thread
// (coalesced index loads, s_j gather, edge_attr summary) into shared-memory records; a thread per destination runs the
// softmax over its records; then sub-warp groups aggregate, reading the records as broadcasts and gathering x_j chunks
// with two edges in flight.  Dependent global round trips per tile: rowptr -> (src, edge_attr) -> s_j -> x_j gathers.
constexpr int kTileDst = 128;
constexpr int kTileEdges = 1536;

template <int H, int G, int CPL, bool USE_EP>
__global__ void __launch_bounds__(kVecThreads)
edge_tile_fwd_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ w_edge,
                     const float* __restrict__ att_edge, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs,
                     int64_t N, int C, int De, float slope, float* __restrict__ agg, float* __restrict__ alpha) {
    extern __shared__ float4 We4[];                                   // [De][nq]
    __shared__ float Ae[kVecMaxDe * H];
    __shared__ int rec_src[kTileEdges], rec_ty[kTileEdges], rp[kTileDst + 1];
    __shared__ float rec_val[kTileEdges], rec_a[kTileEdges][H];
    const int HC = H * C, nq = HC >> 2;
    if (USE_EP)
        for (int i = threadIdx.x; i < De * nq; i += blockDim.x) We4[i] = ldg4(w_edge + 4 * i);
    for (int i = threadIdx.x; i < De * H; i += blockDim.x) Ae[i] = att_edge[i];
    const int tid = threadIdx.x, lane = tid & 31, grp = tid / G, gl = lane % G;
    constexpr int kGroups = kVecThreads / G;
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (gl + G * t)) / C;
    const int64_t ntiles = (N + kTileDst - 1) / kTileDst;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = tile * kTileDst;
        const int nd = (int)min((int64_t)kTileDst, N - t0);
        __syncthreads();                                              // previous tile's records are dead
        if (tid <= nd) rp[tid] = rowptr[t0 + tid];
        __syncthreads();
        const int e0 = rp[0], ne = rp[nd] - e0;
        if (ne > kTileEdges) {
            // ---- oversized tile (hub destinations): per-destination routine straight from global memory
            for (int d = tid; d < nd; d += blockDim.x) {
                const int64_t i = t0 + d;
                const int beg = rp[d], end = rp[d + 1];
                float si[H], mx[H], sum[H];
#pragma unroll
                for (int h = 0; h < H; ++h) { si[h] = xpe[i * ld + HC + h]; mx[h] = -INFINITY; sum[h] = 0.f; }
                for (int p = beg; p < end; ++p) {
                    const int64_t j = srcs[p];
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        float l = si[h] + xpe[j * ld + HC + H + h];
                        for (int dd = 0; dd < De; ++dd) l = fmaf(ea[(int64_t)p * De + dd], Ae[dd * H + h], l);
                        l = l > 0.f ? l : slope * l;
                        mx[h] = fmaxf(mx[h], l);
                        alpha[(int64_t)p * H + h] = l;
                    }
                }
                for (int p = beg; p < end; ++p)
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        const float e = expf(alpha[(int64_t)p * H + h] - mx[h]);
                        alpha[(int64_t)p * H + h] = e;
                        sum[h] += e;
                    }
                for (int p = beg; p < end; ++p)
#pragma unroll
                    for (int h = 0; h < H; ++h) alpha[(int64_t)p * H + h] = alpha[(int64_t)p * H + h] / (sum[h] + 1e-16f);
            }
            __syncthreads();
            for (int d = grp; d < nd; d += kGroups) {
                const int64_t i = t0 + d;
                float4 acc[CPL];
#pragma unroll
                for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p = rp[d]; p < rp[d + 1]; ++p) {
                    const float* xj = xpe + (int64_t)srcs[p] * ld;
                    const float* earow = ea + (int64_t)p * De;
                    EaRow er{0, 0, 0.f};
                    if (USE_EP) er = scan_ea(earow, De);
                    float a[H];
#pragma unroll
                    for (int h = 0; h < H; ++h) a[h] = alpha[(int64_t)p * H + h];
#pragma unroll
                    for (int t = 0; t < CPL; ++t) {
                        const int q = gl + G * t;
                        if (q < nq) {
                            float4 m = ldg4(xj + 4 * q);
                            if (USE_EP) m = f4mul(m, ep_chunk(er, earow, De, We4, nq, q));
                            acc[t] = f4fma(pickh<H>(a, hq[t]), m, acc[t]);
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const int q = gl + G * t;
                    if (q < nq) *reinterpret_cast<float4*>(agg + i * HC + 4 * q) = acc[t];
                }
            }
            continue;
        }
        // ---- phase A: one thread per edge -> records (source, edge_attr summary, s_j + s_e part of the logit)
        for (int e = tid; e < ne; e += blockDim.x) {
            const int p = e0 + e;
            const int j = srcs[p];
            const float* earow = ea + (int64_t)p * De;
            float l[H];
#pragma unroll
            for (int h = 0; h < H; ++h) l[h] = xpe[(int64_t)j * ld + HC + H + h];
            int nz = 0, ty = 0;
            float val = 0.f;
            for (int d = 0; d < De; ++d) {
                const float v = earow[d];
                if (v != 0.f) { ++nz; ty = d; val = v; }
#pragma unroll
                for (int h = 0; h < H; ++h) l[h] = fmaf(v, Ae[d * H + h], l[h]);
            }
            rec_src[e] = j;
            rec_ty[e] = nz == 1 ? ty : -1;
            rec_val[e] = val;
#pragma unroll
            for (int h = 0; h < H; ++h) rec_a[e][h] = l[h];
        }
        __syncthreads();
        // ---- phase B: one thread per destination -> PyG softmax over its records; alpha to global (saved for backward)
        if (tid < nd) {
            const int64_t i = t0 + tid;
            const int beg = rp[tid] - e0, end = rp[tid + 1] - e0;
            float si[H], mx[H], sum[H];
#pragma unroll
            for (int h = 0; h < H; ++h) { si[h] = xpe[i * ld + HC + h]; mx[h] = -INFINITY; sum[h] = 0.f; }
            for (int e = beg; e < end; ++e)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float l = si[h] + rec_a[e][h];
                    l = l > 0.f ? l : slope * l;
                    rec_a[e][h] = l;
                    mx[h] = fmaxf(mx[h], l);
                }
            for (int e = beg; e < end; ++e)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float x = expf(rec_a[e][h] - mx[h]);
                    rec_a[e][h] = x;
                    sum[h] += x;
                }
            for (int e = beg; e < end; ++e)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float a = rec_a[e][h] / (sum[h] + 1e-16f);
                    rec_a[e][h] = a;
                    alpha[(int64_t)(e0 + e) * H + h] = a;
                }
        }
        __syncthreads();
        // ---- phase C: sub-warp group per destination, lanes over 16-byte chunks, two edges in flight
        for (int d = grp; d < nd; d += kGroups) {
            const int64_t i = t0 + d;
            const int beg = rp[d] - e0, end = rp[d + 1] - e0;
            float4 acc[CPL];
#pragma unroll
            for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = beg; e < end; e += 2) {
                const bool two = e + 1 < end;
                const int e1 = two ? e + 1 : e;
                const float* xj0 = xpe + (int64_t)rec_src[e] * ld;
                const float* xj1 = xpe + (int64_t)rec_src[e1] * ld;
                float4 m0[CPL], m1[CPL];
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const int q = gl + G * t;
                    if (q < nq) { m0[t] = ldg4(xj0 + 4 * q); m1[t] = ldg4(xj1 + 4 * q); }
                }
                float a0[H], a1[H];
#pragma unroll
                for (int h = 0; h < H; ++h) { a0[h] = rec_a[e][h]; a1[h] = two ? rec_a[e1][h] : 0.f; }
                const int ty0 = rec_ty[e], ty1 = rec_ty[e1];
                const float v0 = rec_val[e], v1 = rec_val[e1];
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const int q = gl + G * t;
                    if (q < nq) {
                        float4 x0 = m0[t], x1 = m1[t];
                        float c0 = pickh<H>(a0, hq[t]), c1 = pickh<H>(a1, hq[t]);
                        if (USE_EP) {
                            if (ty0 >= 0) { x0 = f4mul(x0, We4[ty0 * nq + q]); c0 *= v0; }
                            else {
                                float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
                                for (int dd = 0; dd < De; ++dd) ep = f4fma(ea[(int64_t)(e0 + e) * De + dd], We4[dd * nq + q], ep);
                                x0 = f4mul(x0, ep);
                            }
                            if (ty1 >= 0) { x1 = f4mul(x1, We4[ty1 * nq + q]); c1 *= v1; }
                            else {
                                float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
                                for (int dd = 0; dd < De; ++dd) ep = f4fma(ea[(int64_t)(e0 + e1) * De + dd], We4[dd * nq + q], ep);
                                x1 = f4mul(x1, ep);
                            }
                        }
                        acc[t] = f4fma(c0, x0, acc[t]);
                        acc[t] = f4fma(c1, x1, acc[t]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = gl + G * t;
                if (q < nq) *reinterpret_cast<float4*>(agg + i * HC + 4 * q) = acc[t];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward 1: per-edge dots
// g_alpha[p,h] = <g_agg[dst,h,:], e_ij (.) x_j>  (written into g_logit as scratch) and the register-resident partial of
// g_w_edge[d, :] += edge_attr[p,d] * alpha * g_agg[dst] (.) x_j
template <int H, int CPL, bool USE_EP>
__global__ void __launch_bounds__(kVecThreads)
edge_dots_bwd_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ w_edge,
                     const float* __restrict__ alpha, const float* __restrict__ g_agg, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ srcs, int64_t N, int C, int De, float* __restrict__ g_logit,
                     float* __restrict__ gwe_partial) {
    extern __shared__ float4 smem4[];
    const int HC = H * C, nq = HC >> 2;
    float4* We4 = smem4;                               // [De][nq]
    float4* red4 = smem4 + (USE_EP ? De * nq : 0);     // [kVecWarps][De][nq]  (end of kernel only)
    if (USE_EP)
        for (int i = threadIdx.x; i < De * nq; i += blockDim.x) We4[i] = ldg4(w_edge + 4 * i);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (lane + 32 * t)) / C;
    float4 gw[USE_EP ? kVecMaxDe : 1][CPL];
#pragma unroll
    for (int d = 0; d < (USE_EP ? kVecMaxDe : 1); ++d)
#pragma unroll
        for (int t = 0; t < CPL; ++t) gw[d][t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        if (beg == end) continue;
        float4 ga[CPL];
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = lane + 32 * t;
            ga[t] = q < nq ? ldg4(g_agg + i * HC + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = beg; p < end; ++p) {
            const int64_t j = srcs[p];
            float a[H], part[H];
#pragma unroll
            for (int h = 0; h < H; ++h) { a[h] = alpha[(int64_t)p * H + h]; part[h] = 0.f; }
            const float* earow = ea + (int64_t)p * De;
            EaRow er{0, 0, 0.f};
            if (USE_EP) er = scan_ea(earow, De);
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = lane + 32 * t;
                if (q < nq) {
                    const float4 gm = f4mul(ga[t], ldg4(xpe + j * ld + 4 * q));
                    float v;
                    if (USE_EP) {
                        const float ah = pickh<H>(a, hq[t]);
                        v = f4dot(gm, ep_chunk(er, earow, De, We4, nq, q));
                        if (er.nz == 1) {
                            // one-hot bond feature: a single weight_edge row receives this edge's gradient
                            const float c = er.val * ah;
#pragma unroll
                            for (int d = 0; d < kVecMaxDe; ++d)
                                if (d == er.ty) gw[d][t] = f4fma(c, gm, gw[d][t]);
                        } else {
#pragma unroll
                            for (int d = 0; d < kVecMaxDe; ++d)
                                if (d < De) gw[d][t] = f4fma(earow[d] * ah, gm, gw[d][t]);
                        }
                    } else {
                        v = gm.x + gm.y + gm.z + gm.w;
                    }
#pragma unroll
                    for (int h = 0; h < H; ++h) part[h] += (hq[t] == h) ? v : 0.f;
                }
            }
#pragma unroll
            for (int h = 0; h < H; ++h) part[h] = warp_sum(part[h]);
            if (lane < H) g_logit[(int64_t)p * H + lane] = pickh<H>(part, lane);
        }
    }
    if (USE_EP) {
        // fixed-order reduction of the per-warp register partials
#pragma unroll
        for (int d = 0; d < kVecMaxDe; ++d)
            if (d < De)
#pragma unroll
                for (int t = 0; t < CPL; ++t) {
                    const int q = lane + 32 * t;
                    if (q < nq) red4[(wid * De + d) * nq + q] = gw[d][t];
                }
        __syncthreads();
        float4* P = reinterpret_cast<float4*>(gwe_partial) + (int64_t)blockIdx.x * De * nq;
        for (int idx = threadIdx.x; idx < De * nq; idx += blockDim.x) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kVecWarps; ++w) {
                const float4 v = red4[w * De * nq + idx];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            P[idx] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward 1': edge-parallel dots
// g_alpha[p,h] depends on edge p alone, so with dst_dst[p] available the dots need no segment structure: one sub-warp
// group per EDGE (grid-stride), both row gathers (g_agg[dst], xp[src]) issued up front, no degree loops, no divergence.
// g_weight_edge (edge_dim <= 4) accumulates in registers with one predicated FMA per weight_edge row.
constexpr int kEpMaxDe = 4;
template <int H, int G, int CPL, bool USE_EP>
__global__ void __launch_bounds__(kVecThreads)
edge_dots_ep_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ w_edge,
                    const float* __restrict__ alpha, const float* __restrict__ g_agg, const int32_t* __restrict__ srcs,
                    const int32_t* __restrict__ dsts, int64_t E, int C, int De, float* __restrict__ g_logit,
                    float* __restrict__ gwe_partial) {
    extern __shared__ float4 smem4[];
    const int HC = H * C, nq = HC >> 2;
    constexpr int kGroups = kVecThreads / G;
    float4* We4 = smem4;                               // [De][nq]
    float4* red4 = smem4 + (USE_EP ? De * nq : 0);     // [kGroups][De][nq]  (end of kernel only)
    if (USE_EP)
        for (int i = threadIdx.x; i < De * nq; i += blockDim.x) We4[i] = ldg4(w_edge + 4 * i);
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = threadIdx.x / G, gl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (gl + G * t)) / C;
    // g_weight_edge accumulates in this group's private shared-memory slab [De][nq] (zeroed here, reduced over the groups
    // in a fixed order at the end): no registers held across the edge loop, no predicated per-row FMAs
    float4* slab = red4 + grp * De * nq;
    if (USE_EP)
        for (int i = threadIdx.x; i < kGroups * De * nq; i += blockDim.x) red4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int64_t g0 = (int64_t)blockIdx.x * kGroups + grp, gstride = (int64_t)gridDim.x * kGroups;
    for (int64_t p = g0; p < E; p += gstride) {
        const float* xj = xpe + (int64_t)srcs[p] * ld;
        const float* gi = g_agg + (int64_t)dsts[p] * HC;
        float4 xv[CPL], gv[CPL];
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = gl + G * t;
            if (q < nq) { xv[t] = ldg4(xj + 4 * q); gv[t] = ldg4(gi + 4 * q); }
        }
        float a[H], part[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { a[h] = alpha[p * H + h]; part[h] = 0.f; }
        const float* earow = ea + p * De;
        EaRow er{0, 0, 0.f};
        if (USE_EP) er = scan_ea(earow, De);
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = gl + G * t;
            if (q < nq) {
                const float4 gm = f4mul(gv[t], xv[t]);
                float v;
                if (USE_EP) {
                    const float ah = pickh<H>(a, hq[t]);
                    if (er.nz == 1) {
                        v = er.val * f4dot(gm, We4[er.ty * nq + q]);
                        float4* w = slab + er.ty * nq + q;
                        *w = f4fma(er.val * ah, gm, *w);
                    } else {
                        float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int dd = 0; dd < kEpMaxDe; ++dd)
                            if (dd < De) {
                                const float ed = earow[dd];
                                ep = f4fma(ed, We4[dd * nq + q], ep);
                                float4* w = slab + dd * nq + q;
                                *w = f4fma(ed * ah, gm, *w);
                            }
                        v = f4dot(gm, ep);
                    }
                } else {
                    v = gm.x + gm.y + gm.z + gm.w;
                }
#pragma unroll
                for (int h = 0; h < H; ++h) part[h] += (hq[t] == h) ? v : 0.f;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) part[h] += __shfl_xor_sync(gmask, part[h], o);
        }
        if (gl < H) g_logit[p * H + gl] = pickh<H>(part, gl);
    }
    if (USE_EP) {
        // fixed-order reduction of the per-group slabs -> one partial per CTA
        __syncthreads();
        float4* P = reinterpret_cast<float4*>(gwe_partial) + (int64_t)blockIdx.x * De * nq;
        for (int idx = threadIdx.x; idx < De * nq; idx += blockDim.x) {
            float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int w = 0; w < kGroups; ++w) {
                const float4 v = red4[w * De * nq + idx];
                sacc.x += v.x; sacc.y += v.y; sacc.z += v.z; sacc.w += v.w;
            }
            P[idx] = sacc;
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward 2: softmax + leaky
template <int H>
__global__ void __launch_bounds__(256)
edge_softmax_bwd_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ att_edge,
                        const float* __restrict__ alpha, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs,
                        int64_t N, int HC, int De, float slope, float* __restrict__ g_logit, float* __restrict__ g_xpe) {
    __shared__ float Ae[kVecMaxDe * H];
    for (int i = threadIdx.x; i < De * H; i += blockDim.x) Ae[i] = att_edge[i];
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        float si[H], dot[H], gsi[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { si[h] = xpe[i * ld + HC + h]; dot[h] = 0.f; gsi[h] = 0.f; }
        for (int p = beg; p < end; ++p) {
#pragma unroll
            for (int h = 0; h < H; ++h) dot[h] = fmaf(alpha[(int64_t)p * H + h], g_logit[(int64_t)p * H + h], dot[h]);
        }
        for (int p = beg; p < end; ++p) {
            const int64_t j = srcs[p];
            float l[H];
#pragma unroll
            for (int h = 0; h < H; ++h) l[h] = si[h] + xpe[j * ld + HC + H + h];
            for (int d = 0; d < De; ++d) {
                const float e = ea[(int64_t)p * De + d];
#pragma unroll
                for (int h = 0; h < H; ++h) l[h] = fmaf(e, Ae[d * H + h], l[h]);
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float g = alpha[(int64_t)p * H + h] * (g_logit[(int64_t)p * H + h] - dot[h]);
                g *= (l[h] > 0.f ? 1.f : slope);
                g_logit[(int64_t)p * H + h] = g;
                gsi[h] += g;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) g_xpe[i * ld + HC + h] = gsi[h];
    }
}

// ------------------------------------------------------------------------------------------------ backward 3: scatter by source
template <int H, int G, int CPL, bool USE_EP>
__global__ void __launch_bounds__(kVecThreads)
edge_source_bwd_kernel(const float* __restrict__ ea, const float* __restrict__ w_edge, const float* __restrict__ alpha,
                       const float* __restrict__ g_agg, const float* __restrict__ g_logit, const int32_t* __restrict__ src_rowptr,
                       const int32_t* __restrict__ src_pos, const int32_t* __restrict__ src_dst, int64_t N, int C, int De,
                       float* __restrict__ g_xpe, int64_t ld) {
    extern __shared__ float4 We4[];
    const int HC = H * C, nq = HC >> 2;
    if (USE_EP)
        for (int i = threadIdx.x; i < De * nq; i += blockDim.x) We4[i] = ldg4(w_edge + 4 * i);
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
    constexpr int kPerWarp = 32 / G;
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (gl + G * t)) / C;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp0 * kPerWarp; base < N; base += nwarps * kPerWarp) {
        const int64_t j = base + sub;
        if (j >= N) continue;
        const int beg = src_rowptr[j], end = src_rowptr[j + 1];
        float4 acc[CPL];
#pragma unroll
        for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        float gsj = 0.f;
        for (int k = beg; k < end; ++k) {
            const int p = src_pos[k];
            const float* gi = g_agg + (int64_t)src_dst[k] * HC;
            float a[H];
#pragma unroll
            for (int h = 0; h < H; ++h) a[h] = alpha[(int64_t)p * H + h];
            if (gl < H) gsj += g_logit[(int64_t)p * H + gl];
            const float* earow = ea + (int64_t)p * De;
            EaRow er{0, 0, 0.f};
            if (USE_EP) er = scan_ea(earow, De);
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = gl + G * t;
                if (q < nq) {
                    float4 m = ldg4(gi + 4 * q);
                    if (USE_EP) m = f4mul(m, ep_chunk(er, earow, De, We4, nq, q));
                    acc[t] = f4fma(pickh<H>(a, hq[t]), m, acc[t]);
                }
            }
        }
        float* out = g_xpe + j * ld;
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = gl + G * t;
            if (q < nq) *reinterpret_cast<float4*>(out + 4 * q) = acc[t];
        }
        if (gl < H) out[HC + H + gl] = gsj;
        for (int k = HC + 2 * H + gl; k < ld; k += G) out[k] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int vec_warp_grid(int64_t N) {
    int64_t want = (N + kVecWarps - 1) / kVecWarps;
    int64_t cap = (int64_t)kNumSMs * 8;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}
static int vec_thread_grid(int64_t N) {
    int64_t want = (N + 255) / 256;
    int64_t cap = (int64_t)kNumSMs * 8;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

bool edge_vec_eligible(const float* xpe, int64_t ldxp, int heads, int C, int De, const float* a1, const float* a2) {
    if ((C & 3) || (ldxp & 3) || De > kVecMaxDe || heads < 1 || heads > GLAM_MAX_HEADS) return false;
    if (heads * C > 384) return false;                                   // <= 3 chunks per lane
    if (((uintptr_t)xpe & 15) || ((uintptr_t)a1 & 15) || ((uintptr_t)a2 & 15)) return false;
    return true;
}

// sub-warp geometry: G lanes per destination, CPL chunks per lane  (nq = H*C/4 chunks per row)
static void vec_geometry(int nq, int* G, int* cpl) {
    if (nq <= 32) { *G = 8; *cpl = 4; } else if (nq <= 64) { *G = 16; *cpl = 4; } else { *G = 32; *cpl = 3; }
}
#define GLAM_VEC_GEO(H_, EP_, g, ...)                                              \
    switch (g) {                                                                   \
        case 8: { constexpr int HH_ = H_, G_ = 8, CPL_ = 4; constexpr bool UE_ = EP_; __VA_ARGS__; } break;   \
        case 16: { constexpr int HH_ = H_, G_ = 16, CPL_ = 4; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
        default: { constexpr int HH_ = H_, G_ = 32, CPL_ = 3; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
    }
#define GLAM_VEC_GDISPATCH(heads, use_ep, g, ...)                       \
    if (!(use_ep)) { GLAM_VEC_GEO(1, false, g, __VA_ARGS__) }           \
    else switch (heads) {                                               \
        case 1: GLAM_VEC_GEO(1, true, g, __VA_ARGS__) break;            \
        case 2: GLAM_VEC_GEO(2, true, g, __VA_ARGS__) break;            \
        case 3: GLAM_VEC_GEO(3, true, g, __VA_ARGS__) break;            \
        default: GLAM_VEC_GEO(4, true, g, __VA_ARGS__) break;           \
    }
#define GLAM_VEC_CPL(H_, EP_, cpl, ...)                                            \
    switch (cpl) {                                                                 \
        case 1: { constexpr int HH_ = H_, CPL_ = 1; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
        case 2: { constexpr int HH_ = H_, CPL_ = 2; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
        default: { constexpr int HH_ = H_, CPL_ = 3; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
    }
#define GLAM_VEC_DISPATCH(heads, use_ep, cpl, ...)                      \
    if (!(use_ep)) { GLAM_VEC_CPL(1, false, cpl, __VA_ARGS__) }         \
    else switch (heads) {                                               \
        case 1: GLAM_VEC_CPL(1, true, cpl, __VA_ARGS__) break;          \
        case 2: GLAM_VEC_CPL(2, true, cpl, __VA_ARGS__) break;          \
        case 3: GLAM_VEC_CPL(3, true, cpl, __VA_ARGS__) break;          \
        default: GLAM_VEC_CPL(4, true, cpl, __VA_ARGS__) break;         \
    }
#define GLAM_VEC_HEADS(heads, ...)                                      \
    switch (heads) {                                                    \
        case 1: { constexpr int HH_ = 1; __VA_ARGS__; } break;          \
        case 2: { constexpr int HH_ = 2; __VA_ARGS__; } break;          \
        case 3: { constexpr int HH_ = 3; __VA_ARGS__; } break;          \
        default: { constexpr int HH_ = 4; __VA_ARGS__; } break;         \
    }

template <typename F>
static void vec_allow_smem(F fn, size_t bytes) {
    if (bytes > 48 * 1024) ensure_dyn_smem((const void*)fn, (size_t)((int)bytes));
}

int edge_vec_fwd(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge,
                 const int32_t* rowptr, const int32_t* srcs, int64_t N, int heads, int C, int De, float slope, float* agg,
                 float* alpha, cudaStream_t stream) {
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4;
    const size_t smem = use_ep ? sizeof(float4) * De * nq : 0;
    int G, gcpl;
    vec_geometry(nq, &G, &gcpl);
    const int64_t ntiles = (N + kTileDst - 1) / kTileDst;
    int64_t grid = (int64_t)kNumSMs * 4;
    if (grid > ntiles) grid = ntiles;
    GLAM_VEC_GDISPATCH(heads, use_ep, G, {
        auto fn = edge_tile_fwd_kernel<HH_, G_, CPL_, UE_>;
        vec_allow_smem(fn, smem);
        fn<<<(unsigned)grid, kVecThreads, smem, stream>>>(xpe, ldxp, ea, w_edge, att_edge, rowptr, srcs, N, C, De, slope, agg, alpha);
    })
    GLAM_CHECK_LAUNCH();
    return 0;
}

size_t edge_vec_bwd_workspace(int heads, int C, int De) { return sizeof(float) * (size_t)kNumSMs * 8 * De * heads * C; }

int edge_vec_bwd_dst(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge,
                     const float* alpha, const float* g_agg, const int32_t* rowptr, const int32_t* srcs, const int32_t* dsts,
                     int64_t N, int64_t E, int heads, int C, int De, float slope, float* g_logit, float* g_xpe, float* g_w_edge,
                     void* workspace, cudaStream_t stream, int* grid_out) {
    if (dsts != nullptr && De <= kEpMaxDe) {
        const bool use_ep_t = w_edge != nullptr;
        const int HCt = heads * C, nq_t = HCt / 4;
        int G, gcpl;
        vec_geometry(nq_t, &G, &gcpl);
        int tgrid = kNumSMs * 4;
        if (E > 0) {
            GLAM_VEC_GDISPATCH(heads, use_ep_t, G, {
                const size_t smem_t = use_ep_t ? sizeof(float4) * (size_t)(De * nq_t) * (1 + kVecThreads / G_) : 0;
                auto fn = edge_dots_ep_kernel<HH_, G_, CPL_, UE_>;
                vec_allow_smem(fn, smem_t);
                fn<<<tgrid, kVecThreads, smem_t, stream>>>(xpe, ldxp, ea, w_edge, alpha, g_agg, srcs, dsts, E, C, De, g_logit, (float*)workspace);
            })
            GLAM_CHECK_LAUNCH();
        } else if (use_ep_t) {
            cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)tgrid * De * HCt, stream);
        }
        GLAM_VEC_HEADS(heads, {
            edge_softmax_bwd_kernel<HH_><<<vec_thread_grid(N), 256, 0, stream>>>(xpe, ldxp, ea, att_edge, alpha, rowptr, srcs, N, HCt, De, slope,
                                                                              g_logit, g_xpe);
        })
        GLAM_CHECK_LAUNCH();
        *grid_out = tgrid;
        return 0;
    }
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4, cpl = (nq + 31) / 32;
    int grid = vec_warp_grid(N);
    if (grid > kNumSMs * 3) grid = kNumSMs * 3;        // one g_w_edge partial per CTA: keep the reduction small
    const size_t smem = use_ep ? sizeof(float4) * (size_t)(De * nq) * (1 + kVecWarps) : 0;
    GLAM_VEC_DISPATCH(heads, use_ep, cpl, {
        auto fn = edge_dots_bwd_kernel<HH_, CPL_, UE_>;
        vec_allow_smem(fn, smem);
        fn<<<grid, kVecThreads, smem, stream>>>(xpe, ldxp, ea, w_edge, alpha, g_agg, rowptr, srcs, N, C, De, g_logit, (float*)workspace);
    })
    GLAM_CHECK_LAUNCH();
    GLAM_VEC_HEADS(heads, {
        edge_softmax_bwd_kernel<HH_><<<vec_thread_grid(N), 256, 0, stream>>>(xpe, ldxp, ea, att_edge, alpha, rowptr, srcs, N, HC, De, slope,
                                                                          g_logit, g_xpe);
    })
    GLAM_CHECK_LAUNCH();
    *grid_out = grid;
    (void)g_w_edge;
    return 0;
}

int edge_vec_bwd_src(const float* ea, const float* w_edge, const float* alpha, const float* g_agg, const float* g_logit,
                     const int32_t* src_rowptr, const int32_t* src_pos, const int32_t* src_dst, int64_t N, int heads, int C,
                     int De, float* g_xpe, int64_t ldxp, cudaStream_t stream) {
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4, cpl = (nq + 31) / 32;
    const size_t smem = use_ep ? sizeof(float4) * De * nq : 0;
    int G, gcpl;
    vec_geometry(nq, &G, &gcpl);
    GLAM_VEC_GDISPATCH(heads, use_ep, G, {
        auto fn = edge_source_bwd_kernel<HH_, G_, CPL_, UE_>;
        vec_allow_smem(fn, smem);
        fn<<<vec_warp_grid((N + 32 / G_ - 1) / (32 / G_)), kVecThreads, smem, stream>>>(ea, w_edge, alpha, g_agg, g_logit, src_rowptr, src_pos,
                                                                                        src_dst, N, C, De, g_xpe, ldxp);
    })
    GLAM_CHECK_LAUNCH();
    return 0;
}

}  // namespace glam
