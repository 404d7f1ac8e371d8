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
template <int H, int CPL, bool USE_EP>
__global__ void __launch_bounds__(kVecThreads)
edge_aggregate_fwd_kernel(const float* __restrict__ xpe, int64_t ld, const float* __restrict__ ea, const float* __restrict__ w_edge,
                          const float* __restrict__ alpha, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs,
                          int64_t N, int C, int De, float* __restrict__ agg) {
    extern __shared__ float4 We4[];                    // [De][HC/4]
    const int HC = H * C, nq = HC >> 2;
    if (USE_EP)
        for (int i = threadIdx.x; i < De * nq; i += blockDim.x) We4[i] = ldg4(w_edge + 4 * i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (lane + 32 * t)) / C;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        float4 acc[CPL];
#pragma unroll
        for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = beg; p < end; ++p) {
            const int64_t j = srcs[p];
            float a[H];
#pragma unroll
            for (int h = 0; h < H; ++h) a[h] = alpha[(int64_t)p * H + h];
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = lane + 32 * t;
                if (q < nq) {
                    float4 m = ldg4(xpe + j * ld + 4 * q);
                    if (USE_EP) {
                        float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int d = 0; d < De; ++d) {
                            const float ed = ea[(int64_t)p * De + d];
                            if (ed != 0.f) ep = f4fma(ed, We4[d * nq + q], ep);
                        }
                        m = f4mul(m, ep);
                    }
                    acc[t] = f4fma(pickh<H>(a, hq[t]), m, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = lane + 32 * t;
            if (q < nq) *reinterpret_cast<float4*>(agg + i * HC + 4 * q) = acc[t];
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
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = lane + 32 * t;
                if (q < nq) {
                    const float4 gm = f4mul(ga[t], ldg4(xpe + j * ld + 4 * q));
                    float v;
                    if (USE_EP) {
                        float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float ah = pickh<H>(a, hq[t]);
#pragma unroll
                        for (int d = 0; d < kVecMaxDe; ++d) {
                            if (d < De) {
                                const float ed = ea[(int64_t)p * De + d];
                                if (ed != 0.f) {
                                    ep = f4fma(ed, We4[d * nq + q], ep);
                                    gw[d][t] = f4fma(ed * ah, gm, gw[d][t]);
                                }
                            }
                        }
                        v = f4dot(gm, ep);
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
template <int H, int CPL, bool USE_EP>
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
    const int lane = threadIdx.x & 31;
    int hq[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) hq[t] = (4 * (lane + 32 * t)) / C;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp0; j < N; j += nwarps) {
        const int beg = src_rowptr[j], end = src_rowptr[j + 1];
        float4 acc[CPL];
#pragma unroll
        for (int t = 0; t < CPL; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        float gsj = 0.f;
        for (int k = beg; k < end; ++k) {
            const int p = src_pos[k];
            const int64_t i = src_dst[k];
            float a[H];
#pragma unroll
            for (int h = 0; h < H; ++h) a[h] = alpha[(int64_t)p * H + h];
            if (lane < H) gsj += g_logit[(int64_t)p * H + lane];
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int q = lane + 32 * t;
                if (q < nq) {
                    float4 m = ldg4(g_agg + i * HC + 4 * q);
                    if (USE_EP) {
                        float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int d = 0; d < De; ++d) {
                            const float ed = ea[(int64_t)p * De + d];
                            if (ed != 0.f) ep = f4fma(ed, We4[d * nq + q], ep);
                        }
                        m = f4mul(m, ep);
                    }
                    acc[t] = f4fma(pickh<H>(a, hq[t]), m, acc[t]);
                }
            }
        }
        float* out = g_xpe + j * ld;
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int q = lane + 32 * t;
            if (q < nq) *reinterpret_cast<float4*>(out + 4 * q) = acc[t];
        }
        if (lane < H) out[HC + H + lane] = gsj;
        for (int k = HC + 2 * H + lane; k < ld; k += 32) out[k] = 0.f;
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
    if (bytes > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int edge_vec_fwd(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge,
                 const int32_t* rowptr, const int32_t* srcs, int64_t N, int heads, int C, int De, float slope, float* agg,
                 float* alpha, cudaStream_t stream) {
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4, cpl = (nq + 31) / 32;
    GLAM_VEC_HEADS(heads, {
        edge_alpha_fwd_kernel<HH_><<<vec_thread_grid(N), 256, 0, stream>>>(xpe, ldxp, ea, att_edge, rowptr, srcs, N, HC, De, slope, alpha);
    })
    GLAM_CHECK_LAUNCH();
    const size_t smem = use_ep ? sizeof(float4) * De * nq : 0;
    GLAM_VEC_DISPATCH(heads, use_ep, cpl, {
        auto fn = edge_aggregate_fwd_kernel<HH_, CPL_, UE_>;
        vec_allow_smem(fn, smem);
        fn<<<vec_warp_grid(N), kVecThreads, smem, stream>>>(xpe, ldxp, ea, w_edge, alpha, rowptr, srcs, N, C, De, agg);
    })
    GLAM_CHECK_LAUNCH();
    return 0;
}

size_t edge_vec_bwd_workspace(int heads, int C, int De) { return sizeof(float) * (size_t)kNumSMs * 8 * De * heads * C; }

int edge_vec_bwd_dst(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge,
                     const float* alpha, const float* g_agg, const int32_t* rowptr, const int32_t* srcs, int64_t N, int heads,
                     int C, int De, float slope, float* g_logit, float* g_xpe, float* g_w_edge, void* workspace,
                     cudaStream_t stream, int* grid_out) {
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
    GLAM_VEC_DISPATCH(heads, use_ep, cpl, {
        auto fn = edge_source_bwd_kernel<HH_, CPL_, UE_>;
        vec_allow_smem(fn, smem);
        fn<<<vec_warp_grid(N), kVecThreads, smem, stream>>>(ea, w_edge, alpha, g_agg, g_logit, src_rowptr, src_pos, src_dst, N, C, De,
                                                            g_xpe, ldxp);
    })
    GLAM_CHECK_LAUNCH();
    return 0;
}

}  // namespace glam
