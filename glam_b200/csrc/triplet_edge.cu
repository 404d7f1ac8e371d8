// Triplet attention edge phase (see include/glam_b200.h (3); math in SURVEY.md Appendix C).
//
// One warp per destination node over the dst-sorted CSR; lanes run over the H*C message channels, so the
// gathers of a source row xp[j, 0:HC] are fully coalesced.  The per-destination softmax and the weighted
// segment sum happen in registers: no [E,H,3C] triplet tensor, no [E,HC] message tensor, no atomics.
// Gradients that reduce over all edges (g_w_edge) go through per-CTA partials + a fixed-order final pass.
#include "common.cuh"

namespace glam {

constexpr int kEdgeWarps = 8;                 // warps per CTA
constexpr int kEdgeThreads = kEdgeWarps * 32;
constexpr int kEdgeCtasPerSM = 4;

template <int H>
__device__ __forceinline__ float pick(const float (&a)[H], int h) {
    float v = a[0];
#pragma unroll
    for (int q = 1; q < H; ++q) v = (h == q) ? a[q] : v;
    return v;
}

__device__ __forceinline__ float leaky(float x, float s) { return x > 0.f ? x : s * x; }

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
template <int H, int KPL, bool USE_EP>
__global__ void __launch_bounds__(kEdgeThreads)
triplet_edge_fwd_kernel(const float* __restrict__ xpe, int64_t ldxp, const float* __restrict__ ea,
                        const float* __restrict__ w_edge, const float* __restrict__ att_edge,
                        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs, int64_t N, int C, int De,
                        float slope, float* __restrict__ agg, float* __restrict__ alpha) {
    extern __shared__ float smem[];
    const int HC = H * C;
    float* We_s = smem;                              // [De][HC]  (USE_EP only)
    float* Ae_s = smem + (USE_EP ? De * HC : 0);     // [De][H]
    for (int i = threadIdx.x; i < (USE_EP ? De * HC : 0); i += blockDim.x) We_s[i] = w_edge[i];
    for (int i = threadIdx.x; i < De * H; i += blockDim.x) Ae_s[i] = att_edge[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    int hidx[KPL];
#pragma unroll
    for (int t = 0; t < KPL; ++t) hidx[t] = (lane + 32 * t) / C;

    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int beg = rowptr[i], end = rowptr[i + 1], deg = end - beg;
        float* out = agg + i * HC;
        if (deg == 0) {
#pragma unroll
            for (int t = 0; t < KPL; ++t) { int k = lane + 32 * t; if (k < HC) out[k] = 0.f; }
            continue;
        }
        const bool small = deg <= 32;                // warp-uniform
        float si[H];
#pragma unroll
        for (int h = 0; h < H; ++h) si[h] = xpe[i * ldxp + HC + h];

        // ---- logits + running max (lane-parallel over edges)
        float lg[H], mx[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { mx[h] = -INFINITY; lg[h] = 0.f; }
        for (int p = beg + lane; p < end; p += 32) {
            const int64_t j = srcs[p];
            float l[H];
#pragma unroll
            for (int h = 0; h < H; ++h) l[h] = si[h] + xpe[j * ldxp + HC + H + h];
            for (int d = 0; d < De; ++d) {
                float e = ea[(int64_t)p * De + d];
#pragma unroll
                for (int h = 0; h < H; ++h) l[h] = fmaf(e, Ae_s[d * H + h], l[h]);
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                l[h] = leaky(l[h], slope);
                mx[h] = fmaxf(mx[h], l[h]);
                if (small) lg[h] = l[h]; else alpha[(int64_t)p * H + h] = l[h];
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) mx[h] = warp_max(mx[h]);
        // ---- exp + sum
        float sum[H];
#pragma unroll
        for (int h = 0; h < H; ++h) sum[h] = 0.f;
        for (int p = beg + lane; p < end; p += 32) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float l = small ? lg[h] : alpha[(int64_t)p * H + h];
                float e = expf(l - mx[h]);
                sum[h] += e;
                if (small) lg[h] = e; else alpha[(int64_t)p * H + h] = e;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) sum[h] = warp_sum(sum[h]) + 1e-16f;
        // ---- normalise (PyG: out / (out_sum + 1e-16)) and publish alpha
        for (int p = beg + lane; p < end; p += 32) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float e = small ? lg[h] : alpha[(int64_t)p * H + h];
                float a = e / sum[h];
                if (small) lg[h] = a;
                alpha[(int64_t)p * H + h] = a;
            }
        }
        __syncwarp();
        // ---- weighted segment sum: lanes over channels, edges in CSR (= stable) order
        float acc[KPL];
#pragma unroll
        for (int t = 0; t < KPL; ++t) acc[t] = 0.f;
        for (int e = 0; e < deg; ++e) {
            const int p = beg + e;
            float a[H];
            if (small) {
#pragma unroll
                for (int h = 0; h < H; ++h) a[h] = __shfl_sync(0xffffffffu, lg[h], e);
            } else {
#pragma unroll
                for (int h = 0; h < H; ++h) a[h] = alpha[(int64_t)p * H + h];
            }
            const int64_t j = srcs[p];
            const float* xj = xpe + j * ldxp;
            float ep[KPL];
            if (USE_EP) {
#pragma unroll
                for (int t = 0; t < KPL; ++t) ep[t] = 0.f;
                for (int d = 0; d < De; ++d) {
                    float ed = ea[(int64_t)p * De + d];
                    if (ed != 0.f) {
#pragma unroll
                        for (int t = 0; t < KPL; ++t) {
                            int k = lane + 32 * t;
                            if (k < HC) ep[t] = fmaf(ed, We_s[d * HC + k], ep[t]);
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                int k = lane + 32 * t;
                if (k < HC) {
                    float m = xj[k];
                    if (USE_EP) m *= ep[t];
                    acc[t] = fmaf(pick<H>(a, hidx[t]), m, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < KPL; ++t) { int k = lane + 32 * t; if (k < HC) out[k] = acc[t]; }
    }
}

// ---------------------------------------------------------------------------------------------------
// backward, destination pass
// ---------------------------------------------------------------------------------------------------
template <int H, int KPL, bool USE_EP>
__global__ void __launch_bounds__(kEdgeThreads)
triplet_edge_bwd_dst_kernel(const float* __restrict__ xpe, int64_t ldxp, const float* __restrict__ ea,
                            const float* __restrict__ w_edge, const float* __restrict__ att_edge,
                            const float* __restrict__ alpha, const float* __restrict__ g_agg,
                            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ srcs, int64_t N, int C,
                            int De, float slope, float* __restrict__ g_logit, float* __restrict__ g_xpe,
                            float* __restrict__ gwe_partial) {
    extern __shared__ float smem[];
    const int HC = H * C;
    float* We_s = smem;
    float* Ae_s = smem + (USE_EP ? De * HC : 0);
    float* gw_s = Ae_s + De * H;                     // [kEdgeWarps][De][HC] (USE_EP only)
    for (int i = threadIdx.x; i < (USE_EP ? De * HC : 0); i += blockDim.x) We_s[i] = w_edge[i];
    for (int i = threadIdx.x; i < De * H; i += blockDim.x) Ae_s[i] = att_edge[i];
    for (int i = threadIdx.x; i < (USE_EP ? kEdgeWarps * De * HC : 0); i += blockDim.x) gw_s[i] = 0.f;
    __syncthreads();

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* gw = gw_s + wid * De * HC;
    int hidx[KPL];
#pragma unroll
    for (int t = 0; t < KPL; ++t) hidx[t] = (lane + 32 * t) / C;

    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        if (beg == end) {
            if (lane < H) g_xpe[i * ldxp + HC + lane] = 0.f;
            continue;
        }
        float ga[KPL];
#pragma unroll
        for (int t = 0; t < KPL; ++t) { int k = lane + 32 * t; ga[t] = k < HC ? g_agg[i * HC + k] : 0.f; }
        float si[H], dot[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { si[h] = xpe[i * ldxp + HC + h]; dot[h] = 0.f; }

        // ---- pass 1 (lanes over channels): g_alpha[p,h] = <g_agg[i,h,:], e_ij (.) x_j>, g_w_edge partials
        for (int p = beg; p < end; ++p) {
            const int64_t j = srcs[p];
            const float* xj = xpe + j * ldxp;
            float a[H], part[H];
#pragma unroll
            for (int h = 0; h < H; ++h) { a[h] = alpha[(int64_t)p * H + h]; part[h] = 0.f; }
            float gm[KPL], ep[KPL];
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                int k = lane + 32 * t;
                gm[t] = k < HC ? ga[t] * xj[k] : 0.f;
                ep[t] = USE_EP ? 0.f : 1.f;
            }
            if (USE_EP) {
                for (int d = 0; d < De; ++d) {
                    float ed = ea[(int64_t)p * De + d];
                    if (ed != 0.f) {
#pragma unroll
                        for (int t = 0; t < KPL; ++t) {
                            int k = lane + 32 * t;
                            if (k < HC) {
                                ep[t] = fmaf(ed, We_s[d * HC + k], ep[t]);
                                gw[d * HC + k] = fmaf(ed, pick<H>(a, hidx[t]) * gm[t], gw[d * HC + k]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                float v = gm[t] * ep[t];
#pragma unroll
                for (int h = 0; h < H; ++h) part[h] += (hidx[t] == h) ? v : 0.f;
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                part[h] = warp_sum(part[h]);
                dot[h] = fmaf(a[h], part[h], dot[h]);
            }
            if (lane < H) g_logit[(int64_t)p * H + lane] = pick<H>(part, lane);
        }
        __syncwarp();
        // ---- pass 2 (lanes over edges): softmax + leaky_relu backward
        float gsi[H];
#pragma unroll
        for (int h = 0; h < H; ++h) gsi[h] = 0.f;
        for (int p = beg + lane; p < end; p += 32) {
            const int64_t j = srcs[p];
            float l[H];
#pragma unroll
            for (int h = 0; h < H; ++h) l[h] = si[h] + xpe[j * ldxp + HC + H + h];
            for (int d = 0; d < De; ++d) {
                float e = ea[(int64_t)p * De + d];
#pragma unroll
                for (int h = 0; h < H; ++h) l[h] = fmaf(e, Ae_s[d * H + h], l[h]);
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float a = alpha[(int64_t)p * H + h];
                float g = a * (g_logit[(int64_t)p * H + h] - dot[h]);
                g *= (l[h] > 0.f ? 1.f : slope);
                g_logit[(int64_t)p * H + h] = g;
                gsi[h] += g;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) gsi[h] = warp_sum(gsi[h]);
        if (lane < H) g_xpe[i * ldxp + HC + lane] = pick<H>(gsi, lane);
        __syncwarp();
    }
    if (USE_EP) {
        __syncthreads();
        float* P = gwe_partial + (int64_t)blockIdx.x * De * HC;
        for (int idx = threadIdx.x; idx < De * HC; idx += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kEdgeWarps; ++w) s += gw_s[w * De * HC + idx];
            P[idx] = s;
        }
    }
}

// out[i] = sum_s partial[s][i]: 64 outputs per CTA, the S range split over 4 thread rows combined in a fixed order
__global__ void __launch_bounds__(256)
reduce_cta_partials_kernel(const float* __restrict__ partial, int S, int count, float* __restrict__ out) {
    __shared__ float red[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + tx;
    float s = 0.f;
    if (i < count) {
        const int per = (S + 3) / 4, k0 = ty * per, k1 = min(S, k0 + per);
        for (int k = k0; k < k1; ++k) s += partial[(int64_t)k * count + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < count) out[i] = ((red[0][tx] + red[1][tx]) + red[2][tx]) + red[3][tx];
}

// ---------------------------------------------------------------------------------------------------
// backward, source pass
// ---------------------------------------------------------------------------------------------------
template <int H, int KPL, bool USE_EP>
__global__ void __launch_bounds__(kEdgeThreads)
triplet_edge_bwd_src_kernel(const float* __restrict__ ea, const float* __restrict__ w_edge, const float* __restrict__ alpha,
                            const float* __restrict__ g_agg, const float* __restrict__ g_logit,
                            const int32_t* __restrict__ src_rowptr, const int32_t* __restrict__ src_pos,
                            const int32_t* __restrict__ src_dst, int64_t N, int C, int De, float* __restrict__ g_xpe,
                            int64_t ldxp) {
    extern __shared__ float smem[];
    const int HC = H * C;
    float* We_s = smem;
    for (int i = threadIdx.x; i < (USE_EP ? De * HC : 0); i += blockDim.x) We_s[i] = w_edge[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int hidx[KPL];
#pragma unroll
    for (int t = 0; t < KPL; ++t) hidx[t] = (lane + 32 * t) / C;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp0; j < N; j += nwarps) {
        const int beg = src_rowptr[j], end = src_rowptr[j + 1];
        float acc[KPL], gsj[H];
#pragma unroll
        for (int t = 0; t < KPL; ++t) acc[t] = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) gsj[h] = 0.f;
        for (int q = beg; q < end; ++q) {
            const int p = src_pos[q];
            const int64_t i = src_dst[q];
            float a[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                a[h] = alpha[(int64_t)p * H + h];
                gsj[h] += g_logit[(int64_t)p * H + h];
            }
            float ep[KPL];
            if (USE_EP) {
#pragma unroll
                for (int t = 0; t < KPL; ++t) ep[t] = 0.f;
                for (int d = 0; d < De; ++d) {
                    float ed = ea[(int64_t)p * De + d];
                    if (ed != 0.f) {
#pragma unroll
                        for (int t = 0; t < KPL; ++t) {
                            int k = lane + 32 * t;
                            if (k < HC) ep[t] = fmaf(ed, We_s[d * HC + k], ep[t]);
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                int k = lane + 32 * t;
                if (k < HC) {
                    float m = g_agg[i * HC + k];
                    if (USE_EP) m *= ep[t];
                    acc[t] = fmaf(pick<H>(a, hidx[t]), m, acc[t]);
                }
            }
        }
        float* out = g_xpe + j * ldxp;
#pragma unroll
        for (int t = 0; t < KPL; ++t) { int k = lane + 32 * t; if (k < HC) out[k] = acc[t]; }
        if (lane < H) out[HC + H + lane] = pick<H>(gsj, lane);
        for (int k = HC + 2 * H + lane; k < ldxp; k += 32) out[k] = 0.f;
    }
}

static int edge_grid(int64_t N) {
    int64_t want = (N + kEdgeWarps - 1) / kEdgeWarps;
    int64_t cap = (int64_t)kNumSMs * kEdgeCtasPerSM;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

static int pick_kpl(int HC) {
    if (HC <= 64) return 2;
    if (HC <= 128) return 4;
    if (HC <= 192) return 6;
    if (HC <= 288) return 9;
    return 0;
}

template <typename F>
static cudaError_t allow_smem(F fn, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return ensure_dyn_smem((const void*)fn, (size_t)((int)bytes));
}

bool edge_vec_eligible(const float* xpe, int64_t ldxp, int heads, int C, int De, const float* a1, const float* a2);
int edge_vec_fwd(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge,
                 const int32_t* rowptr, const int32_t* srcs, int64_t N, int heads, int C, int De, float slope, float* agg,
                 float* alpha, cudaStream_t stream);
size_t edge_vec_bwd_workspace(int heads, int C, int De);
int edge_vec_bwd_dst(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge,
                     const float* alpha, const float* g_agg, const int32_t* rowptr, const int32_t* srcs, const int32_t* dsts, int64_t N, int64_t E, int heads,
                     int C, int De, float slope, float* g_logit, float* g_xpe, float* g_w_edge, void* workspace,
                     cudaStream_t stream, int* grid_out);
int edge_vec_bwd_src(const float* ea, const float* w_edge, const float* alpha, const float* g_agg, const float* g_logit,
                     const int32_t* src_rowptr, const int32_t* src_pos, const int32_t* src_dst, int64_t N, int heads, int C,
                     int De, float* g_xpe, int64_t ldxp, cudaStream_t stream);

bool edge_win_eligible(int heads, int C, int De, int64_t ldxp);
int edge_tile_rows(int64_t N);
int64_t edge_tile_count(int64_t N);
int edge_win_build_tiles(const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* src_rowptr, const int32_t* src_dst,
                         int64_t N, int32_t* dst_tiles, int32_t* src_tiles, cudaStream_t stream);
int edge_win_fwd(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge, const int32_t* rowptr,
                 const int32_t* srcs, const int32_t* tiles, int64_t N, int heads, int C, int De, float slope, float* agg, float* alpha,
                 cudaStream_t stream, int* launched);
int edge_win_bwd_dst_grid(int64_t N);
int edge_win_bwd_dst(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge, const float* alpha,
                     const float* g_agg, const int32_t* rowptr, const int32_t* srcs, const int32_t* tiles, int64_t N, int heads, int C,
                     int De, float slope, float* g_logit, float* g_xpe, float* gwe_partial, cudaStream_t stream, int* launched);
int edge_win_bwd_src(const float* ea, const float* w_edge, const float* alpha, const float* g_agg, const float* g_logit,
                     const int32_t* src_rowptr, const int32_t* src_pos, const int32_t* src_dst, const int32_t* tiles, int64_t N, int heads,
                     int C, int De, float* g_xpe, int64_t ldxp, cudaStream_t stream, int* launched);

}  // namespace glam

using namespace glam;

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

extern "C" int glam_edge_tile_rows(int64_t num_nodes) { return edge_tile_rows(num_nodes); }
extern "C" int64_t glam_edge_tile_count(int64_t num_nodes) { return edge_tile_count(num_nodes); }
extern "C" int glam_build_edge_tiles(const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* src_rowptr,
                                     const int32_t* src_dst, int64_t N, int64_t E, int32_t* dst_tiles, int32_t* src_tiles,
                                     void* stream) {
    if (N == 0) return 0;
    GLAM_REQUIRE(dst_rowptr && dst_tiles && (E == 0 || dst_src), "glam_build_edge_tiles: null pointer");
    GLAM_REQUIRE(!src_tiles || (src_rowptr && (E == 0 || src_dst)), "glam_build_edge_tiles: src_tiles needs src_rowptr/src_dst");
    GLAM_REQUIRE(aligned16(dst_tiles) && aligned16(src_tiles), "glam_build_edge_tiles: tile arrays must be 16-byte aligned");
    return edge_win_build_tiles(dst_rowptr, dst_src, src_rowptr, src_dst, N, dst_tiles, src_tiles, (cudaStream_t)stream);
}

#define GLAM_KPL_CASE(H_, EP_, K_, ...) \
    { constexpr int KPL_ = K_; constexpr int HH_ = H_; constexpr bool UE_ = EP_; __VA_ARGS__; }
#define GLAM_DISPATCH_KPL(H_, EP_, kpl, ...)                      \
    switch (kpl) {                                                \
        case 2: GLAM_KPL_CASE(H_, EP_, 2, __VA_ARGS__) break;     \
        case 4: GLAM_KPL_CASE(H_, EP_, 4, __VA_ARGS__) break;     \
        case 6: GLAM_KPL_CASE(H_, EP_, 6, __VA_ARGS__) break;     \
        default: GLAM_KPL_CASE(H_, EP_, 9, __VA_ARGS__) break;    \
    }
#define GLAM_DISPATCH_EDGE(heads, use_ep, kpl, ...)                        \
    if (!(use_ep)) { GLAM_DISPATCH_KPL(1, false, kpl, __VA_ARGS__) }       \
    else switch (heads) {                                                  \
        case 1: GLAM_DISPATCH_KPL(1, true, kpl, __VA_ARGS__) break;        \
        case 2: GLAM_DISPATCH_KPL(2, true, kpl, __VA_ARGS__) break;        \
        case 3: GLAM_DISPATCH_KPL(3, true, kpl, __VA_ARGS__) break;        \
        default: GLAM_DISPATCH_KPL(4, true, kpl, __VA_ARGS__) break;       \
    }

static int check_edge_args(const char* fn, int heads, int channels, int edge_dim, bool use_ep, int64_t ldxp) {
    GLAM_REQUIRE(heads >= 1 && heads <= GLAM_MAX_HEADS, "%s: heads must be in [1,%d], got %d", fn, GLAM_MAX_HEADS, heads);
    GLAM_REQUIRE(use_ep || heads == 1, "%s: the Light layer (w_edge == NULL) is single-head", fn);
    GLAM_REQUIRE(channels >= 1 && edge_dim >= 1 && edge_dim <= 64, "%s: bad channels/edge_dim", fn);
    GLAM_REQUIRE(pick_kpl(heads * channels) != 0, "%s: heads*channels = %d exceeds 288", fn, heads * channels);
    GLAM_REQUIRE(ldxp >= (int64_t)heads * channels + 2 * heads, "%s: ldxp too small", fn);
    return 0;
}

extern "C" int glam_triplet_edge_fwd(const float* xpe, int64_t ldxp, const float* edge_attr, const float* w_edge,
                                     const float* att_edge, const int32_t* dst_rowptr, const int32_t* dst_src,
                                     const int32_t* dst_tiles, int64_t N, int64_t E, int heads, int C, int De, float slope,
                                     float* agg, float* alpha, void* stream_) {
    const bool use_ep = w_edge != nullptr;
    if (int rc = check_edge_args("glam_triplet_edge_fwd", heads, C, De, use_ep, ldxp)) return rc;
    if (N == 0) return 0;
    GLAM_REQUIRE(xpe && att_edge && dst_rowptr && agg && (E == 0 || (edge_attr && dst_src && alpha)),
                 "glam_triplet_edge_fwd: null pointer");
    const int HC = heads * C, kpl = pick_kpl(HC);
    const size_t smem = sizeof(float) * ((use_ep ? De * HC : 0) + De * heads);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (dst_tiles && E > 0 && edge_win_eligible(heads, C, De, ldxp) && aligned16(xpe) && aligned16(agg) && aligned16(w_edge) &&
        aligned16(dst_tiles)) {
        int launched = 0;
        if (int rc = edge_win_fwd(xpe, ldxp, edge_attr, w_edge, att_edge, dst_rowptr, dst_src, dst_tiles, N, heads, C, De, slope, agg,
                                  alpha, stream, &launched)) return rc;
        if (launched) return 0;
    }
    if (edge_vec_eligible(xpe, ldxp, heads, C, De, agg, w_edge))
        return edge_vec_fwd(xpe, ldxp, edge_attr, w_edge, att_edge, dst_rowptr, dst_src, N, heads, C, De, slope, agg, alpha, stream);
    GLAM_DISPATCH_EDGE(heads, use_ep, kpl, {
        auto fn = triplet_edge_fwd_kernel<HH_, KPL_, UE_>;
        allow_smem(fn, smem);
        fn<<<edge_grid(N), kEdgeThreads, smem, stream>>>(xpe, ldxp, edge_attr, w_edge, att_edge, dst_rowptr, dst_src, N, C, De,
                                                         slope, agg, alpha);
    })
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t glam_triplet_bwd_workspace_bytes(int heads, int channels, int edge_dim) {
    size_t a = sizeof(float) * (size_t)kNumSMs * kEdgeCtasPerSM * (size_t)edge_dim * heads * channels;
    size_t b = edge_vec_bwd_workspace(heads, channels, edge_dim);
    return a > b ? a : b;
}

extern "C" int glam_triplet_edge_bwd_dst(const float* xpe, int64_t ldxp, const float* edge_attr, const float* w_edge,
                                         const float* att_edge, const float* alpha, const float* g_agg,
                                         const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* dst_dst,
                                         const int32_t* dst_tiles, int64_t N, int64_t E,
                                         int heads, int C, int De, float slope, float* g_logit, float* g_xpe,
                                         float* g_w_edge, void* workspace, size_t workspace_bytes, void* stream_) {
    const bool use_ep = w_edge != nullptr;
    if (int rc = check_edge_args("glam_triplet_edge_bwd_dst", heads, C, De, use_ep, ldxp)) return rc;
    const int HC = heads * C, kpl = pick_kpl(HC);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N == 0) {
        if (use_ep && g_w_edge) cudaMemsetAsync(g_w_edge, 0, sizeof(float) * De * HC, stream);
        return 0;
    }
    GLAM_REQUIRE(xpe && att_edge && g_agg && dst_rowptr && g_xpe && (E == 0 || (edge_attr && dst_src && alpha && g_logit)),
                 "glam_triplet_edge_bwd_dst: null pointer");
    GLAM_REQUIRE(!use_ep || (g_w_edge && workspace && workspace_bytes >= glam_triplet_bwd_workspace_bytes(heads, C, De)),
                 "glam_triplet_edge_bwd_dst: g_w_edge/workspace missing or too small");
    if (dst_tiles && E > 0 && edge_win_eligible(heads, C, De, ldxp) && aligned16(xpe) && aligned16(g_agg) && aligned16(w_edge) &&
        aligned16(dst_tiles) && aligned16(workspace)) {
        int launched = 0;
        if (int rc = edge_win_bwd_dst(xpe, ldxp, edge_attr, w_edge, att_edge, alpha, g_agg, dst_rowptr, dst_src, dst_tiles, N, heads, C, De,
                                      slope, g_logit, g_xpe, (float*)workspace, stream, &launched)) return rc;
        if (launched) {
            if (use_ep)
                return launch_reduce_partials((const float*)workspace, edge_win_bwd_dst_grid(N), 1, De * HC, 0, g_w_edge, De * HC, 0,
                                              nullptr, stream);
            return 0;
        }
    }
    if (edge_vec_eligible(xpe, ldxp, heads, C, De, g_agg, w_edge)) {
        int vgrid = 0;
        if (int rc = edge_vec_bwd_dst(xpe, ldxp, edge_attr, w_edge, att_edge, alpha, g_agg, dst_rowptr, dst_src, dst_dst, N, E, heads, C, De,
                                      slope, g_logit, g_xpe, g_w_edge, workspace, stream, &vgrid)) return rc;
        if (use_ep) {
            return launch_reduce_partials((const float*)workspace, vgrid, 1, De * HC, 0, g_w_edge, De * HC, 0, nullptr, stream);
        }
        return 0;
    }
    const size_t smem = sizeof(float) * ((use_ep ? De * HC : 0) + De * heads + (use_ep ? kEdgeWarps * De * HC : 0));
    GLAM_REQUIRE(smem <= 200 * 1024, "glam_triplet_edge_bwd_dst: edge_dim*heads*channels too large for shared memory");
    const int grid = edge_grid(N);
    GLAM_DISPATCH_EDGE(heads, use_ep, kpl, {
        auto fn = triplet_edge_bwd_dst_kernel<HH_, KPL_, UE_>;
        allow_smem(fn, smem);
        fn<<<grid, kEdgeThreads, smem, stream>>>(xpe, ldxp, edge_attr, w_edge, att_edge, alpha, g_agg, dst_rowptr, dst_src, N, C,
                                                 De, slope, g_logit, g_xpe, (float*)workspace);
    })
    GLAM_CHECK_LAUNCH();
    if (use_ep) {
        if (int rc = launch_reduce_partials((const float*)workspace, grid, 1, De * HC, 0, g_w_edge, De * HC, 0, nullptr, stream)) return rc;
    }
    return 0;
}

extern "C" int glam_triplet_edge_bwd_src(const float* edge_attr, const float* w_edge, const float* alpha,
                                         const float* g_agg, const float* g_logit, const int32_t* src_rowptr,
                                         const int32_t* src_pos, const int32_t* src_dst, const int32_t* src_tiles, int64_t N,
                                         int64_t E, int heads, int C, int De, float* g_xpe, int64_t ldxp, void* stream_) {
    const bool use_ep = w_edge != nullptr;
    if (int rc = check_edge_args("glam_triplet_edge_bwd_src", heads, C, De, use_ep, ldxp)) return rc;
    if (N == 0) return 0;
    GLAM_REQUIRE(g_agg && src_rowptr && g_xpe && (E == 0 || (edge_attr && alpha && g_logit && src_pos && src_dst)),
                 "glam_triplet_edge_bwd_src: null pointer");
    const int HC = heads * C, kpl = pick_kpl(HC);
    const size_t smem = sizeof(float) * (use_ep ? De * HC : 0);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (src_tiles && E > 0 && edge_win_eligible(heads, C, De, ldxp) && aligned16(g_xpe) && aligned16(g_agg) && aligned16(w_edge) &&
        aligned16(src_tiles)) {
        int launched = 0;
        if (int rc = edge_win_bwd_src(edge_attr, w_edge, alpha, g_agg, g_logit, src_rowptr, src_pos, src_dst, src_tiles, N, heads, C, De,
                                      g_xpe, ldxp, stream, &launched)) return rc;
        if (launched) return 0;
    }
    if (edge_vec_eligible(g_xpe, ldxp, heads, C, De, g_agg, w_edge))
        return edge_vec_bwd_src(edge_attr, w_edge, alpha, g_agg, g_logit, src_rowptr, src_pos, src_dst, N, heads, C, De, g_xpe, ldxp,
                                stream);
    GLAM_DISPATCH_EDGE(heads, use_ep, kpl, {
        auto fn = triplet_edge_bwd_src_kernel<HH_, KPL_, UE_>;
        allow_smem(fn, smem);
        fn<<<edge_grid(N), kEdgeThreads, smem, stream>>>(edge_attr, w_edge, alpha, g_agg, g_logit, src_rowptr, src_pos, src_dst, N,
                                                         C, De, g_xpe, ldxp);
    })
    GLAM_CHECK_LAUNCH();
    return 0;
}
