// Triplet attention edge phase, WINDOWED path: shared-memory staging of contiguous node windows with the bulk async
// copy engine (cp.async.bulk + mbarrier, SASS: UBLKCP).
//
// Observation: a PyG batch is block-diagonal and its nodes are numbered graph by graph, so all the
// sources of a tile of consecutive destinations lie in ONE short contiguous range of node ids (the graphs the tile
// touches), and the rows xpe[lo:hi) are one contiguous span of HBM.  Instead of gathering x_j per edge through
// L1/L2 (the dependent rowptr -> src -> row chain that kept the gather kernels at 20-30 % of the HBM roofline), a CTA
//   (1) reads its tile descriptor {lo, hi, e0, e1} (glam_build_edge_tiles, once per batch),
//   (2) lets the copy engine stream the whole window into shared memory with a handful of bulk copies — every row
//       is fetched exactly once per tile, fully coalesced, no registers or LSU slots spent on it —
//   (3) meanwhile builds the per-edge records (source slot, bond type, edge part of the logit) from the index arrays,
//   (4) runs softmax (a thread per destination and head) and aggregation (a warp per destination, a lane per
//       16-byte channel chunk) entirely out of shared memory.
// Three CTAs per SM keep two windows in flight while one computes.  Tiles whose window does not fit ("far": proteins,
// very large molecules) keep the records but gather rows from global memory; tiles with more edges than the record
// buffer (hub nodes) take a per-destination routine straight from global memory.  Same arithmetic and summation order
// as triplet_edge_vec.cu; no atomics.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace glam {

constexpr int kWinWarps = 8;
constexpr int kWinThreads = kWinWarps * 32;
constexpr int kWinCtasPerSM = 3;
constexpr int kWinMaxTile = 64;          // nodes per tile (upper bound; the actual size balances the grid, see edge_tile_rows)
constexpr int kWinMaxEdges = 512;        // edge records per tile
constexpr int kWinMaxDe = 8;
constexpr int kWinRegDe = 4;             // g_weight_edge rows kept in registers by the backward-dst kernel
constexpr uint32_t kBulkChunk = 8192;    // bytes per bulk copy
constexpr int kWinSmemBudget = 75 * 1024;   // per CTA: 3 CTAs per SM inside the 227 KB

// nodes per tile: the number of tiles is (just under) a multiple of the resident CTA count, so every CTA gets the same share
int edge_tile_rows(int64_t N) {
    const int64_t ctas = (int64_t)kNumSMs * kWinCtasPerSM;
    int64_t k = (N + ctas * kWinMaxTile - 1) / (ctas * kWinMaxTile);
    if (k < 1) k = 1;
    int64_t D = (N + ctas * k - 1) / (ctas * k);
    if (D < 16) D = 16;
    if (D > kWinMaxTile) D = kWinMaxTile;
    return (int)D;
}
int64_t edge_tile_count(int64_t N) { const int D = edge_tile_rows(N); return (N + D - 1) / D; }

// ------------------------------------------------------------------------------------------------ tile descriptors
// one warp per tile: {first row of the window, one past its last row, first edge, one past the last edge}
__global__ void __launch_bounds__(256)
edge_tiles_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ other, int64_t N, int D, int64_t T, int own_rows,
                  int4* __restrict__ tiles) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    const int64_t t0 = t * D;
    const int nd = (int)min((int64_t)D, N - t0);
    const int e0 = rowptr[t0], e1 = rowptr[t0 + nd];
    int mn = own_rows ? (int)t0 : INT32_MAX, mx = own_rows ? (int)(t0 + nd - 1) : -1;
    for (int e = e0 + lane; e < e1; e += 32) {
        const int j = other[e];
        mn = min(mn, j);
        mx = max(mx, j);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (mx < 0) { mn = 0; mx = -1; }
    if (lane == 0) tiles[t] = make_int4(mn, mx + 1, e0, e1);
}

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
// warp 0: stream `bytes` (multiple of 16, both sides 16-byte aligned) into shared memory; completion on `bar`
__device__ __forceinline__ void bulk_window(float* dst, const float* src, uint32_t bytes, uint64_t* bar, int lane) {
    if (lane == 0) tc::mbar_expect_tx(bar, bytes);
    __syncwarp();
    for (uint32_t off = (uint32_t)lane * kBulkChunk; off < bytes; off += 32u * kBulkChunk)
        bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off, min(kBulkChunk, bytes - off), bar);
}

// e_ij chunk for an edge whose edge_attr row is not one-hot (protein contact features)
__device__ __forceinline__ float4 ep_general(const float* __restrict__ earow, int De, const float4* We4, int nq, int q) {
    float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 0; d < De; ++d) ep = fma4(earow[d], We4[d * nq + q], ep);
    return ep;
}

struct WinArgs {
    const float* xpe; int64_t ld;
    const float* ea; const float* w_edge; const float* att_edge;
    const int32_t* rowptr; const int32_t* other;      // dst pass: dst_rowptr / dst_src;  src pass: src_rowptr / src_dst
    const int4* tiles;
    int64_t N; int C, De, D, rmax; int64_t T;
    float slope;
};

// shared-memory carve-up (all kernels): rows | We4 | records ... ; offsets in floats from a 128-byte aligned base
struct WinSmem {
    float* rows; float4* We4; float* Ae; int* rp; uint64_t* bar; float* rec;
};
__device__ __forceinline__ WinSmem win_carve(float* base, int rmax, int64_t ldrow, int De, int nq, int H, bool use_ep) {
    WinSmem s;
    s.rows = base;
    float* p = base + (size_t)rmax * ldrow;
    s.We4 = reinterpret_cast<float4*>(p);
    p += use_ep ? De * nq * 4 : 0;
    s.Ae = p;
    p += kWinMaxDe * GLAM_MAX_HEADS;
    s.rp = reinterpret_cast<int*>(p);
    p += kWinMaxTile + 4;                       // keeps 16-byte alignment (68 ints)
    s.bar = reinterpret_cast<uint64_t*>(p);
    p += 4;
    s.rec = p;
    return s;
}
static size_t win_smem_bytes(int rmax, int64_t ldrow, int De, int nq, bool use_ep, int rec_floats_per_edge, int extra_floats) {
    return sizeof(float) * ((size_t)rmax * ldrow + (use_ep ? De * nq * 4 : 0) + kWinMaxDe * GLAM_MAX_HEADS + kWinMaxTile + 4 + 4 +
                            (size_t)kWinMaxEdges * rec_floats_per_edge + extra_floats);
}

// ------------------------------------------------------------------------------------------------ work items
// The channel row of a node (HC floats = nq 16-byte chunks) is cut into `ni` items of CPI consecutive chunks, CPI the
// largest of {3,2,1} dividing the chunks per head (C/4), so an item never straddles heads.  The aggregation phases give
// one THREAD per (node, item): consecutive lanes take consecutive items, so a node's row is read and written as one
// contiguous run, lanes of one warp cover ~3 nodes, and the per-edge bookkeeping (record loads, addressing, loop) is
// amortised over CPI chunks instead of being repeated by every lane of a warp.
struct ItemGeom { int ni, tph; };      // items per row, items per head
static int pick_cpi(int C) { const int c4 = C / 4; return c4 % 3 == 0 ? 3 : (c4 % 2 == 0 ? 2 : 1); }

// ------------------------------------------------------------------------------------------------ forward
// records: {src slot, type} | value | c[H]  (edge part of the logit, then alpha * value)
template <int H, bool USE_EP, typename IDX>
__device__ __forceinline__ void fwd_softmax(const float* rows, int ld, IDX own0, int HC, int nd, int e0, float slope, const int* rp,
                                            const int2* rec_st, const float* rec_val, float* rec_c, float* __restrict__ alpha) {
    const int idx = threadIdx.x;
    if (idx >= nd * H) return;
    const int d = idx / H, h = idx - d * H;
    const int beg = rp[d], end = rp[d + 1], deg = end - beg;
    const float si = rows[(own0 + d) * ld + HC + h];
    if (deg <= 4) {
        // molecular graphs: valence-bounded in-degree — everything stays in registers, no loops
        float l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            l[k] = -INFINITY;
            if (k < deg) {
                const int e = beg + k;
                float v = si + rec_c[e * H + h] + rows[(IDX)rec_st[e].x * ld + HC + H + h];
                l[k] = v > 0.f ? v : slope * v;
            }
        }
        const float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                        // same summation order as the loop below
            l[k] = k < deg ? expf(l[k] - mx) : 0.f;
            if (k < deg) sum += l[k];
        }
        sum += 1e-16f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < deg) {
                const int e = beg + k;
                const float a = l[k] / sum;
                alpha[(int64_t)(e0 + e) * H + h] = a;
                rec_c[e * H + h] = USE_EP ? a * rec_val[e] : a;
            }
        return;
    }
    float mx = -INFINITY, sum = 0.f;
    for (int e = beg; e < end; ++e) {
        float l = si + rec_c[e * H + h] + rows[(IDX)rec_st[e].x * ld + HC + H + h];
        l = l > 0.f ? l : slope * l;
        rec_c[e * H + h] = l;
        mx = fmaxf(mx, l);
    }
    for (int e = beg; e < end; ++e) {
        const float x = expf(rec_c[e * H + h] - mx);
        rec_c[e * H + h] = x;
        sum += x;
    }
    sum += 1e-16f;
    for (int e = beg; e < end; ++e) {
        const float a = rec_c[e * H + h] / sum;
        alpha[(int64_t)(e0 + e) * H + h] = a;
        rec_c[e * H + h] = USE_EP ? a * rec_val[e] : a;
    }
}

// out[t0+d, item] = sum over the records of node d of c[e, head(item)] * (e_ij (.) row[slot(e)])[item]
// shared by the forward (rows = xpe window, out = agg) and the backward source pass (rows = g_agg window, out = g_xpe)
template <int H, int CPI, bool USE_EP, typename IDX>
__device__ __forceinline__ void win_aggregate(const float* rows, int ld, int nq, ItemGeom ig, int nd, int64_t t0, int De,
                                              const float* __restrict__ ea, const int32_t* __restrict__ ea_pos, int e0,
                                              const float4* We4, const int* rp, const int2* rec_st, const float* rec_c,
                                              float* __restrict__ out, int64_t ldo) {
    const int total = nd * ig.ni;
    int d = threadIdx.x / ig.ni, g = threadIdx.x - d * ig.ni;
    const int dstep = kWinThreads / ig.ni, gstep = kWinThreads - dstep * ig.ni;
    for (int item = threadIdx.x; item < total; item += kWinThreads) {
        const int beg = rp[d], end = rp[d + 1];
        const int q0 = g * CPI, h = g / ig.tph;
        float4 acc[CPI];
#pragma unroll
        for (int k = 0; k < CPI; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = beg; e < end; ++e) {
            const int2 st = rec_st[e];
            const float c = rec_c[e * H + h];
            const float4* xj = reinterpret_cast<const float4*>(rows + (IDX)st.x * ld) + q0;
            float4 m[CPI];
#pragma unroll
            for (int k = 0; k < CPI; ++k) m[k] = xj[k];
            if (USE_EP) {
                if (st.y >= 0) {
                    const float4* w = We4 + st.y * nq + q0;
#pragma unroll
                    for (int k = 0; k < CPI; ++k) m[k] = mul4(m[k], w[k]);
                } else {
                    const int64_t p = ea_pos ? (int64_t)ea_pos[e0 + e] : (int64_t)(e0 + e);
#pragma unroll
                    for (int k = 0; k < CPI; ++k) m[k] = mul4(m[k], ep_general(ea + p * De, De, We4, nq, q0 + k));
                }
            }
#pragma unroll
            for (int k = 0; k < CPI; ++k) acc[k] = fma4(c, m[k], acc[k]);
        }
        float4* o = reinterpret_cast<float4*>(out + (t0 + d) * ldo) + q0;
#pragma unroll
        for (int k = 0; k < CPI; ++k) o[k] = acc[k];
        g += gstep; d += dstep;
        if (g >= ig.ni) { g -= ig.ni; ++d; }
    }
}

// tile with more edges than the record buffer (hub destinations): per-destination routine from global memory;
// alpha doubles as the scratch for logits and exponentials
template <int H, bool USE_EP>
__device__ __noinline__ void fwd_overflow_tile(const WinArgs& a, const float4* We4, const float* Ae, const int* rp, int nd, int64_t t0,
                                               int e0, float* __restrict__ agg, float* __restrict__ alpha) {
    const int HC = H * a.C, nq = HC >> 2, De = a.De;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = warp; d < nd; d += kWinWarps) {
        const int64_t i = t0 + d;
        const int beg = e0 + rp[d], end = e0 + rp[d + 1];
        float si[H], mx[H], sum[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { si[h] = a.xpe[i * a.ld + HC + h]; mx[h] = -INFINITY; sum[h] = 0.f; }
        for (int p = beg + lane; p < end; p += 32) {
            const int64_t j = a.other[p];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float le = 0.f;
                for (int dd = 0; dd < De; ++dd) le = fmaf(a.ea[(int64_t)p * De + dd], Ae[dd * H + h], le);
                float l = si[h] + le + a.xpe[j * a.ld + HC + H + h];
                l = l > 0.f ? l : a.slope * l;
                mx[h] = fmaxf(mx[h], l);
                alpha[(int64_t)p * H + h] = l;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) mx[h] = warp_max(mx[h]);
        for (int p = beg + lane; p < end; p += 32)
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float x = expf(alpha[(int64_t)p * H + h] - mx[h]);
                alpha[(int64_t)p * H + h] = x;
                sum[h] += x;
            }
#pragma unroll
        for (int h = 0; h < H; ++h) sum[h] = warp_sum(sum[h]) + 1e-16f;
        for (int p = beg + lane; p < end; p += 32)
#pragma unroll
            for (int h = 0; h < H; ++h) alpha[(int64_t)p * H + h] = alpha[(int64_t)p * H + h] / sum[h];
        __syncwarp();
        for (int q = lane; q < nq; q += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const int hh = (4 * q) / a.C;
            for (int p = beg; p < end; ++p) {
                float4 m = ld4(a.xpe + (int64_t)a.other[p] * a.ld + 4 * q);
                if (USE_EP) m = mul4(m, ep_general(a.ea + (int64_t)p * De, De, We4, nq, q));
                acc = fma4(alpha[(int64_t)p * H + hh], m, acc);
            }
            *reinterpret_cast<float4*>(agg + i * HC + 4 * q) = acc;
        }
    }
}

constexpr int kWin2CtasPerSM = 2;
constexpr int kWin2MaxThreads = 256;      // 8 warps x 2 CTAs = 4 warps per SM sub-partition at up to 128 registers
constexpr int kWin2SmemBudget = 108 * 1024;
constexpr int kWin2PrefDe = 4;               // edge_attr words prefetched per edge (wider rows are read at the top of the tile)

template <int H>
__device__ __forceinline__ int item_head(int g, int tph) {
    int h = 0;
#pragma unroll
    for (int q = 1; q < H; ++q) h += (g >= q * tph) ? 1 : 0;
    return h;
}

template <int H, bool USE_EP, typename IDX>
__device__ __forceinline__ void fwd_softmax2(const float* rows, int ld, IDX own0, int HC, int nd, int e0, float slope, const int* rp,
                                             const int2* rec_st, const float* rec_val, float* rec_c, float* __restrict__ alpha) {
    for (int idx = threadIdx.x; idx < nd * H; idx += blockDim.x) {
        const int d = idx / H, h = idx - d * H;
        const int beg = rp[d], end = rp[d + 1], deg = end - beg;
        const float si = rows[(own0 + d) * ld + HC + h];
        if (deg <= 4) {
            float l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                l[k] = -INFINITY;
                if (k < deg) {
                    const int e = beg + k;
                    const float v = si + rec_c[e * H + h] + rows[(IDX)rec_st[e].x * ld + HC + H + h];
                    l[k] = v > 0.f ? v : slope * v;
                }
            }
            const float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                l[k] = k < deg ? expf(l[k] - mx) : 0.f;
                if (k < deg) sum += l[k];
            }
            sum += 1e-16f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < deg) {
                    const int e = beg + k;
                    const float a = l[k] / sum;
                    alpha[(int64_t)(e0 + e) * H + h] = a;
                    rec_c[e * H + h] = USE_EP ? a * rec_val[e] : a;
                }
            continue;
        }
        float mx = -INFINITY, sum = 0.f;
        for (int e = beg; e < end; ++e) {
            float l = si + rec_c[e * H + h] + rows[(IDX)rec_st[e].x * ld + HC + H + h];
            l = l > 0.f ? l : slope * l;
            rec_c[e * H + h] = l;
            mx = fmaxf(mx, l);
        }
        for (int e = beg; e < end; ++e) {
            const float x = expf(rec_c[e * H + h] - mx);
            rec_c[e * H + h] = x;
            sum += x;
        }
        sum += 1e-16f;
        for (int e = beg; e < end; ++e) {
            const float a = rec_c[e * H + h] / sum;
            alpha[(int64_t)(e0 + e) * H + h] = a;
            rec_c[e * H + h] = USE_EP ? a * rec_val[e] : a;
        }
    }
}

template <int H, int CPI, bool USE_EP, typename IDX>
__device__ __forceinline__ void win_aggregate2(const float* rows, int ld, int nq, ItemGeom ig, int nd, int64_t t0, int De,
                                               const float* __restrict__ ea, const int32_t* __restrict__ ea_pos, int e0,
                                               const float4* We4, const int* rp, const int2* rec_st, const float* rec_c,
                                               float* __restrict__ out, int64_t ldo, int d, int g, int dstep, int gstep) {
    const int total = nd * ig.ni;
    for (int item = threadIdx.x; item < total; item += blockDim.x) {
        const int beg = rp[d], end = rp[d + 1];
        const int q0 = g * CPI, h = item_head<H>(g, ig.tph);
        float4 acc[CPI];
#pragma unroll
        for (int k = 0; k < CPI; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = beg; e < end; ++e) {
            const int2 st = rec_st[e];
            const float c = rec_c[e * H + h];
            const float4* xj = reinterpret_cast<const float4*>(rows + (IDX)st.x * ld) + q0;
            float4 m[CPI];
#pragma unroll
            for (int k = 0; k < CPI; ++k) m[k] = xj[k];
            if (USE_EP) {
                if (st.y >= 0) {
                    const float4* w = We4 + st.y * nq + q0;
#pragma unroll
                    for (int k = 0; k < CPI; ++k) m[k] = mul4(m[k], w[k]);
                } else {
                    const int64_t p = ea_pos ? (int64_t)ea_pos[e0 + e] : (int64_t)(e0 + e);
#pragma unroll
                    for (int k = 0; k < CPI; ++k) m[k] = mul4(m[k], ep_general(ea + p * De, De, We4, nq, q0 + k));
                }
            }
#pragma unroll
            for (int k = 0; k < CPI; ++k) acc[k] = fma4(c, m[k], acc[k]);
        }
        float4* o = reinterpret_cast<float4*>(out + (t0 + d) * ldo) + q0;
#pragma unroll
        for (int k = 0; k < CPI; ++k) o[k] = acc[k];
        g += gstep; d += dstep;
        if (g >= ig.ni) { g -= ig.ni; ++d; }
    }
}

__device__ __forceinline__ void win2_issue_window(const WinArgs& a, const int4& dsc, float* dst, uint64_t* bar, int warp, int lane) {
    const int nrows = dsc.y - dsc.x, ne = dsc.w - dsc.z;
    const bool near = ne <= kWinMaxEdges && nrows <= a.rmax && ne > 0;
    if (near && warp == 0) {
        tc::fence_proxy_async_smem();
        bulk_window(dst, a.xpe + (int64_t)dsc.x * a.ld, (uint32_t)nrows * (uint32_t)a.ld * 4u, bar, lane);
    }
}
// index / edge_attr words of a tile into registers (nothing here is waited for: the values are first used at the top of
// the tile they belong to).  Plain scalars: arrays here ended up in local memory, which turns every prefetch into a
// blocking load.
struct WinPref { int j0, j1, rp; float a0, a1, a2, a3, b0, b1, b2, b3; };
__device__ __forceinline__ void win2_prefetch(const WinArgs& a, const int4& dsc, int64_t tile, int tid, int nthr, bool pref_ok,
                                              WinPref& pf) {
    const int e0 = dsc.z, ne = dsc.w - dsc.z, De = a.De;
    const int64_t t0 = tile * a.D;
    const int nd = (int)min((int64_t)a.D, a.N - t0);
    if (tid <= nd) pf.rp = a.rowptr[t0 + tid];
    if (ne <= kWinMaxEdges && pref_ok) {
        if (tid < ne) {
            const int64_t p = e0 + tid;
            const float* r = a.ea + p * De;
            pf.j0 = a.other[p];
            pf.a0 = r[0];
            if (De > 1) pf.a1 = r[1];
            if (De > 2) pf.a2 = r[2];
            if (De > 3) pf.a3 = r[3];
        }
        if (tid + nthr < ne) {
            const int64_t p = e0 + tid + nthr;
            const float* r = a.ea + p * De;
            pf.j1 = a.other[p];
            pf.b0 = r[0];
            if (De > 1) pf.b1 = r[1];
            if (De > 2) pf.b2 = r[2];
            if (De > 3) pf.b3 = r[3];
        }
    }
}

// one prefetched edge -> record
template <int H>
__device__ __forceinline__ void win2_make_record(int e, int j, float v0, float v1, float v2, float v3, int De, int base, const float* Ae,
                                                 int2* rec_st, float* rec_val, float* rec_c) {
    float l[H];
#pragma unroll
    for (int h = 0; h < H; ++h) l[h] = 0.f;
    int nz = 0, ty = 0;
    float val = 0.f;
    const float v[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int dd = 0; dd < 4; ++dd)
        if (dd < De) {
            if (v[dd] != 0.f) { ++nz; ty = dd; val = v[dd]; }
#pragma unroll
            for (int h = 0; h < H; ++h) l[h] = fmaf(v[dd], Ae[dd * H + h], l[h]);
        }
    rec_st[e] = make_int2(j - base, nz == 1 ? ty : -1);
    rec_val[e] = nz == 1 ? val : 1.f;
#pragma unroll
    for (int h = 0; h < H; ++h) rec_c[e * H + h] = l[h];
}

template <int H, int CPI, bool USE_EP>
__global__ void __launch_bounds__(kWin2MaxThreads, kWin2CtasPerSM)
edge_win2_fwd_kernel(const __grid_constant__ WinArgs a, float* __restrict__ agg, float* __restrict__ alpha) {
    extern __shared__ __align__(128) float smem_f[];
    const int HC = H * a.C, nq = HC >> 2, De = a.De, ld = (int)a.ld;
    const ItemGeom ig{nq / CPI, (a.C >> 2) / CPI};
    WinSmem s = win_carve(smem_f, 2 * a.rmax, a.ld, De, nq, H, USE_EP);
    const int bufstride = a.rmax * ld;                            // floats between the two window buffers
    uint64_t* bar = s.bar;                                        // two mbarriers (16 bytes reserved)
    int2* rec_st = reinterpret_cast<int2*>(s.rec);
    float* rec_val = s.rec + 2 * kWinMaxEdges;
    float* rec_c = rec_val + kWinMaxEdges;                        // [kWinMaxEdges][H]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    if (USE_EP)
        for (int i = tid; i < De * nq; i += nthr) s.We4[i] = ld4(a.w_edge + 4 * i);
    for (int i = tid; i < De * H; i += nthr) s.Ae[i] = a.att_edge[i];
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::fence_mbar_init(); }
    const int d_first = tid / ig.ni, g_first = tid - d_first * ig.ni, dstep = nthr / ig.ni, gstep = nthr - dstep * ig.ni;
    const bool pref_ok = De <= kWin2PrefDe;
    WinPref pf{0, 0, 0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};    // two edges per thread + one rowptr word
    __syncthreads();
    int4 desc = blockIdx.x < a.T ? a.tiles[blockIdx.x] : make_int4(0, 0, 0, 0);
    int4 desc_n = make_int4(0, 0, 0, 0);
    if (blockIdx.x < a.T) {
        win2_issue_window(a, desc, s.rows, bar, warp, lane);
        win2_prefetch(a, desc, blockIdx.x, tid, nthr, pref_ok, pf);
        if ((int64_t)blockIdx.x + gridDim.x < a.T) desc_n = a.tiles[blockIdx.x + gridDim.x];
    }
    uint32_t par = 0;
    int k = 0;
    for (int64_t tile = blockIdx.x; tile < a.T; tile += gridDim.x, ++k) {
        const int cur = k & 1;
        const int64_t t0 = tile * a.D;
        const int nd = (int)min((int64_t)a.D, a.N - t0);
        const int lo = desc.x, nrows = desc.y - desc.x, e0 = desc.z, ne = desc.w - desc.z;
        const bool overflow = ne > kWinMaxEdges;
        const bool near = !overflow && nrows <= a.rmax && ne > 0;
        const int base = near ? lo : 0;
        // ---- records of this tile from the prefetched registers
        if (tid <= nd) s.rp[tid] = pf.rp - e0;
        if (!overflow) {
            if (pref_ok) {
                if (tid < ne) win2_make_record<H>(tid, pf.j0, pf.a0, pf.a1, pf.a2, pf.a3, De, base, s.Ae, rec_st, rec_val, rec_c);
                if (tid + nthr < ne)
                    win2_make_record<H>(tid + nthr, pf.j1, pf.b0, pf.b1, pf.b2, pf.b3, De, base, s.Ae, rec_st, rec_val, rec_c);
            } else {
                for (int e = tid; e < ne; e += nthr) {
                    const int p = e0 + e;
                    const float* earow = a.ea + (int64_t)p * De;
                    float l[H];
#pragma unroll
                    for (int h = 0; h < H; ++h) l[h] = 0.f;
                    int nz = 0, ty = 0;
                    float val = 0.f;
                    for (int dd = 0; dd < De; ++dd) {
                        const float v = earow[dd];
                        if (v != 0.f) { ++nz; ty = dd; val = v; }
#pragma unroll
                        for (int h = 0; h < H; ++h) l[h] = fmaf(v, s.Ae[dd * H + h], l[h]);
                    }
                    rec_st[e] = make_int2(a.other[p] - base, nz == 1 ? ty : -1);
                    rec_val[e] = nz == 1 ? val : 1.f;
#pragma unroll
                    for (int h = 0; h < H; ++h) rec_c[e * H + h] = l[h];
                }
            }
        }
        __syncthreads();                                            // records visible; every thread is done with tile k-1
        // ---- tile k+1: window copy into the other buffer, index words into registers, descriptor of k+2
        const int64_t nxt = tile + gridDim.x;
        int4 desc_nn = make_int4(0, 0, 0, 0);
        if (nxt < a.T) {
            win2_issue_window(a, desc_n, s.rows + (cur ^ 1) * bufstride, bar + (cur ^ 1), warp, lane);
            win2_prefetch(a, desc_n, nxt, tid, nthr, pref_ok, pf);
            if (nxt + gridDim.x < a.T) desc_nn = a.tiles[nxt + gridDim.x];
        }
        if (overflow) {
            fwd_overflow_tile<H, USE_EP>(a, s.We4, s.Ae, s.rp, nd, t0, e0, agg, alpha);
        } else if (near) {
            tc::mbar_wait(bar + cur, (par >> cur) & 1u);
            par ^= 1u << cur;
            const float* rows = s.rows + cur * bufstride;
            fwd_softmax2<H, USE_EP, int>(rows, ld, (int)(t0 - lo), HC, nd, e0, a.slope, s.rp, rec_st, rec_val, rec_c, alpha);
            __syncthreads();
            win_aggregate2<H, CPI, USE_EP, int>(rows, ld, nq, ig, nd, t0, De, a.ea, nullptr, e0, s.We4, s.rp, rec_st, rec_c, agg, HC,
                                                d_first, g_first, dstep, gstep);
        } else {
            fwd_softmax2<H, USE_EP, int64_t>(a.xpe, ld, t0, HC, nd, e0, a.slope, s.rp, rec_st, rec_val, rec_c, alpha);
            __syncthreads();
            win_aggregate2<H, CPI, USE_EP, int64_t>(a.xpe, ld, nq, ig, nd, t0, De, a.ea, nullptr, e0, s.We4, s.rp, rec_st, rec_c, agg, HC,
                                                    d_first, g_first, dstep, gstep);
        }
        __syncthreads();                                            // records are dead
        desc = desc_n;
        desc_n = desc_nn;
    }
}

// ------------------------------------------------------------------------------------------------ backward, destination pass
// Two CTAs per SM (the g_weight_edge accumulators live in registers).  Staged per tile: the xpe window AND the tile's own
// g_agg rows.  records: {src slot, local dst, type, value} | alpha[H] | g[H] (g_alpha, then consumed by the softmax phase)
// Dots are EDGE-parallel: a warp takes 32/ni edges at a time, lane = (edge slot, item); the per-head sum over an edge's
// items is ni/H - 1 shuffles; no degree loops, no divergence.
constexpr int kWinDstCtasPerSM = 2;
constexpr int kWinDstSmemBudget = 112 * 1024;

// math of one (edge, item) pair once its operands are in registers: returns this item's part of g_alpha and adds the
// edge's contribution to the g_weight_edge accumulators
template <int H, int CPI, bool USE_EP, int DR>
__device__ __forceinline__ float dots_edge(const int4 r, const float4 (&gm)[CPI], float ah, int q0, int nq, int De, int64_t p,
                                           const float* __restrict__ ea, const float4* We4, float4 (&gw)[DR][CPI]) {
    float v = 0.f;
    if (USE_EP) {
        const float val = __int_as_float(r.w);
        if (r.z >= 0) {
            const float4* w = We4 + r.z * nq + q0;
#pragma unroll
            for (int k = 0; k < CPI; ++k) v += dot4(gm[k], w[k]);
            v *= val;
            const float c = val * ah;
            if (r.z == 0) {
#pragma unroll
                for (int k = 0; k < CPI; ++k) gw[0][k] = fma4(c, gm[k], gw[0][k]);
            } else if (r.z == 1) {
#pragma unroll
                for (int k = 0; k < CPI; ++k) gw[1][k] = fma4(c, gm[k], gw[1][k]);
            } else if (r.z == 2) {
#pragma unroll
                for (int k = 0; k < CPI; ++k) gw[2][k] = fma4(c, gm[k], gw[2][k]);
            } else if (DR > 3) {
#pragma unroll
                for (int k = 0; k < CPI; ++k) gw[DR - 1][k] = fma4(c, gm[k], gw[DR - 1][k]);
            }
        } else {
            const float* earow = ea + p * De;
#pragma unroll
            for (int dd = 0; dd < DR; ++dd)
                if (dd < De) {
                    const float ed = earow[dd];
#pragma unroll
                    for (int k = 0; k < CPI; ++k) {
                        v = fmaf(ed, dot4(gm[k], We4[dd * nq + q0 + k]), v);
                        gw[dd][k] = fma4(ed * ah, gm[k], gw[dd][k]);
                    }
                }
        }
    } else {
#pragma unroll
        for (int k = 0; k < CPI; ++k) v += (gm[k].x + gm[k].y) + (gm[k].z + gm[k].w);
    }
    return v;
}

// edge-parallel dots: a warp takes 32/ni edges per pass, lane = (edge slot, item)
template <int H, int CPI, bool USE_EP, int DR, typename IDX>
__device__ __forceinline__ void bwd_dots(const float* rows, int ld, const float4* gagg4, int nq, ItemGeom ig, int ne, int e0, int De,
                                         const float* __restrict__ ea, const float4* We4, const int4* rec4, const float* rec_a,
                                         float* rec_g, float4 (&gw)[DR][CPI]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int epw = 32 / ig.ni;                                   // edges per warp pass
    const int el = lane / ig.ni, g = lane - el * ig.ni;
    const int q0 = g * CPI, h = item_head<H>(min(g, ig.ni - 1), ig.tph);
    const bool lane_on = el < epw;
    const bool writer = lane_on && (g - (g / ig.tph) * ig.tph) == 0;
    for (int eb = warp * epw; eb < ne; eb += kWinWarps * epw) {
        const int e = eb + el;
        const bool on = lane_on && e < ne;
        float v = 0.f;
        if (on) {
            const int4 r = rec4[e];
            const float4* xj = reinterpret_cast<const float4*>(rows + (IDX)r.x * ld) + q0;
            const float4* gi = gagg4 + r.y * nq + q0;
            float4 gm[CPI];
#pragma unroll
            for (int k = 0; k < CPI; ++k) gm[k] = mul4(gi[k], xj[k]);
            const float ah = USE_EP ? rec_a[e * H + h] : 0.f;
            v = dots_edge<H, CPI, USE_EP, DR>(r, gm, ah, q0, nq, De, (int64_t)(e0 + e), ea, We4, gw);
        }
        // per-head sum over the items of this edge: lanes g .. g + tph - 1 (fixed order)
        float sdot = v;
        for (int k = 1; k < ig.tph; ++k) sdot += __shfl_down_sync(0xffffffffu, v, k);
        if (on && writer) rec_g[e * H + h] = sdot;
    }
}

template <int H, typename IDX>
__device__ __forceinline__ void bwd_softmax(const float* rows, int ld, IDX own0, int HC, int nd, int64_t t0, int e0, int De,
                                            float slope, const float* __restrict__ ea, const float* Ae, const int* rp,
                                            const int4* rec4, const float* rec_a, const float* rec_g, float* __restrict__ g_logit,
                                            float* __restrict__ g_xpe, int64_t ldg) {
    const int idx = threadIdx.x;
    if (idx >= nd * H) return;
    const int d = idx / H, h = idx - d * H;
    const int beg = rp[d], end = rp[d + 1];
    const float si = rows[(own0 + d) * ld + HC + h];
    float dot = 0.f, gsi = 0.f;
    for (int e = beg; e < end; ++e) dot = fmaf(rec_a[e * H + h], rec_g[e * H + h], dot);
    for (int e = beg; e < end; ++e) {
        // same association as the forward: (s_i + edge part) + s_j, the edge part summed from zero
        const int4 r = rec4[e];
        float le = 0.f;
        if (r.z >= 0) le = __fmul_rn(__int_as_float(r.w), Ae[r.z * H + h]);
        else for (int dd = 0; dd < De; ++dd) le = fmaf(ea[(int64_t)(e0 + e) * De + dd], Ae[dd * H + h], le);
        const float l = __fadd_rn(__fadd_rn(si, le), rows[(IDX)r.x * ld + HC + H + h]);
        float g = rec_a[e * H + h] * (rec_g[e * H + h] - dot);
        g *= (l > 0.f ? 1.f : slope);
        g_logit[(int64_t)(e0 + e) * H + h] = g;
        gsi += g;
    }
    g_xpe[(t0 + d) * ldg + HC + h] = gsi;
}

template <int H, bool USE_EP>
__device__ __noinline__ void bwd_dst_overflow_tile(const WinArgs& a, const float4* We4, const float* Ae, const int* rp, int nd, int64_t t0,
                                                   int e0, const float* __restrict__ alpha, const float* __restrict__ g_agg,
                                                   float* __restrict__ g_logit, float* __restrict__ g_xpe, float4* slab_warp) {
    // slab_warp: this warp's [De][nq] float4 scratch (zeroed by the caller); receives the tile's g_weight_edge contribution
    const int HC = H * a.C, nq = HC >> 2, De = a.De;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = warp; d < nd; d += kWinWarps) {
        const int64_t i = t0 + d;
        const int beg = e0 + rp[d], end = e0 + rp[d + 1];
        float dot[H];
#pragma unroll
        for (int h = 0; h < H; ++h) dot[h] = 0.f;
        for (int p = beg; p < end; ++p) {
            const float* xj = a.xpe + (int64_t)a.other[p] * a.ld;
            const float* earow = a.ea + (int64_t)p * De;
            float part[H];
#pragma unroll
            for (int h = 0; h < H; ++h) part[h] = 0.f;
            for (int q = lane; q < nq; q += 32) {
                const float4 gm = mul4(ld4(g_agg + i * HC + 4 * q), ld4(xj + 4 * q));
                const int hh = (4 * q) / a.C;
                float v;
                if (USE_EP) {
                    const float ah = alpha[(int64_t)p * H + hh];
                    float4 ep = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int dd = 0; dd < De; ++dd) {
                        const float ed = earow[dd];
                        ep = fma4(ed, We4[dd * nq + q], ep);
                        slab_warp[dd * nq + q] = fma4(ed * ah, gm, slab_warp[dd * nq + q]);
                    }
                    v = dot4(gm, ep);
                } else {
                    v = (gm.x + gm.y) + (gm.z + gm.w);
                }
#pragma unroll
                for (int h = 0; h < H; ++h) part[h] += (hh == h) ? v : 0.f;
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                part[h] = warp_sum(part[h]);
                dot[h] = fmaf(alpha[(int64_t)p * H + h], part[h], dot[h]);
            }
            if (lane < H) {
                float v = part[0];
#pragma unroll
                for (int h = 1; h < H; ++h) v = lane == h ? part[h] : v;
                g_logit[(int64_t)p * H + lane] = v;
            }
        }
        __syncwarp();
        float si[H], gsi[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { si[h] = a.xpe[i * a.ld + HC + h]; gsi[h] = 0.f; }
        for (int p = beg + lane; p < end; p += 32) {
            const int64_t j = a.other[p];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float le = 0.f;
                for (int dd = 0; dd < De; ++dd) le = fmaf(a.ea[(int64_t)p * De + dd], Ae[dd * H + h], le);
                const float l = __fadd_rn(__fadd_rn(si[h], le), a.xpe[j * a.ld + HC + H + h]);
                float g = alpha[(int64_t)p * H + h] * (g_logit[(int64_t)p * H + h] - dot[h]);
                g *= (l > 0.f ? 1.f : a.slope);
                g_logit[(int64_t)p * H + h] = g;
                gsi[h] += g;
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) gsi[h] = warp_sum(gsi[h]);
        if (lane < H) {
            float v = gsi[0];
#pragma unroll
            for (int h = 1; h < H; ++h) v = lane == h ? gsi[h] : v;
            g_xpe[i * a.ld + HC + lane] = v;
        }
        __syncwarp();
    }
}

// smem: rows (xpe window) | We4 | Ae | rp | bar | rec4 | rec_a | rec_g | gagg [D][HC] | ovf slabs [warps][De][nq] float4
constexpr int kDstMaxEdges = 256;            // records per tile (molecular tiles: <= 4 in-edges per atom)

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct Dst2Smem {
    float* rows; float4* We4; float* Ae; uint64_t* bar; int* rp;
    int4* rec4; float* rec_a; float* rec_g; float* rec_sj; float* rec_si;
    int* raw_src; float* raw_ea; float* raw_al; int* raw_rp;
    float* gagg; float4* ovf;
};
__host__ __device__ inline size_t dst2_fixed_floats(int De, int nq, int H, bool use_ep) {
    const int HC = nq * 4;
    return (size_t)(use_ep ? De * nq * 4 : 0) + 32 + 4 + (kWinMaxTile + 4) + (size_t)kDstMaxEdges * (4 + 3 * H) + kWinMaxTile * 4 +
           (size_t)kDstMaxEdges * (1 + kWinRegDe + H) + (kWinMaxTile + 4) + (size_t)kWinMaxTile * HC +
           (use_ep ? (size_t)kWinWarps * De * nq * 4 : 0);
}
__device__ __forceinline__ Dst2Smem dst2_carve(float* base, int rmax, int ld, int De, int nq, int H, bool use_ep) {
    Dst2Smem s;
    float* p = base;
    s.rows = p; p += (size_t)rmax * ld;
    s.We4 = reinterpret_cast<float4*>(p); p += use_ep ? De * nq * 4 : 0;
    s.Ae = p; p += 32;
    s.bar = reinterpret_cast<uint64_t*>(p); p += 4;
    s.rp = reinterpret_cast<int*>(p); p += kWinMaxTile + 4;
    s.rec4 = reinterpret_cast<int4*>(p); p += kDstMaxEdges * 4;
    s.rec_a = p; p += kDstMaxEdges * H;
    s.rec_g = p; p += kDstMaxEdges * H;
    s.rec_sj = p; p += kDstMaxEdges * H;
    s.rec_si = p; p += kWinMaxTile * 4;
    s.raw_src = reinterpret_cast<int*>(p); p += kDstMaxEdges;
    s.raw_ea = p; p += kDstMaxEdges * kWinRegDe;
    s.raw_al = p; p += kDstMaxEdges * H;
    s.raw_rp = reinterpret_cast<int*>(p); p += kWinMaxTile + 4;
    s.gagg = p; p += (size_t)kWinMaxTile * nq * 4;
    s.ovf = reinterpret_cast<float4*>(p);
    return s;
}

template <int H, int CPI, bool USE_EP, int DR>
__global__ void __launch_bounds__(kWinThreads, kWinDstCtasPerSM)
edge_win_bwd_dst2_kernel(const __grid_constant__ WinArgs a, const float* __restrict__ alpha, const float* __restrict__ g_agg,
                         float* __restrict__ g_logit, float* __restrict__ g_xpe, float* __restrict__ gwe_partial) {
    extern __shared__ __align__(128) float smem_f[];
    const int HC = H * a.C, nq = HC >> 2, De = a.De, ld = (int)a.ld;
    const ItemGeom ig{nq / CPI, (a.C >> 2) / CPI};
    const Dst2Smem s = dst2_carve(smem_f, a.rmax, ld, De, nq, H, USE_EP);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (USE_EP)
        for (int i = tid; i < De * nq; i += kWinThreads) s.We4[i] = ld4(a.w_edge + 4 * i);
    for (int i = tid; i < De * H; i += kWinThreads) s.Ae[i] = a.att_edge[i];
    if (USE_EP)
        for (int i = tid; i < kWinWarps * De * nq; i += kWinThreads) s.ovf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) { tc::mbar_init(s.bar, 1); tc::fence_mbar_init(); }
    float4 gw[DR][CPI];
#pragma unroll
    for (int d = 0; d < DR; ++d)
#pragma unroll
        for (int k = 0; k < CPI; ++k) gw[d][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t parity = 0;

    // raw words of a tile -> shared memory, asynchronously (every thread commits one group, possibly empty)
    auto issue_raw = [&](const int4& dsc, int64_t tile) {
        const int e0 = dsc.z, ne = dsc.w - dsc.z;
        const int64_t t0 = tile * a.D;
        const int nd = (int)min((int64_t)a.D, a.N - t0);
        if (tid <= nd) cp_async4(s.raw_rp + tid, a.rowptr + t0 + tid);
        if (ne <= kDstMaxEdges) {
            for (int i = tid; i < ne; i += kWinThreads) cp_async4(s.raw_src + i, a.other + e0 + i);
            for (int i = tid; i < ne * De; i += kWinThreads) cp_async4(s.raw_ea + i, a.ea + (int64_t)e0 * De + i);
            for (int i = tid; i < ne * H; i += kWinThreads) cp_async4(s.raw_al + i, alpha + (int64_t)e0 * H + i);
        }
        cp_async_commit();
    };
    // window + the tile's own g_agg rows by the copy engine (warp 0)
    auto issue_bulk = [&](const int4& dsc, int64_t tile) {
        const int nrows = dsc.y - dsc.x, ne = dsc.w - dsc.z;
        const int64_t t0 = tile * a.D;
        const int nd = (int)min((int64_t)a.D, a.N - t0);
        const bool staged = ne <= kDstMaxEdges && ne > 0;
        if (staged && warp == 0) {
            const bool near = nrows <= a.rmax;
            tc::fence_proxy_async_smem();
            const uint32_t wbytes = near ? (uint32_t)nrows * (uint32_t)ld * 4u : 0u, gbytes = (uint32_t)nd * (uint32_t)HC * 4u;
            if (lane == 0) tc::mbar_expect_tx(s.bar, wbytes + gbytes);
            __syncwarp();
            for (uint32_t off = (uint32_t)lane * kBulkChunk; off < wbytes; off += 32u * kBulkChunk)
                bulk_g2s(reinterpret_cast<char*>(s.rows) + off, reinterpret_cast<const char*>(a.xpe + (int64_t)dsc.x * a.ld) + off,
                         min(kBulkChunk, wbytes - off), s.bar);
            for (uint32_t off = (uint32_t)lane * kBulkChunk; off < gbytes; off += 32u * kBulkChunk)
                bulk_g2s(reinterpret_cast<char*>(s.gagg) + off, reinterpret_cast<const char*>(g_agg + t0 * HC) + off,
                         min(kBulkChunk, gbytes - off), s.bar);
        }
    };

    __syncthreads();
    int4 desc = blockIdx.x < a.T ? a.tiles[blockIdx.x] : make_int4(0, 0, 0, 0);
    int4 desc_n = make_int4(0, 0, 0, 0);
    if (blockIdx.x < a.T) {
        issue_bulk(desc, blockIdx.x);
        issue_raw(desc, blockIdx.x);
        if ((int64_t)blockIdx.x + gridDim.x < a.T) desc_n = a.tiles[blockIdx.x + gridDim.x];
    }
    const float4* gagg4 = reinterpret_cast<const float4*>(s.gagg);
    for (int64_t tile = blockIdx.x; tile < a.T; tile += gridDim.x) {
        const int64_t t0 = tile * a.D;
        const int nd = (int)min((int64_t)a.D, a.N - t0);
        const int lo = desc.x, nrows = desc.y - desc.x, e0 = desc.z, ne = desc.w - desc.z;
        const int64_t nxt = tile + gridDim.x;
        int4 desc_nn = make_int4(0, 0, 0, 0);
        if (nxt + gridDim.x < a.T) desc_nn = a.tiles[nxt + gridDim.x];
        const bool overflow = ne > kDstMaxEdges;
        const bool staged = !overflow && ne > 0;
        const bool near = staged && nrows <= a.rmax;
        cp_async_wait_all();
        __syncthreads();                                            // raw words of this tile are visible; previous tile is finished
        if (tid <= nd) s.rp[tid] = s.raw_rp[tid] - e0;
        if (overflow) {
            __syncthreads();
            if (nxt < a.T) issue_raw(desc_n, nxt);
            bwd_dst_overflow_tile<H, USE_EP>(a, s.We4, s.Ae, s.rp, nd, t0, e0, alpha, g_agg, g_logit, g_xpe, s.ovf + warp * De * nq);
            __syncthreads();
            if (nxt < a.T) issue_bulk(desc_n, nxt);
            desc = desc_n;
            desc_n = desc_nn;
            continue;
        }
        // ---- records from the staged raw words
        const int base = near ? lo : 0;
        int* rec_w = reinterpret_cast<int*>(s.rec4);
        for (int e = tid; e < ne; e += kWinThreads) {
            const float* earow = s.raw_ea + e * De;
            int nz = 0, ty = 0;
            float val = 0.f;
            for (int dd = 0; dd < De; ++dd) {
                const float v = earow[dd];
                if (v != 0.f) { ++nz; ty = dd; val = v; }
            }
            rec_w[4 * e + 0] = s.raw_src[e] - base;
            rec_w[4 * e + 2] = nz == 1 ? ty : -1;
            rec_w[4 * e + 3] = __float_as_int(nz == 1 ? val : 1.f);
        }
        for (int d = tid; d < nd; d += kWinThreads)
            for (int e = s.raw_rp[d] - e0; e < s.raw_rp[d + 1] - e0; ++e) rec_w[4 * e + 1] = d;
        for (int i = tid; i < ne * H; i += kWinThreads) s.rec_a[i] = s.raw_al[i];
        __syncthreads();                                            // records visible; the raw buffers are free
        if (nxt < a.T) issue_raw(desc_n, nxt);
        if (staged) {
            tc::mbar_wait(s.bar, parity);
            parity ^= 1u;
        }
        // ---- stash s_i / s_j for the softmax phase, then the dots
        if (near) {
            const int own0 = (int)(t0 - lo);
            for (int i = tid; i < ne * H; i += kWinThreads) {
                const int e = i / H, h = i - e * H;
                s.rec_sj[i] = s.rows[s.rec4[e].x * ld + HC + H + h];
            }
            for (int i = tid; i < nd * H; i += kWinThreads) {
                const int d = i / H, h = i - d * H;
                s.rec_si[i] = s.rows[(own0 + d) * ld + HC + h];
            }
            bwd_dots<H, CPI, USE_EP, DR, int>(s.rows, ld, gagg4, nq, ig, ne, e0, De, a.ea, s.We4, s.rec4, s.rec_a, s.rec_g, gw);
        } else {
            for (int i = tid; i < ne * H; i += kWinThreads) {
                const int e = i / H, h = i - e * H;
                s.rec_sj[i] = a.xpe[(int64_t)s.rec4[e].x * ld + HC + H + h];
            }
            for (int i = tid; i < nd * H; i += kWinThreads) {
                const int d = i / H, h = i - d * H;
                s.rec_si[i] = a.xpe[(t0 + d) * ld + HC + h];
            }
            if (ne > 0)
                bwd_dots<H, CPI, USE_EP, DR, int64_t>(a.xpe, ld, gagg4, nq, ig, ne, e0, De, a.ea, s.We4, s.rec4, s.rec_a, s.rec_g, gw);
        }
        __syncthreads();                                            // dots done: window and g_agg buffers are free
        if (nxt < a.T) issue_bulk(desc_n, nxt);
        // ---- softmax + leaky_relu backward from the records alone (overlaps the next tile's copies)
        for (int idx = tid; idx < nd * H; idx += kWinThreads) {
            const int d = idx / H, h = idx - d * H;
            const int beg = s.rp[d], end = s.rp[d + 1];
            const float si = s.rec_si[idx];
            float dot = 0.f, gsi = 0.f;
            for (int e = beg; e < end; ++e) dot = fmaf(s.rec_a[e * H + h], s.rec_g[e * H + h], dot);
            for (int e = beg; e < end; ++e) {
                // same association as the forward: (s_i + edge part) + s_j, the edge part summed from zero
                const int4 r = s.rec4[e];
                float le = 0.f;
                if (r.z >= 0) le = __fmul_rn(__int_as_float(r.w), s.Ae[r.z * H + h]);
                else for (int dd = 0; dd < De; ++dd) le = fmaf(a.ea[(int64_t)(e0 + e) * De + dd], s.Ae[dd * H + h], le);
                const float l = __fadd_rn(__fadd_rn(si, le), s.rec_sj[e * H + h]);
                float g = s.rec_a[e * H + h] * (s.rec_g[e * H + h] - dot);
                g *= (l > 0.f ? 1.f : a.slope);
                g_logit[(int64_t)(e0 + e) * H + h] = g;
                gsi += g;
            }
            g_xpe[(t0 + d) * a.ld + HC + h] = gsi;
        }
        desc = desc_n;
        desc_n = desc_nn;
    }
    if (USE_EP) {
        // fixed-order reduction: lanes (edge slot, item) of every warp -> staging in the (dead) window area -> one partial per CTA
        __syncthreads();
        float4* stage = reinterpret_cast<float4*>(s.rows);        // [kWinWarps][epw][De][nq]
        const int epw = 32 / ig.ni, el = lane / ig.ni, g = lane - el * ig.ni;
        if (el < epw)
#pragma unroll
            for (int d = 0; d < DR; ++d)
                if (d < De)
#pragma unroll
                    for (int k = 0; k < CPI; ++k) stage[((warp * epw + el) * De + d) * nq + g * CPI + k] = gw[d][k];
        __syncthreads();
        float4* P = reinterpret_cast<float4*>(gwe_partial) + (int64_t)blockIdx.x * De * nq;
        for (int idx = tid; idx < De * nq; idx += kWinThreads) {
            float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int w = 0; w < kWinWarps * epw; ++w) {
                const float4 v = stage[w * De * nq + idx];
                sacc.x += v.x; sacc.y += v.y; sacc.z += v.z; sacc.w += v.w;
            }
            for (int w = 0; w < kWinWarps; ++w) {                 // hub-tile contributions
                const float4 v = s.ovf[w * De * nq + idx];
                sacc.x += v.x; sacc.y += v.y; sacc.z += v.z; sacc.w += v.w;
            }
            P[idx] = sacc;
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward, source pass
// tile of consecutive SOURCES; window = rows g_agg[lo:hi) of the destinations of their out-edges
// records: {dst slot, type} | c[H] = alpha * value | gl[H] = g_logit
template <int H, bool USE_EP>
__device__ __noinline__ void src_overflow_tile(const WinArgs& a, const float4* We4, const int* rp, int nd, int64_t t0, int k0,
                                               const int32_t* __restrict__ src_pos, const float* __restrict__ alpha,
                                               const float* __restrict__ g_agg, const float* __restrict__ g_logit,
                                               float* __restrict__ g_xpe) {
    const int HC = H * a.C, nq = HC >> 2, De = a.De;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = warp; d < nd; d += kWinWarps) {
        const int beg = k0 + rp[d], end = k0 + rp[d + 1];
        float* out = g_xpe + (t0 + d) * a.ld;
        for (int q = lane; q < nq; q += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const int hh = (4 * q) / a.C;
            for (int k = beg; k < end; ++k) {
                const int p = src_pos[k];
                float4 m = ld4(g_agg + (int64_t)a.other[k] * HC + 4 * q);
                if (USE_EP) m = mul4(m, ep_general(a.ea + (int64_t)p * De, De, We4, nq, q));
                acc = fma4(alpha[(int64_t)p * H + hh], m, acc);
            }
            *reinterpret_cast<float4*>(out + 4 * q) = acc;
        }
        if (lane < H) {
            float gsj = 0.f;
            for (int k = beg; k < end; ++k) gsj += g_logit[(int64_t)src_pos[k] * H + lane];
            out[HC + H + lane] = gsj;
        }
        for (int k = HC + 2 * H + lane; k < a.ld; k += 32) out[k] = 0.f;
    }
}

// a.rowptr = src_rowptr, a.other = src_dst, a.ld = ld of g_xpe; the staged rows are g_agg rows (HC floats)
template <int H, int CPI, bool USE_EP>
__global__ void __launch_bounds__(kWinThreads, kWinCtasPerSM)
edge_win_bwd_src_kernel(const __grid_constant__ WinArgs a, const int32_t* __restrict__ src_pos, const float* __restrict__ alpha,
                        const float* __restrict__ g_agg, const float* __restrict__ g_logit, float* __restrict__ g_xpe) {
    extern __shared__ __align__(128) float smem_f[];
    const int HC = H * a.C, nq = HC >> 2, De = a.De;
    const ItemGeom ig{nq / CPI, (a.C >> 2) / CPI};
    WinSmem s = win_carve(smem_f, a.rmax, HC, De, nq, H, USE_EP);
    int2* rec_st = reinterpret_cast<int2*>(s.rec);
    float* rec_c = s.rec + 2 * kWinMaxEdges;                      // [kWinMaxEdges][H]
    float* rec_gl = rec_c + kWinMaxEdges * H;                     // [kWinMaxEdges][H]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (USE_EP)
        for (int i = tid; i < De * nq; i += kWinThreads) s.We4[i] = ld4(a.w_edge + 4 * i);
    if (tid == 0) { tc::mbar_init(s.bar, 1); tc::fence_mbar_init(); }
    uint32_t parity = 0;
    const int tail = (int)a.ld - HC;                              // g_si | g_sj | pad columns of a g_xpe row
    __syncthreads();
    int4 desc = blockIdx.x < a.T ? a.tiles[blockIdx.x] : make_int4(0, 0, 0, 0);
    for (int64_t tile = blockIdx.x; tile < a.T; tile += gridDim.x) {
        const int64_t t0 = tile * a.D;
        const int nd = (int)min((int64_t)a.D, a.N - t0);
        const int lo = desc.x, nrows = desc.y - desc.x, k0 = desc.z, ne = desc.w - desc.z;
        const int64_t nxt = tile + gridDim.x;
        if (nxt < a.T) desc = a.tiles[nxt];
        const bool overflow = ne > kWinMaxEdges;
        const bool near = !overflow && nrows <= a.rmax && ne > 0;
        if (near && warp == 0) {
            tc::fence_proxy_async_smem();
            bulk_window(s.rows, g_agg + (int64_t)lo * HC, (uint32_t)nrows * (uint32_t)HC * 4u, s.bar, lane);
        }
        if (tid <= nd) s.rp[tid] = a.rowptr[t0 + tid] - k0;
        if (overflow) {
            __syncthreads();
            src_overflow_tile<H, USE_EP>(a, s.We4, s.rp, nd, t0, k0, src_pos, alpha, g_agg, g_logit, g_xpe);
            __syncthreads();
            continue;
        }
        const int base = near ? lo : 0;
        for (int e = tid; e < ne; e += kWinThreads) {
            const int k = k0 + e;
            const int p = src_pos[k];
            int nz = 1, ty = 0;
            float val = 1.f;
            if (USE_EP) {
                const float* earow = a.ea + (int64_t)p * De;
                nz = 0;
                for (int d = 0; d < De; ++d) {
                    const float v = earow[d];
                    if (v != 0.f) { ++nz; ty = d; val = v; }
                }
            }
            rec_st[e] = make_int2(a.other[k] - base, nz == 1 ? ty : -1);
            const float sc = nz == 1 ? val : 1.f;
#pragma unroll
            for (int h = 0; h < H; ++h) {
                rec_c[e * H + h] = alpha[(int64_t)p * H + h] * sc;
                rec_gl[e * H + h] = g_logit[(int64_t)p * H + h];
            }
        }
        __syncthreads();
        // g_sj and the pad columns (g_si, the first H tail columns, belongs to the destination pass)
        for (int i = tid; i < nd * tail; i += kWinThreads) {
            const int d = i / tail, c = i - d * tail;
            if (c >= H) {
                float v = 0.f;
                if (c < 2 * H)
                    for (int e = s.rp[d]; e < s.rp[d + 1]; ++e) v += rec_gl[e * H + (c - H)];
                g_xpe[(t0 + d) * a.ld + HC + c] = v;
            }
        }
        if (near) {
            tc::mbar_wait(s.bar, parity);
            parity ^= 1u;
            win_aggregate<H, CPI, USE_EP, int>(s.rows, HC, nq, ig, nd, t0, De, a.ea, src_pos, k0, s.We4, s.rp, rec_st, rec_c, g_xpe, a.ld);
        } else {
            win_aggregate<H, CPI, USE_EP, int64_t>(g_agg, HC, nq, ig, nd, t0, De, a.ea, src_pos, k0, s.We4, s.rp, rec_st, rec_c, g_xpe, a.ld);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ host side
#define GLAM_WIN_CPI(H_, EP_, cpi, ...)                                            \
    switch (cpi) {                                                                 \
        case 1: { constexpr int HH_ = H_, CPI_ = 1; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
        case 2: { constexpr int HH_ = H_, CPI_ = 2; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
        default: { constexpr int HH_ = H_, CPI_ = 3; constexpr bool UE_ = EP_; __VA_ARGS__; } break; \
    }
#define GLAM_WIN_DISPATCH(heads, use_ep, cpi, ...)                      \
    if (!(use_ep)) { GLAM_WIN_CPI(1, false, cpi, __VA_ARGS__) }         \
    else switch (heads) {                                               \
        case 1: GLAM_WIN_CPI(1, true, cpi, __VA_ARGS__) break;          \
        case 2: GLAM_WIN_CPI(2, true, cpi, __VA_ARGS__) break;          \
        case 3: GLAM_WIN_CPI(3, true, cpi, __VA_ARGS__) break;          \
        default: GLAM_WIN_CPI(4, true, cpi, __VA_ARGS__) break;         \
    }

// rows of `ldrow` floats that fit beside the fixed part of the budget
static int win_rmax(size_t budget, int64_t ldrow, int De, int nq, bool use_ep, int rec_floats, int extra_floats) {
    const size_t fixed = win_smem_bytes(0, ldrow, De, nq, use_ep, rec_floats, extra_floats);
    if (fixed + 1024 >= budget) return 0;
    return (int)((budget - fixed) / (sizeof(float) * (size_t)ldrow));
}

bool edge_win_eligible(int heads, int C, int De, int64_t ldxp) {
    if ((C & 3) || (ldxp & 3) || De > kWinMaxDe || heads < 1 || heads > GLAM_MAX_HEADS) return false;
    return heads * C <= 384;
}

template <typename F>
static void win_allow_smem(F fn, size_t bytes) {
    ensure_dyn_smem((const void*)fn, (size_t)((int)bytes));
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static int win_grid(int64_t T, int ctas_per_sm) {
    int64_t g = (int64_t)kNumSMs * ctas_per_sm;
    if (g > T) g = T;
    return (int)(g < 1 ? 1 : g);
}

int edge_win_build_tiles(const int32_t* dst_rowptr, const int32_t* dst_src, const int32_t* src_rowptr, const int32_t* src_dst,
                         int64_t N, int32_t* dst_tiles, int32_t* src_tiles, cudaStream_t stream) {
    const int D = edge_tile_rows(N);
    const int64_t T = (N + D - 1) / D;
    if (T == 0) return 0;
    const int grid = (int)((T * 32 + 255) / 256);
    edge_tiles_kernel<<<grid, 256, 0, stream>>>(dst_rowptr, dst_src, N, D, T, 1, reinterpret_cast<int4*>(dst_tiles));
    GLAM_CHECK_LAUNCH();
    if (src_tiles) {
        edge_tiles_kernel<<<grid, 256, 0, stream>>>(src_rowptr, src_dst, N, D, T, 0, reinterpret_cast<int4*>(src_tiles));
        GLAM_CHECK_LAUNCH();
    }
    return 0;
}

// *launched = 1 when the windowed kernel took the call, 0 when the configuration does not fit (caller takes the gather path)
int edge_win_fwd(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge, const int32_t* rowptr,
                 const int32_t* srcs, const int32_t* tiles, int64_t N, int heads, int C, int De, float slope, float* agg, float* alpha,
                 cudaStream_t stream, int* launched) {
    *launched = 0;
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4, cpi = pick_cpi(C), ni = nq / cpi;
    const int D = edge_tile_rows(N);
    const int rec = 3 + heads;
    const int64_t T = (N + D - 1) / D;
    // pipelined kernel: two window buffers, block size fitted to the (node, item) work of a tile; configurations that do not
    // fit (very wide rows) leave *launched = 0 and the caller takes the per-edge gather kernels
    const int rmax2 = win_rmax(kWin2SmemBudget, 2 * ldxp, De, nq, use_ep, rec, 0);
    const int items = D * ni, passes = (items + kWin2MaxThreads - 1) / kWin2MaxThreads;
    int nthr = ((items + passes - 1) / passes + 31) / 32 * 32;
    if (nthr < 256) nthr = 256;
    if (rmax2 < D + 24 || nthr > kWin2MaxThreads || ni > nthr) return 0;
    WinArgs a{xpe, ldxp, ea, w_edge, att_edge, rowptr, srcs, reinterpret_cast<const int4*>(tiles), N, C, De, D, rmax2, T, slope};
    const size_t smem = win_smem_bytes(2 * rmax2, ldxp, De, nq, use_ep, rec, 0);
    GLAM_WIN_DISPATCH(heads, use_ep, cpi, {
        auto fn = edge_win2_fwd_kernel<HH_, CPI_, UE_>;
        win_allow_smem(fn, smem);
        fn<<<win_grid(T, kWin2CtasPerSM), nthr, smem, stream>>>(a, agg, alpha);
    })
    GLAM_CHECK_LAUNCH();
    *launched = 1;
    return 0;
}

int edge_win_bwd_dst_grid(int64_t N) { return win_grid(edge_tile_count(N), kWinDstCtasPerSM); }

int edge_win_bwd_dst(const float* xpe, int64_t ldxp, const float* ea, const float* w_edge, const float* att_edge, const float* alpha,
                     const float* g_agg, const int32_t* rowptr, const int32_t* srcs, const int32_t* tiles, int64_t N, int heads, int C,
                     int De, float slope, float* g_logit, float* g_xpe, float* gwe_partial, cudaStream_t stream,
                     int* launched) {
    *launched = 0;
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4, cpi = pick_cpi(C), ni = nq / cpi;
    if (De > kWinRegDe || ni > 32) return 0;
    const int D = edge_tile_rows(N);
    // pipelined kernel: cp.async staging of the next tile's raw words, next window issued behind the dots
    const size_t fixed = sizeof(float) * dst2_fixed_floats(De, nq, heads, use_ep);
    const size_t stage2 = sizeof(float4) * (size_t)kWinWarps * (32 / ni) * De * nq;
    int rmax2 = fixed + 1024 < (size_t)kWinDstSmemBudget ? (int)(((size_t)kWinDstSmemBudget - fixed) / (sizeof(float) * (size_t)ldxp)) : 0;
    if (rmax2 < D + 24 || (use_ep && (size_t)rmax2 * ldxp * 4 < stage2)) return 0;
    WinArgs a{xpe, ldxp, ea, w_edge, att_edge, rowptr, srcs, reinterpret_cast<const int4*>(tiles), N, C, De, D, rmax2, (N + D - 1) / D,
              slope};
    const size_t smem = fixed + sizeof(float) * (size_t)rmax2 * ldxp;
    if (De <= 3) {
        GLAM_WIN_DISPATCH(heads, use_ep, cpi, {
            auto fn = edge_win_bwd_dst2_kernel<HH_, CPI_, UE_, 3>;
            win_allow_smem(fn, smem);
            fn<<<win_grid(a.T, kWinDstCtasPerSM), kWinThreads, smem, stream>>>(a, alpha, g_agg, g_logit, g_xpe, gwe_partial);
        })
    } else {
        GLAM_WIN_DISPATCH(heads, use_ep, cpi, {
            auto fn = edge_win_bwd_dst2_kernel<HH_, CPI_, UE_, kWinRegDe>;
            win_allow_smem(fn, smem);
            fn<<<win_grid(a.T, kWinDstCtasPerSM), kWinThreads, smem, stream>>>(a, alpha, g_agg, g_logit, g_xpe, gwe_partial);
        })
    }
    GLAM_CHECK_LAUNCH();
    *launched = 1;
    return 0;
}

int edge_win_bwd_src(const float* ea, const float* w_edge, const float* alpha, const float* g_agg, const float* g_logit,
                     const int32_t* src_rowptr, const int32_t* src_pos, const int32_t* src_dst, const int32_t* tiles, int64_t N, int heads,
                     int C, int De, float* g_xpe, int64_t ldxp, cudaStream_t stream, int* launched) {
    *launched = 0;
    const bool use_ep = w_edge != nullptr;
    const int HC = heads * C, nq = HC / 4, cpi = pick_cpi(C);
    const int D = edge_tile_rows(N);
    const int rec = 2 + 2 * heads;
    const int rmax = win_rmax(kWinSmemBudget, HC, De, nq, use_ep, rec, 0);
    if (rmax < D + 8 || nq / cpi > kWinThreads) return 0;
    WinArgs a{nullptr, ldxp, ea, w_edge, nullptr, src_rowptr, src_dst, reinterpret_cast<const int4*>(tiles), N, C, De, D, rmax,
              (N + D - 1) / D, 0.f};
    const size_t smem = win_smem_bytes(rmax, HC, De, nq, use_ep, rec, 0);
    GLAM_WIN_DISPATCH(heads, use_ep, cpi, {
        auto fn = edge_win_bwd_src_kernel<HH_, CPI_, UE_>;
        win_allow_smem(fn, smem);
        fn<<<win_grid(a.T, kWinCtasPerSM), kWinThreads, smem, stream>>>(a, src_pos, alpha, g_agg, g_logit, g_xpe);
    })
    GLAM_CHECK_LAUNCH();
    *launched = 1;
    return 0;
}

}  // namespace glam
