// Projections on the 5th-generation tensor cores (tcgen05, TF32 operands, fp32 accumulation in TMEM).
//
//   Y[M,N] = epilogue(X[M,K] * W + bias)        M = nodes (10^5..10^8), K,N = 36..288
//
// The contraction is tiny per row, so this is a streaming kernel whose job is to keep HBM busy.  One persistent
// CTA per SM, warp-specialised:
//   warp 0  (1 lane)  TMA producer: cp.async.bulk.tensor loads of 128-row X tiles (32-feature panels, hardware
//                     SWIZZLE_128B, out-of-bounds rows/columns zero-filled) into a 2..4-stage shared-memory ring;
//   warp 1  (1 lane)  MMA issuer: ceil(K/8) tcgen05.mma.kind::tf32 per tile into one of two TMEM accumulators,
//                     tcgen05.commit releases the smem stage and publishes the accumulator;
//   warps 2-5         epilogue: tcgen05.ld (thread = row = TMEM lane) -> bias / CELU / CELU' mask / accumulate ->
//                     16-byte global stores; they also stage W once (K-major swizzled image) at kernel start.
// Loads of tile i+2.., MMA of tile i+1 and the epilogue of tile i overlap; the accumulator is double-buffered.
//
// Output columns [exact_begin, exact_end) bypass the tensor core: the attention-logit columns s_i|s_j of the
// extended node projection are computed with exact fp32 FMAs from the staged operands (see glam_gemm_ex).
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace glam {

using namespace tc;

constexpr int kEpiGroups = 2;                          // two 4-warp epilogue groups, one per TMEM accumulator
constexpr int kTcThreads = 64 + kEpiGroups * 128;
constexpr int kTileM = 128;
constexpr int kMaxStages = 4;
constexpr int kPanelBytes = kTileM * kPanelRowBytes;   // 16 KB: one 32-feature panel of a 128-row tile

enum Epi { EPI_NONE = 0, EPI_CELU = 1, EPI_MUL_CELU_GRAD = 2, EPI_ACCUM = 3 };

struct TcGemmParams {
    const float* W; int64_t w_sk, w_sn;
    const float* bias;
    const float* aux; int64_t ldaux;
    float* Y; int64_t ldy;
    int64_t M; int N, K, epi;
    int KP, Npad, tmem_cols, vec_store, stages, staged_out, og_groups, w_vec;
    int exact_begin, exact_end;   // output columns computed with exact fp32 FMAs (attention-logit columns)
    int n_chunk;                  // > 0: blockIdx.y owns output columns [y * n_chunk, ...) (few row tiles: fill the SMs by columns)
    unsigned long long* dbg;      // optional phase timestamps of CTA 0 (globaltimer ns), 8 slots
};


__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GLAM_DBG(slot) do { if (p.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.dbg[slot] = gtime(); } while (0)

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

template <int EPI>
__device__ __forceinline__ float4 epi_apply(float4 v, float4 b, float4 y) {
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if (EPI == EPI_CELU) { v.x = celu1(v.x); v.y = celu1(v.y); v.z = celu1(v.z); v.w = celu1(v.w); }
    else if (EPI == EPI_MUL_CELU_GRAD) {
        v.x *= (y.x > 0.f ? 1.f : y.x + 1.f); v.y *= (y.y > 0.f ? 1.f : y.y + 1.f);
        v.z *= (y.z > 0.f ? 1.f : y.z + 1.f); v.w *= (y.w > 0.f ? 1.f : y.w + 1.f);
    } else if (EPI == EPI_ACCUM) { v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w; }
    return v;
}

// The whole tile loop of one epilogue group, specialised on the epilogue so that the inner loops carry no mode branches.
// Phase 1 (thread = row = TMEM lane): tcgen05.ld -> row-major staging tile; exact fp32 logit columns from the staged
// operands.  Phase 2: the group's four warps sweep the tile with fully coalesced 16-byte global accesses.
template <int EPI>
__device__ __forceinline__ void staged_epilogue_loop(const TcGemmParams& p, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                     uint64_t* empty_bar, uint32_t tmem_base, uint32_t xs_addr, uint32_t ws_addr,
                                                     uint32_t og_addr, int stage_bytes, int grp, int wg, int q4, int lane,
                                                     int64_t ntiles) {
    const bool has_exact = p.exact_end > p.exact_begin;
    const int nex = has_exact ? min(p.exact_end - p.exact_begin, 8) : 0;
    const int row_in_tile = q4 * 32 + lane;
    const int KQ = p.K >> 2, nq = p.N >> 2, N = p.N;
    const uint32_t my_row = og_addr + (uint32_t)(row_in_tile * N) * 4u;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % p.stages, a = it & 1;
        if (a != grp) continue;
        mbar_wait(&tfull_bar[a], (uint32_t)(it >> 1) & 1u);
        if (it == 0 && wg == 0) GLAM_DBG(4);
        tc_fence_after_sync();
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * p.Npad);
        // ---- phase 1
        float xs[8];
        if (nex > 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) xs[e] = 0.f;
            const uint32_t xrow = xs_addr + (uint32_t)(s * stage_bytes);
            for (int q = 0; q < KQ; ++q) {
                const float4 xv = lds128(xrow + panel_chunk_offset(row_in_tile, q, kTileM));
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < nex) {
                        const float4 wv = lds128(ws_addr + panel_chunk_offset(p.exact_begin + e, q, p.Npad));
                        xs[e] = fmaf(xv.x, wv.x, xs[e]); xs[e] = fmaf(xv.y, wv.y, xs[e]);
                        xs[e] = fmaf(xv.z, wv.z, xs[e]); xs[e] = fmaf(xv.w, wv.w, xs[e]);
                    }
            }
        }
        for (int c0 = 0; c0 < p.Npad; c0 += 32) {
            float v[32];
            tmem_ld32(lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (c0 + j < N) sts128(my_row + (uint32_t)(c0 + j) * 4u, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
        if (nex > 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (e < nex) sts32(my_row + (uint32_t)(p.exact_begin + e) * 4u, xs[e]);
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {                                     // accumulator and X stage are free again
            mbar_arrive(&tempty_bar[a]);
            if (has_exact) mbar_arrive(&empty_bar[s]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        // ---- phase 2
        const int64_t row0 = tile * kTileM;
        const int64_t rows_left = p.M - row0;
        const int rows_here = rows_left < kTileM ? (int)rows_left : kTileM;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nq >= 24) {
            // wide rows: warp per row, lanes over the row's 16-byte chunks (bias hoisted, pointer increments only)
            for (int cq = lane; cq < nq; cq += 32) {
                const int c = cq << 2;
                const float4 b = p.bias ? *reinterpret_cast<const float4*>(p.bias + c) : zero4;
                uint32_t og = og_addr + (uint32_t)(wg * N + c) * 4u;
                float* y = p.Y + (row0 + wg) * p.ldy + c;
                const float* ax = (EPI == EPI_MUL_CELU_GRAD) ? p.aux + (row0 + wg) * p.ldaux + c : nullptr;
                const uint32_t og_step = (uint32_t)(4 * N) * 4u;
                const int64_t y_step = 4 * p.ldy, a_step = 4 * p.ldaux;
#pragma unroll 4
                for (int r = wg; r < rows_here; r += 4) {
                    float4 yv = zero4;
                    if (EPI == EPI_MUL_CELU_GRAD) yv = *reinterpret_cast<const float4*>(ax);
                    else if (EPI == EPI_ACCUM) yv = *reinterpret_cast<const float4*>(y);
                    *reinterpret_cast<float4*>(y) = epi_apply<EPI>(lds128(og), b, yv);
                    og += og_step; y += y_step;
                    if (EPI == EPI_MUL_CELU_GRAD) ax += a_step;
                }
            }
        } else {
            // narrow rows: flat sweep over the tile's 16-byte chunks, 4 independent chunks per thread in flight
            const int total = rows_here * nq;
            const float inv_nq = 1.0f / (float)nq;
            for (int i0 = wg * 32 + lane; i0 < total; i0 += 128 * 4) {
                float4 v[4], yv[4], b[4];
                float* yp[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 128 * u;
                    yv[u] = zero4; b[u] = zero4; yp[u] = nullptr;
                    if (i < total) {
                        const int r = (int)(((float)i + 0.5f) * inv_nq);      // exact for i < 2^13
                        const int c = (i - r * nq) << 2;
                        v[u] = lds128(og_addr + (uint32_t)i * 16u);
                        yp[u] = p.Y + (row0 + r) * p.ldy + c;
                        if (p.bias) b[u] = *reinterpret_cast<const float4*>(p.bias + c);
                        if (EPI == EPI_MUL_CELU_GRAD) yv[u] = *reinterpret_cast<const float4*>(p.aux + (row0 + r) * p.ldaux + c);
                        else if (EPI == EPI_ACCUM) yv[u] = *reinterpret_cast<const float4*>(yp[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (yp[u]) *reinterpret_cast<float4*>(yp[u]) = epi_apply<EPI>(v[u], b[u], yv[u]);
            }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // staging tile free for this group's next tile
        if (it == 0 && wg == 0) GLAM_DBG(5);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const TcGemmParams p_in) {
    TcGemmParams p = p_in;
    if (p.n_chunk > 0) {                                         // this CTA's column block: every N-side pointer moves with it
        const int n0 = blockIdx.y * p.n_chunk;
        p.W += (int64_t)n0 * p.w_sn;
        if (p.bias) p.bias += n0;
        if (p.aux) p.aux += n0;
        p.Y += n0;
        p.N = min(p.n_chunk, p.N - n0);
        p.Npad = (p.N + 15) / 16 * 16;
    }
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int panels = (p.KP + kPanelFeatures - 1) / kPanelFeatures;
    const int stage_bytes = panels * kPanelBytes;
    uint8_t* Xs = smem;                                          // [stages][panels][128][128 B]
    uint8_t* Ws = smem + (size_t)p.stages * stage_bytes;         // [panels][Npad][128 B]
    float* Os = reinterpret_cast<float*>(Ws + (size_t)panels * p.Npad * kPanelRowBytes);   // [128][N] output staging
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const bool has_exact = p.exact_end > p.exact_begin;
    if (t == 0) GLAM_DBG(0);

    if (t == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], has_exact ? 5 : 1); }   // MMA commit (+ 4 epilogue warps)
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        mbar_init(&w_bar, 4 * kEpiGroups);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t ntiles = (p.M + kTileM - 1) / kTileM;
    if (t == 0) GLAM_DBG(1);

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            prefetch_tensormap(&tmap_x);
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                for (int pn = 0; pn < panels; ++pn)
                    tma_load_2d(Xs + (size_t)s * stage_bytes + (size_t)pn * kPanelBytes, &tmap_x, pn * kPanelFeatures,
                                (int)(tile * kTileM), &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(kTileM, p.Npad, 0, 0);
            const uint32_t xs_addr = smem_u32(Xs), ws_addr = smem_u32(Ws);
            const int ksteps = p.KP >> 3;
            mbar_wait(&w_bar, 0);
            GLAM_DBG(2);
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int s = it % p.stages, a = it & 1;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u, aph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&tempty_bar[a], aph ^ 1u);
                mbar_wait(&full_bar[s], ph);
                if (it == 0) GLAM_DBG(3);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * p.Npad);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t pan = ks >> 2, within = (ks & 3) * 32;
                    const uint64_t da = make_smem_desc(xs_addr + s * stage_bytes + pan * kPanelBytes + within, 16, 1024);
                    const uint64_t db = make_smem_desc(ws_addr + pan * (p.Npad * kPanelRowBytes) + within, 16, 1024);
                    mma_tf32_ss(d_tmem, da, db, idesc, ks > 0 ? 1u : 0u);
                }
                mma_commit(&empty_bar[s]);       // smem stage free once these MMAs have read it
                mma_commit(&tfull_bar[a]);       // accumulator ready
            }
        }
    } else {
        // ================================ epilogue warps ===============================
        const int eall = t - 64;                                 // 0 .. 128*kEpiGroups-1 (W staging)
        const int grp = (warp - 2) >> 2;                         // epilogue group = TMEM accumulator index
        const int wg = (warp - 2) & 3;                           // warp within the group
        const int q4 = warp & 3;                                 // TMEM lane quarter this warp may read
        const int row_in_tile = q4 * 32 + lane;
        float* Og = Os + (size_t)(p.og_groups == 1 ? 0 : grp) * kTileM * p.N;   // this group's staging tile (one CTA-tile launches only use group 0)
        // ---- stage W once: K-major swizzled image [panels][Npad][128 B], zero padded to [Npad][KP]
        {
            const int KPQ = p.KP >> 2;
            const int items = p.Npad * KPQ;
            constexpr int kWT = 128 * kEpiGroups;
            constexpr int kWU = 8;                               // loads in flight per thread
            for (int base = 0; base < items; base += kWT * kWU) {
                float4 v[kWU];
#pragma unroll
                for (int u = 0; u < kWU; ++u) {
                    const int idx = base + u * kWT + eall;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < items) {
                        int n, q;
                        if (p.w_sn == 1) { q = idx / p.Npad; n = idx - q * p.Npad; } else { n = idx / KPQ; q = idx - n * KPQ; }
                        if (n < p.N) {
                            const int k = 4 * q;
                            const float* w = p.W + (int64_t)k * p.w_sk + (int64_t)n * p.w_sn;
                            if (p.w_vec) { if (k < p.K) v[u] = __ldg(reinterpret_cast<const float4*>(w)); continue; }   // K % 4 == 0
                            if (k < p.K) v[u].x = w[0];
                            if (k + 1 < p.K) v[u].y = w[p.w_sk];
                            if (k + 2 < p.K) v[u].z = w[2 * p.w_sk];
                            if (k + 3 < p.K) v[u].w = w[3 * p.w_sk];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kWU; ++u) {
                    const int idx = base + u * kWT + eall;
                    if (idx < items) {
                        int n, q;
                        if (p.w_sn == 1) { q = idx / p.Npad; n = idx - q * p.Npad; } else { n = idx / KPQ; q = idx - n * KPQ; }
                        *reinterpret_cast<float4*>(Ws + panel_chunk_offset(n, q, p.Npad)) = v[u];
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_bar);
            if (has_exact) {                                     // the exact columns read W rows from shared memory
                mbar_wait(&w_bar, 0);
                // (redundant with the mbarrier's release/acquire; named barrier 3 over the 256 epilogue threads makes the ordering
                // visible to compute-sanitizer's racecheck, which does not model inline-PTX mbarrier waits)
                asm volatile("bar.sync 3, %0;" ::"n"(128 * kEpiGroups) : "memory");
            }
        }
        if (p.staged_out) {
            const uint32_t xs_a = smem_u32(Xs), ws_a = smem_u32(Ws), og_a = smem_u32(Og);
            switch (p.epi) {
                case EPI_CELU: staged_epilogue_loop<EPI_CELU>(p, tfull_bar, tempty_bar, empty_bar, tmem_base, xs_a, ws_a, og_a, stage_bytes, grp, wg, q4, lane, ntiles); break;
                case EPI_MUL_CELU_GRAD: staged_epilogue_loop<EPI_MUL_CELU_GRAD>(p, tfull_bar, tempty_bar, empty_bar, tmem_base, xs_a, ws_a, og_a, stage_bytes, grp, wg, q4, lane, ntiles); break;
                case EPI_ACCUM: staged_epilogue_loop<EPI_ACCUM>(p, tfull_bar, tempty_bar, empty_bar, tmem_base, xs_a, ws_a, og_a, stage_bytes, grp, wg, q4, lane, ntiles); break;
                default: staged_epilogue_loop<EPI_NONE>(p, tfull_bar, tempty_bar, empty_bar, tmem_base, xs_a, ws_a, og_a, stage_bytes, grp, wg, q4, lane, ntiles); break;
            }
        } else {
        const int KQ = p.K >> 2;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % p.stages, a = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            if (a != grp) continue;
            mbar_wait(&tfull_bar[a], aph);
            tc_fence_after_sync();
            const int64_t m = tile * kTileM + row_in_tile;
            const bool row_ok = m < p.M;
            float* yrow = p.Y + m * p.ldy;
            const float* arow = p.aux ? p.aux + m * p.ldaux : nullptr;
            const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * p.Npad);
            const uint8_t* Xrow = Xs + (size_t)s * stage_bytes;
            for (int c0 = 0; c0 < p.Npad; c0 += 16) {
                float v[16];
                tmem_ld16(lane_base + (uint32_t)c0, v);
                if (!row_ok) continue;
                if (c0 < p.exact_end && c0 + 16 > p.exact_begin) {
                    for (int n = max(c0, p.exact_begin); n < min(c0 + 16, p.exact_end); ++n) {
                        float acc = 0.f;
                        for (int q = 0; q < KQ; ++q) {
                            const float4 xv = *reinterpret_cast<const float4*>(Xrow + panel_chunk_offset(row_in_tile, q, kTileM));
                            const float4 wv = *reinterpret_cast<const float4*>(Ws + panel_chunk_offset(n, q, p.Npad));
                            acc = fmaf(xv.x, wv.x, acc); acc = fmaf(xv.y, wv.y, acc);
                            acc = fmaf(xv.z, wv.z, acc); acc = fmaf(xv.w, wv.w, acc);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (c0 + j == n) v[j] = acc;
                    }
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = c0 + j;
                    if (n < p.N) {
                        float acc = v[j];
                        if (p.bias) acc += p.bias[n];
                        if (p.epi == EPI_CELU) acc = celu1(acc);
                        else if (p.epi == EPI_MUL_CELU_GRAD) { float y = arow[n]; acc *= (y > 0.f ? 1.f : y + 1.f); }
                        else if (p.epi == EPI_ACCUM) acc += yrow[n];
                        v[j] = acc;
                    }
                }
                if (p.vec_store && c0 + 16 <= p.N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.N) yrow[c0 + j] = v[j];
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tempty_bar[a]);
                if (has_exact) mbar_arrive(&empty_bar[s]);
            }
        }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (t == 64) GLAM_DBG(6);
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (t == 32) GLAM_DBG(7);
}

static unsigned long long* g_tc_dbg = nullptr;
static int g_math_mode = 1;   // 0 = fp32 on the CUDA cores (exact), 1 = TF32 tensor cores (default)
int g_math_mode_get() { return g_math_mode; }

// ---- host: tensor map encoding through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D fp32 row-major [rows, cols] with leading dimension ld; box = box_rows x 32 features
int make_tmap_rows(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return -1; }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kPanelFeatures, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows, (long long)cols, (long long)ld);
        return -1;
    }
    return 0;
}

// Shared-memory plan: W image + (optionally) the [128][N] output staging tile + as many X stages as fit (<= 4).
// single_tile: every CTA processes at most one tile (M <= 128 * grid), so one load stage and one epilogue group's
// staging tile are enough — that keeps the coalesced epilogue for wide N x K that would not fit otherwise.
static size_t tc_smem_bytes(int64_t N, int64_t K, bool want_staged, int* stages_out, int* staged_out, bool single_tile = false) {
    const int64_t KP = (K + 7) / 8 * 8, Npad = (N + 15) / 16 * 16;
    const int64_t panels = (KP + kPanelFeatures - 1) / kPanelFeatures;
    const int64_t wbytes = panels * Npad * kPanelRowBytes, stage = panels * kPanelBytes;
    const int64_t obytes = (int64_t)(single_tile ? 1 : kEpiGroups) * kTileM * N * sizeof(float);
    const int64_t limit = 222 * 1024;
    int64_t staged = want_staged ? 1 : 0;
    int64_t stages = (limit - wbytes - staged * obytes) / stage;
    if (staged && stages < (single_tile ? 1 : 2)) { staged = 0; stages = (limit - wbytes) / stage; }   // keep the load pipeline alive first
    if (single_tile && stages > 1) stages = 1;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages_out) *stages_out = (int)stages;
    if (staged_out) *staged_out = (int)staged;
    return stages < 1 ? 0 : (size_t)(wbytes + staged * obytes + stages * stage + 1024);
}

bool tc_gemm_eligible(const float* X, int64_t ldx, const float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K) {
    if (g_math_mode == 0) return false;
    if (N > 256 || K > 288 || K < 4 || (K & 3) || (ldx & 3) || ((uintptr_t)X & 15)) return false;
    if (M < 1 || M >= ((int64_t)1 << 31)) return false;
    if (tc_smem_bytes(N, K, false, nullptr, nullptr) == 0) return false;  // operands must fit in shared memory
    (void)Y; (void)ldy;
    return true;
}

int tc_gemm_launch(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                   const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, int epi,
                   int exact_begin, int exact_end, cudaStream_t stream) {
    TcGemmParams p;
    p.exact_begin = exact_begin; p.exact_end = exact_end;
    p.dbg = g_tc_dbg;
    p.W = W; p.w_sk = w_sk; p.w_sn = w_sn; p.bias = bias; p.aux = aux; p.ldaux = ldaux;
    p.Y = Y; p.ldy = ldy; p.M = M; p.N = (int)N; p.K = (int)K; p.epi = epi;
    p.KP = (int)((K + 7) / 8 * 8);
    p.Npad = (int)((N + 15) / 16 * 16);
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)(p.Npad + (p.Npad + 31) / 32 * 32));   // epilogue reads 32-column groups
    p.vec_store = ((ldy & 3) == 0 && ((uintptr_t)Y & 15) == 0) ? 1 : 0;
    const bool al16 = ((N & 3) == 0) && ((ldy & 3) == 0) && (((uintptr_t)Y & 15) == 0) && (((uintptr_t)bias & 15) == 0) &&
                      (aux == nullptr || (((ldaux & 3) == 0) && (((uintptr_t)aux & 15) == 0)));
    const bool single_tile = M <= (int64_t)kTileM * kNumSMs;
    size_t smem = tc_smem_bytes(N, K, al16, &p.stages, &p.staged_out, single_tile);
    p.og_groups = single_tile ? 1 : kEpiGroups;
    // Many row tiles, but the two [128][N] staging tiles of the coalesced epilogue do not fit beside W and the load ring
    // (N = 144, K = 108: Set2Set's gate GEMM over 65 536 screening graphs ran the row-per-thread epilogue at 1 TB/s): split the
    // output columns over blockIdx.y until they do — X tiles are then read once per column block, out of L2.
    p.n_chunk = 0;
    unsigned gy = 1;
    if (!single_tile && al16 && !p.staged_out && exact_end <= exact_begin && N >= 64 && smem > 0) {
        for (int ways = 2; ways <= 4; ++ways) {
            const int64_t chunk = ((N + ways - 1) / ways + 15) / 16 * 16;
            int st = 0, sg = 0;
            const size_t sm = tc_smem_bytes(chunk, K, true, &st, &sg, false);
            if (sm > 0 && sg == 1 && st >= 2 && chunk < N) {
                p.n_chunk = (int)chunk; gy = (unsigned)((N + chunk - 1) / chunk);
                p.stages = st; p.staged_out = 1; smem = sm;
                break;
            }
        }
    }
    p.w_vec = (w_sk == 1 && (w_sn & 3) == 0 && ((uintptr_t)W & 15) == 0) ? 1 : 0;
    GLAM_REQUIRE(smem > 0, "tc_gemm: operands do not fit in shared memory (N=%lld K=%lld)", (long long)N, (long long)K);
    CUtensorMap tmap;
    if (int rc = make_tmap_rows(&tmap, X, M, K, ldx, kTileM, (int)CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    {                                                            // the attribute is per device: set it on every call
        cudaError_t e = ensure_dyn_smem((const void*)tc_gemm_kernel, (size_t)((int)smem));
        if (e != cudaSuccess) { set_error("tc_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int64_t ntiles = (M + kTileM - 1) / kTileM;
    int64_t grid = kNumSMs;
    if (grid > ntiles) grid = ntiles;
    // Few row tiles (graph-level GEMMs: 4096 rows = 32 tiles): split the output columns over blockIdx.y so that the W image
    // build and the epilogue, which dominate such launches, shrink with it and more SMs take part.
    if (exact_end <= exact_begin && ntiles * 2 <= kNumSMs && N >= 64) {
        const int64_t ways = kNumSMs / ntiles;
        int64_t chunk = ((N + ways - 1) / ways + 15) / 16 * 16;
        if (chunk < 32) chunk = 32;
        if (chunk < N) { p.n_chunk = (int)chunk; gy = (unsigned)((N + chunk - 1) / chunk); }
    }
    tc_gemm_kernel<<<dim3((unsigned)grid, gy), kTcThreads, smem, stream>>>(tmap, p);
    GLAM_CHECK_LAUNCH();
    return 0;
}

}  // namespace glam

extern "C" int glam_set_math_mode(int mode) {
    GLAM_REQUIRE(mode == 0 || mode == 1, "glam_set_math_mode: 0 = fp32 (CUDA cores), 1 = tf32 (tcgen05)");
    glam::g_math_mode = mode;
    return 0;
}
extern "C" int glam_get_math_mode(void) { return glam::g_math_mode; }
// debugging aid (not part of the public header): device buffer of 8 u64 receiving CTA-0 phase timestamps of tc_gemm
extern "C" void glam_debug_tc_timestamps(void* buf) { glam::g_tc_dbg = (unsigned long long*)buf; }
