// GRU node update with the two gate GEMMs and the gate arithmetic in ONE kernel (MessageBlock, src_1gp/layer.py:260-266):
//
//   gi = m W_ih^T + b_ih,  gh = h W_hh^T + b_hh                 (torch.nn.GRU, gate order r, z, n)
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) n + z h, x_out = act(h' + identity)
//
// Unfused, the two [N,3C] pre-activation tensors are written by the GEMMs and read back by the gate kernel (that gate
// kernel runs at 82 % of HBM peak: there is nothing left to tune in it, only bytes to remove).  Here they never leave
// the SM: per 128-row tile TMA brings the m and h rows (32-feature SWIZZLE_128B panels), one lane issues
// tcgen05.mma.kind::tf32 for both products into ONE TMEM accumulator whose columns are permuted per channel to
// (r_pre, z_pre, gi_n, gh_n) — the h-side product accumulates onto the m-side one — double-buffered across tiles, and the
// epilogue warps (thread = row = TMEM lane) read 8 channels per tcgen05.ld.x32, apply bias / sigmoid / tanh / the
// residual and activation; r|z|n go straight to HBM, h' and gh_n are staged row-major in the (by then dead) operand
// stage and leave as contiguous tiles together with x_out = act(h' + identity) — row-strided 16-byte stores of those
// three cost 60 us per launch in L2 partial-line fills when tried.  HBM traffic per node row:
// 3C floats in, 6C out (vs 3C in + 6C out + 6C in + 6C out... unfused).
//
// Same warp roles and barriers as tc_gemm.cu: warp 0 TMA producer, warp 1 MMA issuer, then two epilogue groups of 8
// warps (two warps per TMEM lane quarter, interleaved over the channel chunks: the epilogue is a latency chain per row,
// so it wants every warp the register file allows).
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace glam {

using namespace tc;

int make_tmap_rows(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int swizzle);
int g_math_mode_get();

namespace {

constexpr int kGruEpiWarps = 8;                       // per epilogue group: two warps per TMEM lane quarter, channel chunks interleaved
constexpr int kGruThreads = 64 + 2 * kGruEpiWarps * 32;
constexpr int kGruTileM = 128;
constexpr int kGruStages = 2;
constexpr int kGruPanelBytes = kGruTileM * kPanelRowBytes;

struct TcGruParams {
    const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;
    const float* identity;
    float* rzn; float* gh; float* h_new; float* x_out;
    int64_t N; int C, KP, Npad, act; float act_param;
};

__device__ __forceinline__ float4 lds128g(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// The epilogue is a latency chain per row (few warps per SM), so the gate non-linearities use the fast intrinsics:
// ex2.approx / rcp.approx, relative error ~2^-22 on sigmoid, absolute ~2e-7 on tanh — far below TF32 operand rounding.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, fast_sigmoid(2.f * x), -1.f); }

// K-major swizzled images of the two [3C][C] row-major GRU weights, with the OUTPUT COLUMNS PERMUTED so that one TMEM
// accumulator holds, for channel c, the four values the gate arithmetic needs side by side:
//   column 4c+0 = r pre-activation (m W_ir + h W_hr), 4c+1 = z pre-activation, 4c+2 = gi_n = m W_in, 4c+3 = gh_n = h W_hn.
// The m-side image has zero rows at 4c+3, the h-side image at 4c+2; the second MMA accumulates onto the first.
__device__ __forceinline__ void stage_gru_weights(const float* __restrict__ w_ih, const float* __restrict__ w_hh, uint8_t* Wi,
                                                  uint8_t* Wh, int C, int KP, int Npad, int tid, int nthreads) {
    const int KPQ = KP >> 2, items = Npad * KPQ;
    for (int idx = tid; idx < items; idx += nthreads) {
        const int n = idx / KPQ, q = idx - n * KPQ;
        const int c = n >> 2, g = n & 3;
        float4 vi = make_float4(0.f, 0.f, 0.f, 0.f), vh = vi;
        if (c < C && 4 * q < C) {                                                    // C % 4 == 0
            if (g != 3) vi = __ldg(reinterpret_cast<const float4*>(w_ih + (size_t)(g * C + c) * C + 4 * q));
            if (g != 2) vh = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)((g == 3 ? 2 : g) * C + c) * C + 4 * q));
        }
        const uint32_t off = panel_chunk_offset(n, q, Npad);
        *reinterpret_cast<float4*>(Wi + off) = vi;
        *reinterpret_cast<float4*>(Wh + off) = vh;
    }
}

// CQ = C / 4 (compile time: the channel loop is fully unrolled, the residual row is prefetched into registers before the
// accumulator is ready, and no global load sits on the per-chunk dependency chain)
template <int CQ>
__global__ void __launch_bounds__(kGruThreads, 1)
tc_gru_fwd_kernel(const __grid_constant__ CUtensorMap tmap_m, const __grid_constant__ CUtensorMap tmap_h, const TcGruParams p) {
    constexpr int C = 4 * CQ;
    constexpr int NCH = (C + 7) / 8;                             // 8-channel (32-column) chunks
    constexpr int kJ = (NCH + 1) / 2;                            // chunks per warp (ch = 2j + half)
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kGruStages], empty_bar[kGruStages], tfull_bar[2], tempty_bar[2], w_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[4 * C];                // per channel: b_ir+b_hr, b_iz+b_hz, b_in, b_hn
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int panels = (p.KP + kPanelFeatures - 1) / kPanelFeatures;
    const int op_bytes = panels * kGruPanelBytes;               // one operand tile (m or h)
    const int stage_bytes = 2 * op_bytes;
    const int w_bytes = panels * p.Npad * kPanelRowBytes;
    uint8_t* Xs = smem;                                          // [stages][m | h][panels][128][128 B]
    uint8_t* Wi = smem + (size_t)kGruStages * stage_bytes;       // [panels][Npad][128 B]
    uint8_t* Wh = Wi + w_bytes;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) {
        for (int s = 0; s < kGruStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1 + kGruEpiWarps); }   // MMA commit + the group's epilogue warps (h rows are read from the stage)
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kGruEpiWarps); }
        mbar_init(&w_bar, 2 * kGruEpiWarps);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512u);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t ntiles = (p.N + kGruTileM - 1) / kGruTileM;

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tensormap(&tmap_m);
            prefetch_tensormap(&tmap_h);
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int s = it % kGruStages;
                const uint32_t ph = (uint32_t)(it / kGruStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                uint8_t* dst = Xs + (size_t)s * stage_bytes;
                for (int pn = 0; pn < panels; ++pn) {
                    tma_load_2d(dst + (size_t)pn * kGruPanelBytes, &tmap_m, pn * kPanelFeatures, (int)(tile * kGruTileM), &full_bar[s]);
                    tma_load_2d(dst + op_bytes + (size_t)pn * kGruPanelBytes, &tmap_h, pn * kPanelFeatures, (int)(tile * kGruTileM), &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(kGruTileM, p.Npad, 0, 0);
            const uint32_t xs_addr = smem_u32(Xs), wi_addr = smem_u32(Wi), wh_addr = smem_u32(Wh);
            const int ksteps = p.KP >> 3;
            mbar_wait(&w_bar, 0);
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int s = it % kGruStages, a = it & 1;
                const uint32_t ph = (uint32_t)(it / kGruStages) & 1u, aph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&tempty_bar[a], aph ^ 1u);
                mbar_wait(&full_bar[s], ph);
                tc_fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)(a * p.Npad);
                const uint32_t m_addr = xs_addr + s * stage_bytes, h_addr = m_addr + op_bytes;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t pan = ks >> 2, within = (ks & 3) * 32;
                    mma_tf32_ss(d, make_smem_desc(m_addr + pan * kGruPanelBytes + within, 16, 1024),
                                make_smem_desc(wi_addr + pan * (p.Npad * kPanelRowBytes) + within, 16, 1024), idesc, ks > 0 ? 1u : 0u);
                }
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t pan = ks >> 2, within = (ks & 3) * 32;
                    mma_tf32_ss(d, make_smem_desc(h_addr + pan * kGruPanelBytes + within, 16, 1024),
                                make_smem_desc(wh_addr + pan * (p.Npad * kPanelRowBytes) + within, 16, 1024), idesc, 1u);
                }
                mma_commit(&empty_bar[s]);
                mma_commit(&tfull_bar[a]);
            }
        }
    } else {
        // ================================ epilogue warps ===============================
        constexpr int kEpiThreads = 2 * kGruEpiWarps * 32;
        const int eall = t - 64;
        const int grp = (warp - 2) / kGruEpiWarps;
        const int half = ((warp - 2) >> 2) & 1;                  // which interleaved half of the channel chunks
        const int q4 = warp & 3;                                 // TMEM lane quarter this warp may read
        const int row_in_tile = q4 * 32 + lane;
        stage_gru_weights(p.w_ih, p.w_hh, Wi, Wh, C, p.KP, p.Npad, eall, kEpiThreads);
        for (int i = eall; i < 4 * C; i += kEpiThreads) {
            const int c = i >> 2, g = i & 3;
            bias_s[i] = g == 0 ? p.b_ih[c] + p.b_hh[c] : g == 1 ? p.b_ih[C + c] + p.b_hh[C + c] : g == 2 ? p.b_ih[2 * C + c] : p.b_hh[2 * C + c];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&w_bar);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");   // bias_s visible to both epilogue groups
        const uint32_t xs_addr = smem_u32(Xs), bias_addr = smem_u32(bias_s);
        const int gtid = (warp - 2 - grp * kGruEpiWarps) * 32 + lane;     // 0..255 within the group
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % kGruStages, a = it & 1;
            if (a != grp) continue;
            const int64_t row0 = tile * kGruTileM;
            const int64_t m = row0 + row_in_tile;
            const bool row_ok = m < p.N;
            constexpr int kIdPer = (kGruTileM * CQ + kGruEpiWarps * 32 - 1) / (kGruEpiWarps * 32);
            const int64_t rows_left = p.N - row0;
            const int nq_tile = (int)(rows_left < kGruTileM ? rows_left : kGruTileM) * CQ;               // float4 count
            mbar_wait(&tfull_bar[a], (uint32_t)(it >> 1) & 1u);
            tc_fence_after_sync();
            const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(a * p.Npad);
            const uint32_t stage_addr = xs_addr + (uint32_t)(s * stage_bytes);
            const uint32_t hrow = stage_addr + (uint32_t)op_bytes;
            // this row's h (own channel chunks) out of the operand stage; after the group barrier the whole stage — both
            // operand tiles are dead once the MMAs have completed — is reused as the row-major output staging area
            float4 hv[kJ][2];
#pragma unroll
            for (int j = 0; j < kJ; ++j)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int cq = 2 * (2 * j + half) + u;
                    hv[j][u] = cq < CQ ? lds128g(hrow + panel_chunk_offset(row_in_tile, cq, kGruTileM)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGruEpiWarps * 32) : "memory");
            const uint32_t st_h = stage_addr, st_g = stage_addr + (uint32_t)(kGruTileM * C * 4);     // [128][C] h', [128][C] gh_n
            float* rzn = p.rzn + m * 3 * C;
#pragma unroll
            for (int j = 0; j < kJ; ++j) {
                const int ch = 2 * j + half;
                const bool live = ch < NCH;                      // warp-uniform
                float v[32];
                tmem_ld32(lane_base + (uint32_t)((live ? ch : 0) * 32), v);
                if (j == kJ - 1) {                               // last read of the accumulator
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[a]);
                }
                if (!live) continue;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int cq = 2 * ch + u;
                    if (cq >= CQ) continue;
                    const int c = cq << 2;
                    float4 r, z, nn, gn, hw;
#define GRU_LANE(k, i)                                                                   \
    {                                                                                    \
        const float4 b = lds128g(bias_addr + 16u * (uint32_t)(c + i));                   \
        gn.k = v[16 * u + 4 * i + 3] + b.w;                                              \
        r.k = fast_sigmoid(v[16 * u + 4 * i] + b.x);                                     \
        z.k = fast_sigmoid(v[16 * u + 4 * i + 1] + b.y);                                 \
        nn.k = fast_tanh((v[16 * u + 4 * i + 2] + b.z) + r.k * gn.k);                    \
        hw.k = (1.f - z.k) * nn.k + z.k * hv[j][u].k;                                    \
    }
                    GRU_LANE(x, 0) GRU_LANE(y, 1) GRU_LANE(z, 2) GRU_LANE(w, 3)
#undef GRU_LANE
                    const uint32_t so = (uint32_t)(row_in_tile * C + c) * 4u;
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_h + so), "f"(hw.x), "f"(hw.y), "f"(hw.z), "f"(hw.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_g + so), "f"(gn.x), "f"(gn.y), "f"(gn.z), "f"(gn.w) : "memory");
                    if (row_ok) {
                        *reinterpret_cast<float4*>(rzn + c) = r;
                        *reinterpret_cast<float4*>(rzn + C + c) = z;
                        *reinterpret_cast<float4*>(rzn + 2 * C + c) = nn;
                    }
                }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGruEpiWarps * 32) : "memory");
            // coalesced copy-out of the tile: h_new, x_out = act(h' + identity), gh_n — contiguous [rows][C] blocks in HBM
            {
                float* hn = p.h_new + row0 * C;
                float* xo = p.x_out + row0 * C;
                float* gn = p.gh + row0 * C;
                // all residual loads of this thread are requested before the first use (ncu: 22 % of the kernel's warp
                // samples sat on these loads when each iteration waited for its own DRAM round trip)
                float4 idv[kIdPer];
#pragma unroll
                for (int k = 0; k < kIdPer; ++k) {
                    const int i = gtid + k * kGruEpiWarps * 32;
                    idv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.identity && i < nq_tile) idv[k] = __ldg(reinterpret_cast<const float4*>(p.identity + row0 * C) + i);
                }
#pragma unroll
                for (int k = 0; k < kIdPer; ++k) {
                    const int i = gtid + k * kGruEpiWarps * 32;
                    if (i < nq_tile) {
                        const float4 hw = lds128g(st_h + 16u * (uint32_t)i), g4 = lds128g(st_g + 16u * (uint32_t)i);
                        float4 xv;
                        xv.x = act_fwd(hw.x + idv[k].x, p.act, p.act_param); xv.y = act_fwd(hw.y + idv[k].y, p.act, p.act_param);
                        xv.z = act_fwd(hw.z + idv[k].z, p.act, p.act_param); xv.w = act_fwd(hw.w + idv[k].w, p.act, p.act_param);
                        reinterpret_cast<float4*>(hn)[i] = hw;
                        reinterpret_cast<float4*>(xo)[i] = xv;
                        reinterpret_cast<float4*>(gn)[i] = g4;
                    }
                }
            }
            fence_proxy_async_smem();                            // generic-proxy accesses to the stage before TMA refills it
            asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGruEpiWarps * 32) : "memory");
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

size_t gru_smem_bytes(int C, int* KP_out, int* Npad_out) {
    const int KP = (C + 7) / 8 * 8, Npad = (4 * C + 15) / 16 * 16;
    const int panels = (KP + kPanelFeatures - 1) / kPanelFeatures;
    if (KP_out) *KP_out = KP;
    if (Npad_out) *Npad_out = Npad;
    return (size_t)kGruStages * 2 * panels * kGruPanelBytes + (size_t)2 * panels * Npad * kPanelRowBytes + 1024;
}

bool al16p(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace
}  // namespace glam

using namespace glam;

// 1 when glam_gru_fused_fwd can run for this shape in the current math mode (tf32), else 0: the caller then uses
// glam_gemm_ex x 2 + glam_gru_gates_fwd (exact-fp32 mode, wide C, unaligned rows).
extern "C" int glam_gru_fused_supported(int channels) {
    if (g_math_mode_get() == 0) return 0;
    if (channels < 4 || (channels & 3)) return 0;
    int KP, Npad;
    const size_t smem = gru_smem_bytes(channels, &KP, &Npad);
    return ((channels == 32 || channels == 36 || channels == 40 || channels == 44) && 2 * Npad + 32 <= 512 && smem <= 222 * 1024) ? 1 : 0;
}

extern "C" int glam_gru_fused_fwd(const float* m, int64_t ldm, const float* h, int64_t ldh, const float* identity,
                                  const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int64_t N,
                                  int channels, int act, float act_param, float* rzn, float* gh, float* h_new, float* x_out,
                                  void* stream_) {
    GLAM_REQUIRE(N >= 0 && channels > 0, "glam_gru_fused_fwd: bad shape");
    GLAM_REQUIRE(glam_gru_fused_supported(channels), "glam_gru_fused_fwd: unsupported (channels=%d, math mode %d); use the unfused calls",
                 channels, g_math_mode_get());
    if (N == 0) return 0;
    GLAM_REQUIRE(m && h && w_ih && w_hh && b_ih && b_hh && rzn && gh && h_new && x_out, "glam_gru_fused_fwd: null pointer");
    GLAM_REQUIRE(ldm >= channels && ldh >= channels && (ldm & 3) == 0 && (ldh & 3) == 0 && al16p(m) && al16p(h) && al16p(identity) &&
                 al16p(w_ih) && al16p(w_hh) && al16p(b_ih) && al16p(b_hh) && al16p(rzn) && al16p(gh) && al16p(h_new) && al16p(x_out),
                 "glam_gru_fused_fwd: rows must be 16-byte aligned");
    GLAM_REQUIRE(N < ((int64_t)1 << 31), "glam_gru_fused_fwd: too many rows");
    TcGruParams p;
    p.w_ih = w_ih; p.w_hh = w_hh; p.b_ih = b_ih; p.b_hh = b_hh; p.identity = identity;
    p.rzn = rzn; p.gh = gh; p.h_new = h_new; p.x_out = x_out; p.N = N; p.C = channels; p.act = act; p.act_param = act_param;
    const size_t smem = gru_smem_bytes(channels, &p.KP, &p.Npad);
    CUtensorMap tm, th;
    if (int rc = make_tmap_rows(&tm, m, N, channels, ldm, kGruTileM, (int)CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_tmap_rows(&th, h, N, channels, ldh, kGruTileM, (int)CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const int64_t ntiles = (N + kGruTileM - 1) / kGruTileM;
    const int64_t grid = ntiles < kNumSMs ? ntiles : kNumSMs;
    auto launch = [&](auto kernel) -> int {
        cudaError_t e = ensure_dyn_smem((const void*)kernel, (size_t)((int)smem));
        if (e != cudaSuccess) { set_error("glam_gru_fused_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        kernel<<<(unsigned)grid, kGruThreads, smem, (cudaStream_t)stream_>>>(tm, th, p);
        return 0;
    };
    int rc = 0;
    switch (channels) {
        case 32: rc = launch(tc_gru_fwd_kernel<8>); break;
        case 36: rc = launch(tc_gru_fwd_kernel<9>); break;
        case 40: rc = launch(tc_gru_fwd_kernel<10>); break;
        default: rc = launch(tc_gru_fwd_kernel<11>); break;
    }
    if (rc) return rc;
    GLAM_CHECK_LAUNCH();
    return 0;
}
