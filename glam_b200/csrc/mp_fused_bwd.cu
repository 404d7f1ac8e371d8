// Backward of the message stack as ONE kernel: `message_steps` MessageBlock steps (src_1gp/layer.py:252-267 looped by
// src_1gp/model.py:53-54) differentiated in reverse on the same graph-aligned tiles as the forward (mp_fused.cu).
//
// A tile of whole graphs is closed under message passing in BOTH directions (block-diagonal batch: every in-edge AND every
// out-edge of a tile row stays inside the tile), so the whole reverse chain of a step
//   (g_x', g_h') --gates'--> G_GI | G_GH --W_ih, W_hh--> g_m, g_h --CELU'--> g_pre --W_scale^T--> g_agg
//        --(edge backward: dots + softmax' + leaky' by destination, scatter by SOURCE)--> g_xpe --W_ext^T--> g_x
// runs out of shared / tensor memory, and the carried gradients (g_x, g_h) never leave the SM between steps.  Unfused this was
// 8 launches per step (gate backward, 4 projection GEMMs, 2 edge kernels + their partial reduction) with g_m, g_agg, g_logit
// and the carried gradients round-tripping through HBM.  What still leaves the SM — once, as tile-sized contiguous copies —
// is what the weight-gradient contractions read (G_GI, G_GH, G_PRE, G_XPE) and g_x0; the small parameter gradients of the
// edge phase (weight_edge, att_edge) are accumulated in registers across all tiles and steps of a CTA and leave as one
// partial per CTA (fixed order: bitwise reproducible).
//
// Per 128-row tile and step (512 threads; warp 15 lane 0 issues the MMAs; CTA barriers between phases):
//   T0  gate backward, thread = (row, 4-channel chunk): saved r|z|n, gh_n, h, x' from HBM, the carried gradients from TENSOR
//       MEMORY (the previous step's MMA accumulators + two parked tiles) -> G = [g_r | g_z | g_n | g_n r] operand panels
//   T1  tcgen05.mma  g_m = G_GI W_ih, g_h = G_GH W_hh (one operand image, two K-slice selections); G_GI / G_GH copy-out
//   T2  epilogue: g_pre = g_m * CELU'(m) -> operand panels; the xp tile and alpha rows of the step arrive by cp.async
//   T3  tcgen05.mma  g_agg = g_pre W_scale^T; G_PRE copy-out
//   T4  TMEM -> g_agg tile (row-major)
//   T5  destination pass, a warp per row, lane = 16-byte chunk: per in-edge the dot <g_agg_i, w_e * xp_j> per head (segmented
//       shuffles), softmax' and leaky' -> g_logit, g_s_i; weight_edge / att_edge gradients in registers
//   T6  source pass, a warp per row: g_xp_j = sum over out-edges alpha w_e * g_agg_i, g_s_j -> g_xpe operand panels
//   T7  tcgen05.mma  g_x = g_xpe W_ext^T (consumed by the next step's T0 straight from tensor memory); G_XPE copy-out
// W_ih / W_hh as MMA operands (48 KB) do not fit next to the edge phase's tiles: they live in the arena during T0-T1 and are
// brought back from a pre-swizzled global image by the copy engine (cp.async.bulk) while T7 / T0 run.
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "mp_common.cuh"

namespace glam {

using namespace tc;
using namespace mp;

int g_math_mode_get();
extern unsigned long long* g_mp_phase_clock;

namespace {

constexpr int kBwMaxSteps = 8;
constexpr int kBwThreads = 512;

template <int CQ, int H>
struct BwGeom {
    static constexpr int NT = kBwThreads;
    static constexpr int C = 4 * CQ, HC = H * C, NQ = HC / 4;
    static constexpr int LD = (HC + 2 * H + 3) / 4 * 4, LQ = LD / 4;
    // G image (chunks of 4 fp32): [g_r CQ | g_z CQ | g_n CQ | pad to even] [g_n*r CQ | pad to even]
    static constexpr int KI = (3 * CQ + 1) / 2 * 2, KN = (CQ + 1) / 2 * 2, NR0 = KI, GCH = KI + KN, GPAN = (GCH + 7) / 8;
    static constexpr int KHC = 2 * CQ + KN;                              // K chunks of the W_hh operand: [r z | n]
    static constexpr int IPAN = (KI + 7) / 8, HPAN = (KHC + 7) / 8;
    static constexpr int KP = KN, PPAN = (KP + 7) / 8;                   // g_pre operand
    static constexpr int KX = (LQ + 1) / 2 * 2, XPAN = (KX + 7) / 8;     // g_xpe operand
    static constexpr int NS = (C + 15) / 16 * 16, NA = (HC + 15) / 16 * 16;
    static constexpr int TM_M = 0, TM_H = NS, TM_AGG = 2 * NS, TM_X = 2 * NS + NA, TM_GID = TM_X + NS, TM_GHZ = TM_GID + C,
                         TM_COLS = TM_GHZ + C;
    static constexpr int WQ = NT / 128, JPW = (CQ + WQ - 1) / WQ, NW = NT / 32, RPW = kMpM / NW;
    // arena (bytes): regions alias by phase, see the header comment
    static constexpr int BW_I_BYTES = IPAN * NS * 128, BW_H_BYTES = HPAN * NS * 128, BW_BYTES = BW_I_BYTES + BW_H_BYTES;
    static constexpr int A_G = 0, A_BW = GPAN * kMpPanel, A_PRE = A_BW;
    static constexpr int A_XP = 0, A_GX = 0;
    static constexpr int XP_BYTES = kMpM * LD * 4, GX_BYTES = XPAN * kMpPanel;
    static constexpr int A_AGG = ((XP_BYTES > GX_BYTES ? XP_BYTES : GX_BYTES) + 1023) / 1024 * 1024;
    static constexpr int A_ALPHA = A_AGG + kMpM * HC * 4, A_GL = A_ALPHA + kMpMaxEdges * H * 4;
    static constexpr int E0_ = A_GL + kMpMaxEdges * H * 4, E1_ = A_BW + BW_BYTES, E2_ = A_PRE + PPAN * kMpPanel;
    static constexpr int ARENA = ((E0_ > E1_ ? (E0_ > E2_ ? E0_ : E2_) : (E1_ > E2_ ? E1_ : E2_)) + 1023) / 1024 * 1024;
    static constexpr int SPAN = (KP + 7) / 8;
    static constexpr int OFF_WS = ARENA, OFF_WX = OFF_WS + SPAN * NA * 128, OFF_MISC = OFF_WX + XPAN * NS * 128;
    // misc (32-bit words)
    static constexpr int M_BAR = 0, M_RP = 8, M_RPS = M_RP + 132, M_REC = M_RPS + 132, M_RECS = M_REC + kMpMaxEdges,
                         M_GSI = M_RECS + kMpMaxEdges, M_WE = M_GSI + kMpM * 4, M_AE = M_WE + kMpMaxDe * HC,
                         M_AEW = (M_AE + kMpMaxDe * H + 3) / 4 * 4, M_CLK = M_AEW + NW * kMpMaxDe * H, M_END = M_CLK + 16;
    static constexpr int SMEM = OFF_MISC + M_END * 4;
    static_assert(H == 3, "the logit columns are laid out for three heads");
    static_assert(TM_COLS <= 512, "TMEM columns");
    static_assert(NQ + H <= 32 && KX / 2 * 2 <= 32, "a lane per 16-byte chunk of a row");
    static_assert(A_PRE + PPAN * kMpPanel <= A_ALPHA, "g_pre panels must not reach the alpha rows that arrive meanwhile");
    static_assert(A_BW >= XP_BYTES, "the xp tile arrives while the g_pre panels are live");
    static_assert(RPW * NW == kMpM, "rows per warp");
    static_assert(SMEM <= 232448, "shared memory");
};

struct BwParams {
    const float* X; const float* HH; const float* XPE; const float* ALPHA; const float* M; const float* RZN; const float* GH;
    const float* GT;                                  // tile-blocked gate save of the forward (replaces RZN, GH, HH, X', M here), or NULL
    const float* g_ext[kBwMaxSteps]; const float* g_h_final;
    const float* w_ext; int ldw; const float* w_edge; const float* att_edge; const float* w_scale;
    const uint8_t* bw_image;
    const int4* tiles; const int32_t* meta;
    const int32_t* rowptr; const int32_t* src; const uint8_t* etype;
    const int32_t* src_rowptr; const int32_t* src_pos; const int32_t* src_dst;
    int64_t N, E;
    int De, steps, act, res;
    float slope, act_param;
    float* G_GI; float* G_GH; float* G_PRE; float* G_XPE; float* g_x0;
    float* G4;                                        // [steps][N][4C] = g_r | g_z | g_n | g_n r (replaces G_GI / G_GH), or NULL
    float* g_h0;                                      // separate gradient of the initial GRU state (h0 was its own tensor), or NULL
    int pre_act; float pre_act_param;                 // activation of the input LinearBlock the forward applied (x0 = act(..)): g_x0 *= act'(x0)
    float* partial;                                   // [grid][De*HC + De*H]
    unsigned long long* phase_clock;                  // profiling aid (glam_message_stack_phase_clock), or NULL
};

// W_ih / W_hh [3C][C] -> the two B-operand images (rows = output channel n < C padded to NS, K-major SWIZZLE_128B panels):
//   BW_I[n][k] = W_ih[k][n], k < 3C;   BW_H[n][k'] = W_hh[k'][n] for k' < 2C (r, z), W_hh[2C + (k' - 2C)][n] for the n block
template <int CQ, int H>
__global__ void __launch_bounds__(256)
bw_image_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, uint8_t* __restrict__ image) {
    using G = BwGeom<CQ, H>;
    constexpr int C = G::C;
    const int total_i = G::IPAN * 8 * G::NS, total_h = G::HPAN * 8 * G::NS;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_i + total_h; i += gridDim.x * blockDim.x) {
        const bool is_h = i >= total_i;
        const int ii = is_h ? i - total_i : i;
        const int q = ii / G::NS, n = ii - q * G::NS;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < C) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!is_h) {
                    const int k = 4 * q + e;
                    if (k < 3 * C) v[e] = w_ih[(size_t)k * C + n];
                } else if (q < 2 * CQ) {
                    v[e] = w_hh[(size_t)(4 * q + e) * C + n];
                } else {
                    const int idx = 4 * (q - 2 * CQ) + e;
                    if (idx < C) v[e] = w_hh[(size_t)(2 * C + idx) * C + n];
                }
            }
        }
        uint8_t* base = image + (is_h ? G::BW_I_BYTES : 0) + (q >> 3) * (G::NS * 128) + pan_off(n, q);
        *reinterpret_cast<float4*>(base) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 tmem_ld4v(uint32_t taddr) {
    float v[4];
    tmem_ld4(taddr, v);
    return make_float4(v[0], v[1], v[2], v[3]);
}
struct GateIn { float4 r, z, n, ghn, hv, xo, gx, ghc; };
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int CQ, int H>
__global__ void __launch_bounds__(kBwThreads, 1)
mp_fused_bwd_kernel(const BwParams p) {
    using G = BwGeom<CQ, H>;
    constexpr int NT = G::NT, C = G::C, HC = G::HC, NQ = G::NQ, LD = G::LD, LQ = G::LQ, WQ = G::WQ;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* GP = sm + G::A_G;    uint8_t* BW = sm + G::A_BW;  uint8_t* PRE = sm + G::A_PRE;  uint8_t* GX = sm + G::A_GX;
    float* xp = reinterpret_cast<float*>(sm + G::A_XP);
    float* gagg = reinterpret_cast<float*>(sm + G::A_AGG);
    float* alpha_s = reinterpret_cast<float*>(sm + G::A_ALPHA);
    float* gl = reinterpret_cast<float*>(sm + G::A_GL);
    uint8_t* WS = sm + G::OFF_WS;  uint8_t* WX = sm + G::OFF_WX;
    float* misc = reinterpret_cast<float*>(sm + G::OFF_MISC);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(misc + G::M_BAR);
    uint64_t* bw_bar = mma_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + G::M_BAR + 4);
    int* rp = reinterpret_cast<int*>(misc + G::M_RP);   int* rps = reinterpret_cast<int*>(misc + G::M_RPS);
    int* rec = reinterpret_cast<int*>(misc + G::M_REC); int* recs = reinterpret_cast<int*>(misc + G::M_RECS);
    float* gsi = misc + G::M_GSI;  float* We = misc + G::M_WE;  float* Ae = misc + G::M_AE;
    float* aew = misc + G::M_AEW;                             // [warp][type][head] att_edge gradient accumulators
    unsigned int* clk = reinterpret_cast<unsigned int*>(misc + G::M_CLK);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool mma_thread = tid == NT - 32;
    const int ntiles = p.meta[0], flags = p.meta[1];
    const int PW = p.De * HC + p.De * H;                     // floats of one CTA's parameter-gradient partial

    if (flags != 0) {                                        // precondition violated (see mp_fused.cu): poison, the host checks meta[1]
        const int64_t nc = p.N * C;
        for (int64_t i = blockIdx.x * (int64_t)NT + tid; i < nc; i += (int64_t)gridDim.x * NT) p.g_x0[i] = CUDART_NAN_F;
        for (int i = tid; i < PW; i += NT) p.partial[(size_t)blockIdx.x * PW + i] = CUDART_NAN_F;
        return;
    }
    if ((int)blockIdx.x >= ntiles) {
        for (int i = tid; i < PW; i += NT) p.partial[(size_t)blockIdx.x * PW + i] = 0.f;
        return;
    }
    if (smem_u32(sm) & 1023u) __trap();

    // ---------------------------------------------------------------- set-up: barriers, TMEM, resident weights
    if (tid == 0) { mbar_init(mma_bar, 1); mbar_init(bw_bar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < (G::OFF_MISC - G::OFF_WS) / 16; i += NT) sts128(WS + 16 * i, z4);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    {
        // W_scale^T operand: B[n][k] = w_scale[n][k] (n < HC, k < C): rows of w_scale as they are stored
        for (int i = tid; i < HC * CQ; i += NT) {
            const int n = i / CQ, q = i - n * CQ;
            sts128(WS + (q >> 3) * (G::NA * 128) + pan_off(n, q), ldg4(p.w_scale + (size_t)n * C + 4 * q));
        }
        // W_ext^T operand: B[n][k] = w_ext[n][k] (n < C, k < LD)
        for (int i = tid; i < C * LQ; i += NT) {
            const int n = i / LQ, q = i - n * LQ;
            sts128(WX + (q >> 3) * (G::NS * 128) + pan_off(n, q), ldg4(p.w_ext + (size_t)n * p.ldw + 4 * q));
        }
        for (int i = tid; i < p.De * HC; i += NT) We[i] = p.w_edge[i];
        for (int i = tid; i < p.De * H; i += NT) Ae[i] = p.att_edge[i];
    }
    fence_proxy_async_smem();
    __syncthreads();
    uint32_t bw_ph = 0;
    if (mma_thread) {                                        // first copy of the W_ih | W_hh operand image
        mbar_expect_tx(bw_bar, (uint32_t)G::BW_BYTES);
        bulk_g2s(BW, p.bw_image, (uint32_t)G::BW_BYTES, bw_bar);
    }

    const uint64_t d_g = make_smem_desc(smem_u32(GP), 16, 1024), d_bwi = make_smem_desc(smem_u32(BW), 16, 1024);
    const uint64_t d_bwh = make_smem_desc(smem_u32(BW + G::BW_I_BYTES), 16, 1024), d_pre = make_smem_desc(smem_u32(PRE), 16, 1024);
    const uint64_t d_ws = make_smem_desc(smem_u32(WS), 16, 1024), d_gx = make_smem_desc(smem_u32(GX), 16, 1024);
    const uint64_t d_wx = make_smem_desc(smem_u32(WX), 16, 1024);
    // descriptor of the k-step that starts at 16-byte chunk q of an image whose panels are `pan` bytes apart
    auto dsc = [](uint64_t base, int q, int pan) { return base + (uint64_t)(((q >> 3) * pan + (q & 7) * 16) >> 4); };
    const int q4 = warp & 3, cg = warp >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    uint32_t ph = 0;
    const int64_t NC = p.N * C;
    const int S = p.steps;
    // edge-phase roles of this lane: 16-byte chunk lq of a row (head hq); lanes NQ..NQ+H-1 collect g_s_j in the source pass
    const int lq = lane < NQ ? lane : 0, hq = lq / CQ, hidx = lq - hq * CQ;
    const bool head_leader = lane < NQ && hidx == 0;
    float4 wacc[kMpMaxDe];
#pragma unroll
    for (int ty = 0; ty < kMpMaxDe; ++ty) wacc[ty] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < G::NW * kMpMaxDe * H; i += NT) aew[i] = 0.f;                 // (the first tile's barrier orders this)

    unsigned int clk_last = (unsigned int)clock64();
    if (p.phase_clock && tid < 16) clk[tid] = 0;
    __syncthreads();
#define BW_TICK(I)                                                              \
    if (p.phase_clock && tid == 0) {                                            \
        const unsigned int now_ = (unsigned int)clock64();                      \
        clk[I] += now_ - clk_last;                                              \
        clk_last = now_;                                                        \
    }
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int4 td = p.tiles[t];
        const int4 tdn = t + (int)gridDim.x < ntiles ? p.tiles[t + gridDim.x] : td;
        const int n0 = td.x, nd = td.y - td.x, e0 = td.z, ne = td.w - td.z;
        // ------------------------------------------------------------ tile index words (both directions)
        const int ks0 = p.src_rowptr[n0];
        for (int i = tid; i <= kMpM; i += NT) {
            rp[i] = i <= nd ? p.rowptr[n0 + i] - e0 : ne;
            rps[i] = i <= nd ? p.src_rowptr[n0 + i] - ks0 : ne;
        }
        for (int e = tid; e < ne; e += NT) {
            rec[e] = (p.src[e0 + e] - n0) | ((int)p.etype[e0 + e] << 8);
            const int pos = p.src_pos[ks0 + e] - e0;
            recs[e] = pos | ((p.src_dst[ks0 + e] - n0) << 10) | ((int)p.etype[e0 + pos] << 18);
        }
        __syncthreads();
        BW_TICK(0)

        for (int s = S - 1; s >= 0; --s) {
            const bool first = s == S - 1;                   // the first step processed: carried gradients come from outside
            // -------------------------------------------------------- T0: gate backward -> G operand panels, parked g_id / g_h z
            // the saved activations of a 4-channel chunk travel one round ahead of the gate arithmetic (and the first round is
            // requested before the wait on the previous step's g_x MMA)
            auto gate_load = [&](int jj) {
                GateIn in;
                const float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f);
                in.r = in.z = in.n = in.ghn = in.hv = in.xo = in.gx = in.ghc = z0;
                const int j = cg + WQ * jj;
                if (j < CQ && row < nd) {
                    if (p.GT) {                              // 32 rows of one chunk = 512 contiguous bytes
                        const float4* gt = reinterpret_cast<const float4*>(p.GT + ((size_t)s * p.N + n0) * 7 * C);
                        in.r = __ldg(gt + j * nd + row); in.z = __ldg(gt + (CQ + j) * nd + row); in.n = __ldg(gt + (2 * CQ + j) * nd + row);
                        in.ghn = __ldg(gt + (3 * CQ + j) * nd + row); in.hv = __ldg(gt + (4 * CQ + j) * nd + row);
                        in.xo = __ldg(gt + (5 * CQ + j) * nd + row);
                    } else {
                        const size_t o = ((size_t)s * p.N + n0 + row) * C + 4 * j, o3 = ((size_t)s * p.N + n0 + row) * 3 * C + 4 * j;
                        in.r = ldg4(p.RZN + o3); in.z = ldg4(p.RZN + o3 + C); in.n = ldg4(p.RZN + o3 + 2 * C);
                        in.ghn = ldg4(p.GH + o); in.hv = ldg4(p.HH + o); in.xo = ldg4(p.X + o + NC);
                    }
                    if (p.g_ext[s]) in.gx = ldg4(p.g_ext[s] + (size_t)(n0 + row) * C + 4 * j);
                    if (first && p.g_h_final) in.ghc = ldg4(p.g_h_final + (size_t)(n0 + row) * C + 4 * j);
                }
                return in;
            };
            GateIn nxt = gate_load(0);
            if (!first) { mbar_wait_guarded(mma_bar, ph); ph ^= 1u; tc_fence_after_sync(); }        // the previous step's g_x MMA
            BW_TICK(10)
#pragma unroll
            for (int jj = 0; jj < G::JPW; ++jj) {
                const int j = cg + WQ * jj;
                const GateIn cur = nxt;
                if (jj + 1 < G::JPW) nxt = gate_load(jj + 1);
                if (j < CQ) {                                // warp-uniform
                    const float4 r = cur.r, z = cur.z, nn = cur.n, ghn = cur.ghn, hv = cur.hv, xo = cur.xo;
                    float4 gx = cur.gx, ghc = cur.ghc;
                    if (!first) {
                        gx = add4(gx, tmem_ld4v(lane_base + G::TM_X + 4 * j));
                        if (p.res) gx = add4(gx, tmem_ld4v(lane_base + G::TM_GID + 4 * j));
                        ghc = add4(tmem_ld4v(lane_base + G::TM_H + 4 * j), tmem_ld4v(lane_base + G::TM_GHZ + 4 * j));
                    }
                    float4 gs, grp, gzp, gnp, gnr, ghz;
#define BW_GATE(k)                                                                              \
    {                                                                                           \
        gs.k = row < nd ? gx.k * act_grad_from_out(xo.k, p.act, p.act_param) : 0.f;             \
        const float ghp = gs.k + ghc.k;                                                         \
        gnp.k = ghp * (1.f - z.k) * (1.f - nn.k * nn.k);                                        \
        gzp.k = ghp * (hv.k - nn.k) * z.k * (1.f - z.k);                                        \
        grp.k = gnp.k * ghn.k * r.k * (1.f - r.k);                                              \
        gnr.k = gnp.k * r.k;                                                                    \
        ghz.k = ghp * z.k;                                                                      \
    }
                    BW_GATE(x) BW_GATE(y) BW_GATE(z) BW_GATE(w)
#undef BW_GATE
                    sts128(GP + (j >> 3) * kMpPanel + pan_off(row, j), grp);
                    sts128(GP + ((CQ + j) >> 3) * kMpPanel + pan_off(row, CQ + j), gzp);
                    sts128(GP + ((2 * CQ + j) >> 3) * kMpPanel + pan_off(row, 2 * CQ + j), gnp);
                    sts128(GP + ((G::NR0 + j) >> 3) * kMpPanel + pan_off(row, G::NR0 + j), gnr);
                    tmem_st4(lane_base + G::TM_GID + 4 * j, gs);
                    tmem_st4(lane_base + G::TM_GHZ + 4 * j, ghz);
                }
            }
            if (cg == 0) {                                   // K padding chunks of the image
                const float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q = 3 * CQ; q < G::KI; ++q) sts128(GP + (q >> 3) * kMpPanel + pan_off(row, q), z0);
                for (int q = G::NR0 + CQ; q < G::GCH; ++q) sts128(GP + (q >> 3) * kMpPanel + pan_off(row, q), z0);
            }
            BW_TICK(11)
            tmem_st_wait();
            tc_fence_before_sync();
            fence_proxy_async_smem();
            __syncthreads();
            BW_TICK(1)
            // -------------------------------------------------------- T1: g_m = G_GI W_ih, g_h = G_GH W_hh; G_GI / G_GH copy-out
            if (mma_thread) {
                mbar_wait_guarded(bw_bar, bw_ph);            // the W_ih | W_hh image has landed
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, G::NS, 0, 0);
#pragma unroll
                for (int i = 0; i < G::KI / 2; ++i)
                    mma_tf32_ss(tmem_base + G::TM_M, dsc(d_g, 2 * i, kMpPanel), dsc(d_bwi, 2 * i, G::NS * 128), idesc, i > 0 ? 1u : 0u);
#pragma unroll
                for (int i = 0; i < CQ; ++i)
                    mma_tf32_ss(tmem_base + G::TM_H, dsc(d_g, 2 * i, kMpPanel), dsc(d_bwh, 2 * i, G::NS * 128), idesc, i > 0 ? 1u : 0u);
#pragma unroll
                for (int i = 0; i < G::KN / 2; ++i)
                    mma_tf32_ss(tmem_base + G::TM_H, dsc(d_g, G::NR0 + 2 * i, kMpPanel), dsc(d_bwh, 2 * CQ + 2 * i, G::NS * 128), idesc, 1u);
                mma_commit(mma_bar);
            }
            bw_ph ^= 1u;
            if (p.G4) {                                     // one [4C] row per node: both GRU weight gradients read it
                float4* g4 = reinterpret_cast<float4*>(p.G4 + ((size_t)s * p.N + n0) * 4 * C);
                for (int i = tid; i < nd * 4 * CQ; i += NT) {
                    const int r = i / (4 * CQ), q = i - r * (4 * CQ), qi = q < 3 * CQ ? q : G::NR0 + q - 3 * CQ;
                    g4[i] = lds128(GP + (qi >> 3) * kMpPanel + pan_off(r, qi));
                }
            } else {
                float4* gi = reinterpret_cast<float4*>(p.G_GI + ((size_t)s * p.N + n0) * 3 * C);
                float4* gh = reinterpret_cast<float4*>(p.G_GH + ((size_t)s * p.N + n0) * 3 * C);
                for (int i = tid; i < nd * 3 * CQ; i += NT) {
                    const int r = i / (3 * CQ), q = i - r * (3 * CQ), qh = q < 2 * CQ ? q : G::NR0 + q - 2 * CQ;
                    gi[i] = lds128(GP + (q >> 3) * kMpPanel + pan_off(r, q));
                    gh[i] = lds128(GP + (qh >> 3) * kMpPanel + pan_off(r, qh));
                }
            }
            float4 m_keep[G::JPW];                           // the CELU outputs of the step (for CELU'), requested while the MMAs run
#pragma unroll
            for (int jj = 0; jj < G::JPW; ++jj) {
                const int j = cg + WQ * jj;
                m_keep[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < CQ && row < nd)
                    m_keep[jj] = p.GT ? __ldg(reinterpret_cast<const float4*>(p.GT + ((size_t)s * p.N + n0) * 7 * C) + (6 * CQ + j) * nd + row)
                                      : ldg4(p.M + ((size_t)s * p.N + n0 + row) * C + 4 * j);
            }
            mbar_wait_guarded(mma_bar, ph); ph ^= 1u;
            tc_fence_after_sync();
            __syncthreads();                                 // copy-outs done, MMA done: the G image and the weight image are dead
            BW_TICK(2)
            // -------------------------------------------------------- T2: the step's xp tile and alpha rows on their way (cp.async);
            // epilogue g_pre = g_m * CELU'(m) -> operand panels
            {
                // what the NEXT step (or the next tile's first step) reads, on its way into L2 meanwhile
                const bool more = s > 0;
                const int4 tp = more ? td : tdn;
                if (more || t + (int)gridDim.x < ntiles) {
                    const int sp = more ? s - 1 : S - 1, pn0 = tp.x, pnd = tp.y - tp.x;
                    const size_t ro = (size_t)sp * p.N + pn0;
                    auto pf = [&](const float* base, int bytes) {
                        const char* b = reinterpret_cast<const char*>(base);
                        for (int o = tid * 128; o < bytes; o += NT * 128) prefetch_l2(b + o);
                    };
                    if (p.GT) {
                        pf(p.GT + ro * 7 * C, pnd * 7 * C * 4);
                    } else {
                        pf(p.RZN + ro * 3 * C, pnd * 3 * C * 4); pf(p.GH + ro * C, pnd * C * 4); pf(p.HH + ro * C, pnd * C * 4);
                        pf(p.X + (ro + p.N) * C, pnd * C * 4); pf(p.M + ro * C, pnd * C * 4);
                    }
                    pf(p.XPE + ro * LD, pnd * LD * 4);
                    pf(p.ALPHA + ((size_t)sp * p.E + tp.z) * H, (tp.w - tp.z) * H * 4);
                    if (p.g_ext[sp]) pf(p.g_ext[sp] + (size_t)pn0 * C, pnd * C * 4);
                }
                const float* xs = p.XPE + ((size_t)s * p.N + n0) * LD;
                for (int i = tid; i < nd * LQ; i += NT) cp_async16(xp + 4 * i, xs + 4 * i);
                const float* as = p.ALPHA + ((size_t)s * p.E + e0) * H;
                for (int i = tid; i < ne * H; i += NT) cp_async4(alpha_s + i, as + i);
            }
#pragma unroll
            for (int jj = 0; jj < G::JPW; ++jj) {
                const int j = cg + WQ * jj;
                if (j < CQ) {
                    const float4 gm = tmem_ld4v(lane_base + G::TM_M + 4 * j), m = m_keep[jj];
                    const float4 gp = make_float4(gm.x * (m.x > 0.f ? 1.f : m.x + 1.f), gm.y * (m.y > 0.f ? 1.f : m.y + 1.f),
                                                  gm.z * (m.z > 0.f ? 1.f : m.z + 1.f), gm.w * (m.w > 0.f ? 1.f : m.w + 1.f));
                    sts128(PRE + (j >> 3) * kMpPanel + pan_off(row, j), gp);
                }
            }
            if (cg == 0)
                for (int q = CQ; q < G::KP; ++q) sts128(PRE + (q >> 3) * kMpPanel + pan_off(row, q), make_float4(0.f, 0.f, 0.f, 0.f));
            tc_fence_before_sync();
            fence_proxy_async_smem();
            __syncthreads();
            BW_TICK(3)
            // -------------------------------------------------------- T3: g_agg = g_pre W_scale^T; G_PRE copy-out
            if (mma_thread) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, G::NA, 0, 0);
#pragma unroll
                for (int i = 0; i < G::KP / 2; ++i)
                    mma_tf32_ss(tmem_base + G::TM_AGG, dsc(d_pre, 2 * i, kMpPanel), dsc(d_ws, 2 * i, G::NA * 128), idesc, i > 0 ? 1u : 0u);
                mma_commit(mma_bar);
            }
            {
                float4* gp = reinterpret_cast<float4*>(p.G_PRE + ((size_t)s * p.N + n0) * C);
                for (int i = tid; i < nd * CQ; i += NT) {
                    const int r = i / CQ, q = i - r * CQ;
                    gp[i] = lds128(PRE + (q >> 3) * kMpPanel + pan_off(r, q));
                }
            }
            mbar_wait_guarded(mma_bar, ph); ph ^= 1u;
            tc_fence_after_sync();
            __syncthreads();                                 // G_PRE copy-out done before the g_agg tile overwrites the panels
            BW_TICK(4)
            // -------------------------------------------------------- T4: TMEM -> g_agg tile (row-major, pitch HC)
            for (int c0 = 16 * cg; c0 < HC; c0 += 16 * WQ) {
                float v[16];
                tmem_ld16(lane_base + G::TM_AGG + c0, v);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c0 + 4 * i < HC) sts128(gagg + row * HC + c0 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
            cp_async_wait_all();
            tc_fence_before_sync();
            __syncthreads();
            BW_TICK(5)
            // -------------------------------------------------------- T5: destination pass (src_1gp/layer.py:46-54 differentiated)
            // A row with in-degree SL <= 4 (valence-bounded molecular graphs) takes a routine specialised for exactly SL edges: their
            // loads are issued together and their per-head shuffle reductions overlap; the bond type selects the code path (it is
            // uniform over the warp), so weight_edge rows and their gradient accumulators are addressed statically.  Hubs loop.
            float4 wt[kMpMaxDe];
#pragma unroll
            for (int ty = 0; ty < kMpMaxDe; ++ty) wt[ty] = ty < p.De ? lds128(We + ty * HC + 4 * lq) : make_float4(0.f, 0.f, 0.f, 0.f);
            const bool hs8 = hidx + 8 < CQ, hs4 = hidx < 4, hs2 = hidx < 2, hs1 = hidx < 1;
            auto head_sum = [&](float v) {                   // fixed-order sum over the CQ lanes of a head; valid on the head's first lane
                float u;
                if (CQ > 8) { u = __shfl_down_sync(0xffffffffu, v, 8); v += hs8 ? u : 0.f; }
                u = __shfl_down_sync(0xffffffffu, v, 4); v += hs4 ? u : 0.f;
                u = __shfl_down_sync(0xffffffffu, v, 2); v += hs2 ? u : 0.f;
                u = __shfl_down_sync(0xffffffffu, v, 1); v += hs1 ? u : 0.f;
                return v;
            };
            // lane partial of g_alpha = <g_agg_i, w_e * xp_j> (4 channels) + this edge's weight_edge gradient alpha * g_agg_i * xp_j
#define BW_EDGE_CASE(K)                                                                                                         \
    case K:                                                                                                                     \
        wacc[K].x = fmaf(al_, gx_.x, wacc[K].x); wacc[K].y = fmaf(al_, gx_.y, wacc[K].y);                                        \
        wacc[K].z = fmaf(al_, gx_.z, wacc[K].z); wacc[K].w = fmaf(al_, gx_.w, wacc[K].w);                                        \
        return fmaf(gx_.x, wt[K].x, fmaf(gx_.y, wt[K].y, fmaf(gx_.z, wt[K].z, gx_.w * wt[K].w)));
            auto edge_dot = [&](int ty_, float al_, const float4& ga_, const float4& xj_) -> float {
                const float4 gx_ = make_float4(ga_.x * xj_.x, ga_.y * xj_.y, ga_.z * xj_.z, ga_.w * xj_.w);
                switch (ty_) {
                    BW_EDGE_CASE(0)
                    BW_EDGE_CASE(1)
                    BW_EDGE_CASE(2)
                    default:
                    BW_EDGE_CASE(3)
                }
            };
#undef BW_EDGE_CASE
            float* ae_w = aew + warp * (kMpMaxDe * H);       // this warp's att_edge gradient accumulators [type][head]
            auto dst_row = [&](auto slots_tag, int d, int beg) {
                constexpr int SL = decltype(slots_tag)::value;
                const float4 ga = lds128(gagg + d * HC + 4 * lq);
                const float si = xp[d * LD + HC + hq];
                int ty[SL]; float al[SL], v[SL], lg[SL];
                float4 xj[SL];
#pragma unroll
                for (int i = 0; i < SL; ++i) {
                    const int rc = rec[beg + i];
                    ty[i] = rc >> 8;
                    al[i] = alpha_s[(beg + i) * H + hq];
                    xj[i] = lds128(xp + (rc & 0xff) * LD + 4 * lq);
                    lg[i] = si + Ae[ty[i] * H + hq] + xp[(rc & 0xff) * LD + HC + H + hq];
                }
#pragma unroll
                for (int i = 0; i < SL; ++i) v[i] = edge_dot(ty[i], al[i], ga, xj[i]);
#pragma unroll
                for (int i = 0; i < SL; ++i) v[i] = head_sum(v[i]);                   // g_alpha of edge i, head hq (on the head's first lane)
                if (head_leader) {
                    float dsum = 0.f, gsi_acc = 0.f;
#pragma unroll
                    for (int i = 0; i < SL; ++i) dsum = fmaf(al[i], v[i], dsum);
#pragma unroll
                    for (int i = 0; i < SL; ++i) {
                        float g = al[i] * (v[i] - dsum);                               // softmax backward
                        g = lg[i] > 0.f ? g : p.slope * g;                             // leaky_relu backward
                        gl[(beg + i) * H + hq] = g;
                        gsi_acc += g;
                        ae_w[ty[i] * H + hq] += g;
                    }
                    gsi[d * 4 + hq] = gsi_acc;
                }
            };
#pragma unroll 1
            for (int k = 0; k < G::RPW; ++k) {
                const int d = warp * G::RPW + k;
                const int beg = rp[d], end = rp[d + 1], deg = end - beg;
                switch (deg) {
                    case 0: if (head_leader) gsi[d * 4 + hq] = 0.f; continue;
                    case 1: dst_row(std::integral_constant<int, 1>{}, d, beg); continue;
                    case 2: dst_row(std::integral_constant<int, 2>{}, d, beg); continue;
                    case 3: dst_row(std::integral_constant<int, 3>{}, d, beg); continue;
                    case 4: dst_row(std::integral_constant<int, 4>{}, d, beg); continue;
                    default: break;
                }
                const float4 ga = lds128(gagg + d * HC + 4 * lq);
                const float si = xp[d * LD + HC + hq];
                float dsum = 0.f;
                for (int e = beg; e < end; ++e) {
                    const int rc = rec[e];
                    const float al = alpha_s[e * H + hq];
                    const float galpha = head_sum(edge_dot(rc >> 8, al, ga, lds128(xp + (rc & 0xff) * LD + 4 * lq)));
                    if (head_leader) { dsum = fmaf(al, galpha, dsum); gl[e * H + hq] = galpha; }
                }
                if (head_leader) {
                    float gsi_acc = 0.f;
                    for (int e = beg; e < end; ++e) {
                        const int rc = rec[e], ty = rc >> 8;
                        float g = alpha_s[e * H + hq] * (gl[e * H + hq] - dsum);       // softmax backward
                        const float l = si + Ae[ty * H + hq] + xp[(rc & 0xff) * LD + HC + H + hq];
                        g = l > 0.f ? g : p.slope * g;                                 // leaky_relu backward
                        gl[e * H + hq] = g;
                        gsi_acc += g;
                        ae_w[ty * H + hq] += g;
                    }
                    gsi[d * 4 + hq] = gsi_acc;
                }
            }
            __syncthreads();
            BW_TICK(6)
            // -------------------------------------------------------- T6: source pass -> g_xpe operand panels (they alias the xp tile)
            // (the weight_edge row of each out-edge comes from shared memory by bond type: no control flow in this pass)
            const int cidx = (lane >= NQ && lane < NQ + H) ? lane - NQ : 0;           // lanes NQ .. NQ+H-1 sum g_logit into g_s_j
            const float cmask = (lane >= NQ && lane < NQ + H) ? 1.f : 0.f;
            const float* we_l = We + 4 * lq;
            auto src_row = [&](auto slots_tag, int beg, float4& a, float& gsj) {
                constexpr int SL = decltype(slots_tag)::value;
                float al[SL], glv[SL]; float4 ga[SL], w[SL];
#pragma unroll
                for (int i = 0; i < SL; ++i) {
                    const int rs = recs[beg + i], pos = rs & 0x3ff;
                    al[i] = alpha_s[pos * H + hq];
                    glv[i] = gl[pos * H + cidx];
                    ga[i] = lds128(gagg + ((rs >> 10) & 0xff) * HC + 4 * lq);
                    w[i] = lds128(we_l + (rs >> 18) * HC);
                }
#pragma unroll
                for (int i = 0; i < SL; ++i) {
                    a.x = fmaf(al[i], w[i].x * ga[i].x, a.x); a.y = fmaf(al[i], w[i].y * ga[i].y, a.y);
                    a.z = fmaf(al[i], w[i].z * ga[i].z, a.z); a.w = fmaf(al[i], w[i].w * ga[i].w, a.w);
                    gsj = fmaf(cmask, glv[i], gsj);
                }
            };
#pragma unroll 1
            for (int k = 0; k < G::RPW; ++k) {
                const int j = warp * G::RPW + k;
                const int beg = rps[j], end = rps[j + 1];
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                float gsj = 0.f;
                const float gsi_l = gsi[j * 4 + (lane & 3)];                           // lanes 0..2 (mod 4): g_s_i of heads 0..2
                switch (end - beg) {
                    case 0: break;
                    case 1: src_row(std::integral_constant<int, 1>{}, beg, a, gsj); break;
                    case 2: src_row(std::integral_constant<int, 2>{}, beg, a, gsj); break;
                    case 3: src_row(std::integral_constant<int, 3>{}, beg, a, gsj); break;
                    case 4: src_row(std::integral_constant<int, 4>{}, beg, a, gsj); break;
                    default:
                        for (int e = beg; e < end; ++e) {
                            const int rs = recs[e], pos = rs & 0x3ff;
                            const float al = alpha_s[pos * H + hq];
                            const float4 ga = lds128(gagg + ((rs >> 10) & 0xff) * HC + 4 * lq), w = lds128(we_l + (rs >> 18) * HC);
                            a.x = fmaf(al, w.x * ga.x, a.x); a.y = fmaf(al, w.y * ga.y, a.y);
                            a.z = fmaf(al, w.z * ga.z, a.z); a.w = fmaf(al, w.w * ga.w, a.w);
                            gsj = fmaf(cmask, gl[pos * H + cidx], gsj);
                        }
                }
                // columns HC .. HC+3 = g_s_i[0..2] | g_s_j[0] (lane NQ); HC+4 .. HC+7 = g_s_j[1], g_s_j[2], padding (lane NQ+1)
                const float g0 = __shfl_sync(0xffffffffu, gsj, NQ), g1 = __shfl_sync(0xffffffffu, gsj, NQ + 1),
                            g2 = __shfl_sync(0xffffffffu, gsj, NQ + 2);
                const float s0 = __shfl_sync(0xffffffffu, gsi_l, 0), s1 = __shfl_sync(0xffffffffu, gsi_l, 1),
                            s2 = __shfl_sync(0xffffffffu, gsi_l, 2);
                if (lane == NQ) a = make_float4(s0, s1, s2, g0);
                else if (lane == NQ + 1) a = make_float4(g1, g2, 0.f, 0.f);
                else if (lane > NQ + 1) a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < G::KX) sts128(GX + (lane >> 3) * kMpPanel + pan_off(j, lane), a);
            }
            fence_proxy_async_smem();
            __syncthreads();
            BW_TICK(7)
            // -------------------------------------------------------- T7: g_x = g_xpe W_ext^T (read by the next T0 from tensor memory);
            // the weight image comes back for the next step; G_XPE copy-out
            if (mma_thread) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, G::NS, 0, 0);
#pragma unroll
                for (int i = 0; i < G::KX / 2; ++i)
                    mma_tf32_ss(tmem_base + G::TM_X, dsc(d_gx, 2 * i, kMpPanel), dsc(d_wx, 2 * i, G::NS * 128), idesc, i > 0 ? 1u : 0u);
                mma_commit(mma_bar);
                if (s > 0 || t + (int)gridDim.x < ntiles) {
                    mbar_expect_tx(bw_bar, (uint32_t)G::BW_BYTES);
                    bulk_g2s(BW, p.bw_image, (uint32_t)G::BW_BYTES, bw_bar);
                }
            }
            {
                float4* gx = reinterpret_cast<float4*>(p.G_XPE + ((size_t)s * p.N + n0) * LD);
                for (int i = tid; i < nd * LQ; i += NT) {
                    const int r = i / LQ, q = i - r * LQ;
                    gx[i] = lds128(GX + (q >> 3) * kMpPanel + pan_off(r, q));
                }
            }
            __syncthreads();                                 // copy-out done before the next step's T0 rewrites the panels
            BW_TICK(8)
        }
        // ------------------------------------------------------------ tile output: g_x0 = g_x + g_h (X[0] and HH[0] are both x0)
        mbar_wait_guarded(mma_bar, ph); ph ^= 1u;
        tc_fence_after_sync();
#pragma unroll
        for (int jj = 0; jj < G::JPW; ++jj) {
            const int j = cg + WQ * jj;
            if (j < CQ) {
                float4 g = tmem_ld4v(lane_base + G::TM_X + 4 * j);
                const float4 gh = add4(tmem_ld4v(lane_base + G::TM_H + 4 * j), tmem_ld4v(lane_base + G::TM_GHZ + 4 * j));
                if (p.res) g = add4(g, tmem_ld4v(lane_base + G::TM_GID + 4 * j));
                if (!p.g_h0) g = add4(g, gh);                // h0 == x0 (layer.py:253-254): one tensor, one gradient
                if (p.pre_act && row < nd) {                 // through the input LinearBlock's activation: what leaves is d/d(pre-activation)
                    const float4 x0v = *reinterpret_cast<const float4*>(p.X + (size_t)(n0 + row) * C + 4 * j);
                    g.x *= act_grad_from_out(x0v.x, p.pre_act, p.pre_act_param); g.y *= act_grad_from_out(x0v.y, p.pre_act, p.pre_act_param);
                    g.z *= act_grad_from_out(x0v.z, p.pre_act, p.pre_act_param); g.w *= act_grad_from_out(x0v.w, p.pre_act, p.pre_act_param);
                }
                if (row < nd) {
                    *reinterpret_cast<float4*>(p.g_x0 + (size_t)(n0 + row) * C + 4 * j) = g;
                    if (p.g_h0) *reinterpret_cast<float4*>(p.g_h0 + (size_t)(n0 + row) * C + 4 * j) = gh;
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();                                     // TMEM reads done before the next tile's MMAs; index words reusable
        BW_TICK(9)
    }

#undef BW_TICK
    if (p.phase_clock && tid < 16) p.phase_clock[blockIdx.x * 32 + 16 + tid] = clk[tid];
    // ---------------------------------------------------------------- parameter-gradient partial of this CTA (fixed order over warps)
    {
        float4* stage = reinterpret_cast<float4*>(sm);       // [NW][De][NQ] float4
        if (lane < NQ)
#pragma unroll
            for (int ty = 0; ty < kMpMaxDe; ++ty) stage[(warp * kMpMaxDe + ty) * NQ + lane] = wacc[ty];
        __syncthreads();
        float* out = p.partial + (size_t)blockIdx.x * PW;
        for (int i = tid; i < p.De * HC; i += NT) {
            const int ty = i / HC, c = i - ty * HC;
            float acc = 0.f;
            for (int w = 0; w < G::NW; ++w) acc += reinterpret_cast<const float*>(stage + (w * kMpMaxDe + ty) * NQ)[c];
            out[i] = acc;
        }
        for (int i = tid; i < p.De * H; i += NT) {
            const int ty = i / H, h = i - ty * H;
            float acc = 0.f;
            for (int w = 0; w < G::NW; ++w) acc += aew[(w * kMpMaxDe + ty) * H + h];
            out[p.De * HC + i] = acc;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

bool al16b(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int CQ, int H>
int bwd_launch(const BwParams& p, const float* w_ih, const float* w_hh, uint8_t* image, cudaStream_t stream) {
    using G = BwGeom<CQ, H>;
    bw_image_kernel<CQ, H><<<8, 256, 0, stream>>>(w_ih, w_hh, image);
    cudaError_t e = ensure_dyn_smem((const void*)mp_fused_bwd_kernel<CQ, H>, (size_t)(G::SMEM));
    if (e != cudaSuccess) { set_error("glam_message_stack_bwd: cudaFuncSetAttribute(%d bytes): %s", G::SMEM, cudaGetErrorString(e)); return (int)e; }
    mp_fused_bwd_kernel<CQ, H><<<kNumSMs, kBwThreads, G::SMEM, stream>>>(p);
    return 0;
}

template <int CQ, int H>
size_t bwd_image_bytes() { return (size_t)BwGeom<CQ, H>::BW_BYTES; }

}  // namespace
}  // namespace glam

using namespace glam;

extern "C" int glam_message_stack_bwd_supported(int channels, int heads, int edge_dim, int steps) {
    if (g_math_mode_get() == 0) return 0;
    if (heads != 3 || edge_dim < 1 || edge_dim > kMpMaxDe || steps < 1 || steps > kBwMaxSteps) return 0;
    return (channels == 32 || channels == 36) ? 1 : 0;      // (40 channels: 30 chunks + 3 collector lanes do not fit a warp)
}

extern "C" size_t glam_message_stack_bwd_workspace_bytes(int channels, int heads, int edge_dim) {
    const size_t image = channels == 32 ? bwd_image_bytes<8, 3>() : bwd_image_bytes<9, 3>();
    return image + (size_t)kNumSMs * (size_t)(edge_dim * heads * channels + edge_dim * heads) * sizeof(float) + 256;
}

extern "C" int glam_message_stack_bwd(const float* save_x, const float* save_h, const float* save_xpe, const float* save_alpha,
                                      const float* save_m, const float* save_rzn, const float* save_gh, const float* save_gt,
                                      const float* const* h_g_ext, const float* g_h_final, const float* w_ext, int64_t ldw,
                                      const float* w_edge, const float* att_edge, const float* w_scale, const float* w_ih,
                                      const float* w_hh, const int32_t* tiles, const int32_t* tile_meta, const int32_t* dst_rowptr,
                                      const int32_t* dst_src, const uint8_t* etype, const int32_t* src_rowptr, const int32_t* src_pos,
                                      const int32_t* src_dst, int64_t num_nodes, int64_t num_edges, int channels, int heads,
                                      int edge_dim, int steps, float negative_slope, int act, float act_param, int res, int pre_act,
                                      float pre_act_param, float* g_gi,
                                      float* g_gh, float* g4, float* g_pre, float* g_xpe, float* g_x0, float* g_h0, float* g_w_edge, float* g_att_edge,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
    GLAM_REQUIRE(glam_message_stack_bwd_supported(channels, heads, edge_dim, steps),
                 "glam_message_stack_bwd: unsupported (channels=%d heads=%d edge_dim=%d steps=%d math mode %d); use the per-op calls",
                 channels, heads, edge_dim, steps, g_math_mode_get());
    GLAM_REQUIRE(num_nodes >= 0 && num_edges >= 0, "glam_message_stack_bwd: bad sizes");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int HC = heads * channels, ld = (HC + 2 * heads + 3) / 4 * 4, PW = edge_dim * HC + edge_dim * heads;
    GLAM_REQUIRE(g_w_edge && g_att_edge, "glam_message_stack_bwd: null pointer");
    if (num_nodes == 0) {
        cudaMemsetAsync(g_w_edge, 0, sizeof(float) * edge_dim * HC, stream);
        cudaMemsetAsync(g_att_edge, 0, sizeof(float) * edge_dim * heads, stream);
        return 0;
    }
    GLAM_REQUIRE(save_xpe && (save_gt || (save_x && save_h && save_m && save_rzn && save_gh)) && w_ext && w_edge && att_edge && w_scale && w_ih && w_hh &&
                 tiles && tile_meta && dst_rowptr && src_rowptr && (g4 || (g_gi && g_gh)) && g_pre && g_xpe && g_x0 && workspace &&
                 (num_edges == 0 || (save_alpha && dst_src && etype && src_pos && src_dst)),
                 "glam_message_stack_bwd: null pointer");
    GLAM_REQUIRE(pre_act == ACT_NONE || (save_x && !g_h0 && pre_act >= ACT_RELU && pre_act <= ACT_CELU),
                 "glam_message_stack_bwd: pre_act needs save_x (x0 = the activated input projection) and the combined g_x0");
    GLAM_REQUIRE(ldw == ld, "glam_message_stack_bwd: w_ext pitch %lld, expected %d", (long long)ldw, ld);
    GLAM_REQUIRE(workspace_bytes >= glam_message_stack_bwd_workspace_bytes(channels, heads, edge_dim) && al16b(workspace),
                 "glam_message_stack_bwd: workspace too small or not 16-byte aligned");
    GLAM_REQUIRE(al16b(save_x) && al16b(save_h) && al16b(save_xpe) && al16b(save_m) && al16b(save_rzn) && al16b(save_gh) && al16b(save_gt) && al16b(w_ext) &&
                 al16b(w_scale) && al16b(g_gi) && al16b(g_gh) && al16b(g4) && al16b(g_pre) && al16b(g_xpe) && al16b(g_x0) && al16b(g_h0) && al16b(g_h_final) && al16b(tiles),
                 "glam_message_stack_bwd: pointers must be 16-byte aligned");
    GLAM_REQUIRE(num_nodes < ((int64_t)1 << 31) && num_edges < ((int64_t)1 << 31), "glam_message_stack_bwd: too large");
    BwParams p;
    p.X = save_x; p.HH = save_h; p.XPE = save_xpe; p.ALPHA = save_alpha; p.M = save_m; p.RZN = save_rzn; p.GH = save_gh; p.GT = save_gt;
    for (int s = 0; s < kBwMaxSteps; ++s) {
        p.g_ext[s] = (h_g_ext && s < steps) ? h_g_ext[s] : nullptr;
        GLAM_REQUIRE(al16b(p.g_ext[s]), "glam_message_stack_bwd: pointers must be 16-byte aligned");
    }
    p.g_h_final = g_h_final; p.w_ext = w_ext; p.ldw = (int)ldw; p.w_edge = w_edge; p.att_edge = att_edge; p.w_scale = w_scale;
    uint8_t* image = reinterpret_cast<uint8_t*>(workspace);
    const size_t image_bytes = channels == 32 ? bwd_image_bytes<8, 3>() : bwd_image_bytes<9, 3>();
    p.bw_image = image;
    p.partial = reinterpret_cast<float*>(image + ((image_bytes + 255) / 256) * 256);
    p.tiles = reinterpret_cast<const int4*>(tiles); p.meta = tile_meta; p.rowptr = dst_rowptr; p.src = dst_src; p.etype = etype;
    p.src_rowptr = src_rowptr; p.src_pos = src_pos; p.src_dst = src_dst; p.N = num_nodes; p.E = num_edges; p.De = edge_dim;
    p.steps = steps; p.act = act; p.res = res; p.slope = negative_slope; p.act_param = act_param;
    p.G_GI = g_gi; p.G_GH = g_gh; p.G4 = g4; p.G_PRE = g_pre; p.G_XPE = g_xpe; p.g_x0 = g_x0; p.g_h0 = g_h0; p.phase_clock = g_mp_phase_clock;
    p.pre_act = pre_act; p.pre_act_param = pre_act_param;
    int rc = 0;
    switch (channels) {
        case 32: rc = bwd_launch<8, 3>(p, w_ih, w_hh, image, stream); break;
        default: rc = bwd_launch<9, 3>(p, w_ih, w_hh, image, stream); break;
    }
    if (rc) return rc;
    GLAM_CHECK_LAUNCH();
    count_launch(1);
    (void)PW;
    return launch_reduce_partials(p.partial, kNumSMs, edge_dim, HC, edge_dim * heads, g_w_edge, HC, 0, g_att_edge, stream);
}
