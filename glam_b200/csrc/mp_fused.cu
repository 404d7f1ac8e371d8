// The message stack of GLAM as ONE kernel: `message_steps` applications of the (weight-tied) MessageBlock
//   conv = TripletMessage (src_1gp/layer.py:36-61) -> CELU -> GRU (layer.py:260-263) -> (+identity) -> act (:265-266)
// looped by src_1gp/model.py:53-54, on graph-aligned tiles that never leave the SM.
//
// Why: a PyG batch is block-diagonal (src_1gp/dataset.py:75-87 + Batch.from_data_list): no edge crosses a graph, and the
// nodes of a graph are consecutive rows.  A tile made of WHOLE graphs (<= 128 nodes: the M of one tcgen05.mma) is
// therefore closed under message passing — every source row of every in-edge is inside the tile — so the whole chain
//   x --Wn--> xp --(edge softmax / aggregate)--> agg --Wscale,+b,CELU--> m --W_ih | h --W_hh--> gates --> h', x'
// can run out of shared memory / tensor memory, step after step, with x and h resident: per tile HBM sees the input rows
// once, the index words once (not once per step), and the output rows.  Unfused, the same work was 3 launches per step
// for the conv alone with xp [N,HC+2H] and agg [N,HC] written and re-read (6.4x the algorithmic bytes).
//
// Per 128-row tile and step (512 threads, all phases by all warps, CTA barriers between them):
//   P1  one lane issues tcgen05.mma.kind::tf32  xp = x Wn  (A = x panels, B = Wn panels, D in TMEM);
//       meanwhile the 2H attention-logit columns s_i|s_j are computed with exact fp32 FMAs from the staged x rows
//       (softmax gradients are zero-sum per destination: TF32 rounding there is amplified)
//   P2  tcgen05.ld: xp -> shared memory, row-major (the edge phase gathers rows x_j by source slot)
//   P3  segment softmax, a thread per (destination, head), PyG form exp(a-max)/(sum+1e-16)
//   P4  aggregation sum alpha * e_ij (.) x_j into registers, a thread per (destination, 12-channel item)
//   P5  registers -> agg operand panels (they alias the xp tile, dead by now)
//   P6  tcgen05.mma  pre = agg Wscale
//   P7  epilogue: + bias, CELU -> m operand panels
//   P8  tcgen05.mma  gi = m W_ih^T, gh = h W_hh^T
//   P9  gate epilogue (thread = row = TMEM lane): r, z, n, h' = (1-z) n + z h, x' = act(h' + x): written IN PLACE into the
//       x / h operand panels — the next step's P1 reads them there.
// In training ("save" mode) the tensors MessageStackFn.backward consumes (xpe, agg, alpha, m, r|z|n, gh_n, x', h') leave
// as contiguous tile-sized copies out of shared memory; in eval mode (virtual screening) only the outputs do.
//
// Shared-memory operand image: tc_common.cuh's SWIZZLE_128B K-major panels (32 features x 128 rows = 16 KB).  C = 36
// features need 32 + 8: the 8-feature tails of x, h, m share ONE panel (k-slices at byte 0 / 32 / 64 of its rows), and
// likewise the tails of Wn, W_ih, W_hh — a K-slice of a panel is addressed by its start byte, so independent operands
// can live side by side in one panel.  One CTA per SM: 192 KB of panels + 16 KB of records.
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
unsigned long long* g_mp_phase_clock = nullptr;                  // glam_message_stack_phase_clock (shared with mp_fused_bwd.cu)

namespace {


enum : int { kFlagNodes = 1, kFlagEdges = 2, kFlagCross = 4, kFlagEdgeAttr = 8 };

// ------------------------------------------------------------------------------------------------ graph tiles
// Greedy packing of consecutive whole graphs into tiles of <= max_nodes nodes and <= max_edges in-edges, in chunks of 512
// graphs (tiles never span chunks: one partly filled tile per chunk of ~100).  Three small launches:
//   (1) a CTA per chunk: the tiles of the chunk into the chunk's own slots of a scratch array (slot c*512 + i) and the
//       chunk's tile count;
//   (2) one CTA: exclusive scan of the chunk counts (fixed order), total -> meta[0];
//   (3) a warp per chunk: scratch slots -> final positions.
// The kernels' time per tile does not depend on how full the tile is (the phases are lane-parallel over its 128 rows), so the
// ORDER of the graphs in the batch matters: synth.tile_order() / permute_graphs() arrange a batch so that consecutive graphs
// fill the tiles (first-fit decreasing), which the greedy pass here then reproduces.
constexpr int kTileChunk = 512;                 // graphs per warp: a tile never spans chunks, so at most one short tile per 512 graphs
constexpr int kTileWarps = 8;

__global__ void __launch_bounds__(256)
graph_tiles_pack_kernel(const int32_t* __restrict__ gptr, int64_t B, const int32_t* __restrict__ rowptr, int max_nodes, int max_edges,
                        int4* __restrict__ scratch, int32_t* __restrict__ counts, int32_t* __restrict__ meta) {
    // one CTA per chunk: (a) all threads fetch the chunk's node / edge end offsets, (b) every graph finds, by bisection over the
    // (monotone) offsets, where a tile that STARTS at it would end, (c) one thread follows those links from the chunk's first
    // graph — a chain of ~100 shared-memory reads, not 512 dependent global ones —, (d) all threads write the tiles out
    __shared__ int nend[kTileChunk + 1], eend[kTileChunk + 1], nxt[kTileChunk], starts[kTileChunk], n_tiles;
    const int tid = threadIdx.x;
    const int64_t c = blockIdx.x;
    const int64_t g0 = c * kTileChunk;
    if (g0 >= B) return;
    const int ng = (int)min((int64_t)kTileChunk, B - g0);
    for (int i = tid; i <= ng; i += blockDim.x) {
        const int nn = gptr[g0 + i];
        nend[i] = nn;
        eend[i] = rowptr[nn];
    }
    __syncthreads();
    for (int g = tid; g < ng; g += blockDim.x) {
        const int n0 = nend[g], e0 = eend[g];
        int lo = g, hi = ng;                             // largest k in [g, ng] whose graphs g..k-1 fit
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (nend[mid] - n0 <= max_nodes && eend[mid] - e0 <= max_edges) lo = mid; else hi = mid - 1;
        }
        nxt[g] = lo == g ? g + 1 : lo;                   // a single graph over the caps: its own (flagged) tile
    }
    __syncthreads();
    if (tid == 0) {
        int count = 0;
        for (int g = 0; g < ng;) {
            const int k = nxt[g];
            if (nend[k] > nend[g]) starts[count++] = g;
            g = k;
        }
        n_tiles = count;
        counts[c] = count;
    }
    __syncthreads();
    int flags = 0;
    for (int i = tid; i < n_tiles; i += blockDim.x) {
        const int g = starts[i], k = nxt[g];
        const int n0 = nend[g], n1 = nend[k], e0 = eend[g], e1 = eend[k];
        flags |= (n1 - n0 > max_nodes ? kFlagNodes : 0) | (e1 - e0 > max_edges ? kFlagEdges : 0);
        scratch[c * kTileChunk + i] = make_int4(n0, n1, e0, e1);
    }
    if (flags) atomicOr(&meta[1], flags);                // meta is zeroed by the caller; every builder ORs its findings in
}

// counts[c] -> exclusive prefix (in place); total -> meta[0].  One CTA, 1024 threads, each a contiguous run of chunks.
__global__ void __launch_bounds__(1024)
graph_tiles_scan_kernel(int32_t* __restrict__ counts, int64_t chunks, int32_t* __restrict__ meta) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int64_t per = (chunks + 1023) / 1024, c0 = min(chunks, t * per), c1 = min(chunks, c0 + per);
    int sum = 0;
    for (int64_t c = c0; c < c1; ++c) sum += counts[c];
    part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {               // Hillis-Steele inclusive scan (integers: order is irrelevant)
        const int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - sum;
    for (int64_t c = c0; c < c1; ++c) { const int n = counts[c]; counts[c] = run; run += n; }
    if (t == 1023) meta[0] = part[1023];
}

__global__ void __launch_bounds__(kTileWarps * 32)
graph_tiles_emit_kernel(const int4* __restrict__ scratch, const int32_t* __restrict__ offsets, int64_t chunks, const int32_t* __restrict__ meta,
                        int4* __restrict__ tiles) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * kTileWarps + warp;
    if (c >= chunks) return;
    const int base = offsets[c], n = (c + 1 < chunks ? offsets[c + 1] : meta[0]) - base;
    for (int i = lane; i < n; i += 32) tiles[base + i] = scratch[c * kTileChunk + i];
}

// a warp per tile: every source of the tile's in-edges must be one of its own rows (block-diagonal batch)
__global__ void __launch_bounds__(256)
graph_tiles_check_kernel(const int4* __restrict__ tiles, const int32_t* __restrict__ src, int32_t* __restrict__ meta, int64_t max_tiles) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (t >= max_tiles || t >= meta[0]) return;
    const int4 td = tiles[t];
    int bad = 0;
    for (int e = td.z + lane; e < td.w; e += 32) {
        const int j = src[e];
        bad |= (j < td.x || j >= td.y) ? 1 : 0;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&meta[1], kFlagCross);
}

// bond type of every (dst-ordered) edge: index of the 1 in an exactly one-hot edge_attr row (src_1gp/dataset.py:82)
__global__ void __launch_bounds__(256)
edge_types_kernel(const float* __restrict__ ea, const int32_t* __restrict__ perm, int64_t E, int De, uint8_t* __restrict__ etype,
                  int32_t* __restrict__ meta) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t row = perm ? perm[e] : e;              // perm: the dst-ordered edge e is row perm[e] of the caller's edge_attr
    int ty = -1, ok = 1;
    for (int d = 0; d < De; ++d) {
        const float v = ea[row * De + d];
        if (v == 1.f) { ok &= (ty < 0); ty = d; }
        else ok &= (v == 0.f);
    }
    ok &= (ty >= 0) && (ty < kMpMaxDe);
    etype[e] = ok ? (uint8_t)ty : (uint8_t)0;
    if (!ok) atomicOr(&meta[1], kFlagEdgeAttr);
}

// ------------------------------------------------------------------------------------------------ the fused kernel
struct MpParams {
    const float* x0; const float* h0;
    const float* x_raw; const float* w_pre; const float* b_pre; int raw_dim, pre_act; float pre_act_param;   // optional input LinearBlock
    const float* w_ext; int ldw;
    const float* w_edge; const float* att_edge;
    const float* w_scale; const float* bias;
    const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;
    const int4* tiles; const int32_t* meta;
    const int32_t* rowptr; const int32_t* src; const uint8_t* etype;
    int64_t N, E;
    int De, steps, act, res, conv_only, keep_all;
    float slope, act_param;
    float* x_out; float* h_out;
    float* sX; float* sHH; float* sXPE; float* sAGG; float* sALPHA; float* sM; float* sRZN; float* sGH;
    float* sGT;                              // tile-blocked gate save for the one-launch backward (replaces sRZN / sGH), or NULL
    float* sMH;                              // [steps][N][2C+4] = m | h_in | 1 0 0 0 (replaces sM): ONE operand for both GRU weight gradients
    const int64_t* pn_batch; float pn_eps;   // PairNorm on every step's block input (evaluation): graph id per node, or NULL
    unsigned long long* phase_clock;         // profiling aid (glam_message_stack_phase_clock): [grid][16] cycles per phase, or NULL
};


template <int CQ, int H, int NT>
struct MpGeom {
    static constexpr int C = 4 * CQ, HC = H * C, NQ = HC / 4;
    static constexpr int LD = (HC + 2 * H + 3) / 4 * 4;                 // row pitch of the xp tile = ldxp of the unfused path
    static constexpr int KP = (C + 7) / 8 * 8, KSX = KP / 8;            // K of the x/h/m products (k-steps of 8)
    static constexpr int TQ = CQ > 8 ? CQ - 8 : 0;                      // 16-byte chunks in the shared tail panel (<= 2)
    static constexpr int NXP = (HC + 15) / 16 * 16;                     // N of the node projection
    static constexpr int KAGG = (HC + 7) / 8 * 8, KSA = KAGG / 8, NPA = (KAGG + 31) / 32;
    static constexpr int NS = (C + 15) / 16 * 16;                       // N of the scale projection
    static constexpr int NG = (3 * C + 15) / 16 * 16;                   // N of each GRU product
    static constexpr int WROWS = NXP > NG ? NXP : NG;
    static constexpr int TM_XP = 0, TM_PRE = NXP, TM_GI = NXP + NS, TM_GH = NXP + NS + NG, TM_XRES = NXP + NS + 2 * NG, TM_COLS = TM_XRES + C;   // XRES: the pre-norm block input (PairNorm mode)
    static constexpr int WQ = NT / 128;                                 // warps per TMEM lane quarter = threads per tile row
    static constexpr int JPW = (CQ + WQ - 1) / WQ;                      // epilogues: 4-channel chunks per warp
    static constexpr int REGB0 = NPA * kMpPanel, REGB1 = kMpM * LD * 4, REGB2 = kMpM * 3 * C * 4;
    static constexpr int REGB = ((REGB0 > REGB1 ? (REGB0 > REGB2 ? REGB0 : REGB2) : (REGB1 > REGB2 ? REGB1 : REGB2)) + 1023) / 1024 * 1024;
    static constexpr int OFF_XM = 0, OFF_HM = kMpPanel, OFF_AT = 2 * kMpPanel, OFF_REG = 3 * kMpPanel;
    static constexpr int OFF_WN = OFF_REG + REGB, OFF_WI = OFF_WN + NXP * 128, OFF_WH = OFF_WI + NG * 128, OFF_WT = OFF_WH + NG * 128;
    static constexpr int OFF_WS = OFF_WT + WROWS * 128, OFF_MISC = OFF_WS + NPA * NS * 128;
    // misc (floats / ints)
    static constexpr int M_BAR = 0, M_RP = 4, M_REC = 136, M_ALPHA = M_REC + kMpMaxEdges, M_WE = M_ALPHA + kMpMaxEdges * H;
    static constexpr int M_AE = M_WE + kMpMaxDe * HC, M_U = M_AE + kMpMaxDe * H, M_BIAS = M_U + C * 2 * H, M_GB = M_BIAS + C;
    static constexpr int M_PRE = (M_GB + 4 * C + 3) / 4 * 4;            // input LinearBlock: W [C][raw_dim <= kMpMaxRaw] | b [C]
    static constexpr int M_CLK = (M_PRE + C * kMpMaxRaw + C + 3) / 4 * 4;  // 16 phase-cycle counters (profiling aid)
    static constexpr int M_GID = M_CLK + 16;                            // PairNorm: local graph id per tile row (bytes)
    static constexpr int M_END = M_GID + kMpM / 4;
    static constexpr int SMEM = OFF_MISC + M_END * 4;
    static_assert(NT % 128 == 0 && NT >= 256 && NT <= 1024, "threads");
    static_assert(KSX <= 5 && TQ <= 2, "tail panel holds 8 features per operand");
    static_assert(TM_COLS <= 512, "TMEM columns");
    static_assert(NPA <= 4 && NXP <= 256 && NG <= 256, "shape");
    static_assert(KAGG / 4 <= 32, "aggregation: one lane per 16-byte chunk of a row");
    static_assert((NS * 128) % 1024 == 0 && (NXP * 128) % 1024 == 0 && (NG * 128) % 1024 == 0, "panel alignment");
};

// copy nd rows x CQ chunks out of an operand image (main panel + tail-panel slot) to contiguous global rows
template <int CQ, int NT>
__device__ __forceinline__ void copy_out_panels(const uint8_t* mainp, const uint8_t* tailp, int slot, float* __restrict__ dst, int nd) {
    for (int i = threadIdx.x; i < nd * CQ; i += NT) {
        const int r = i / CQ, q = i - r * CQ;
        const float4 v = q < 8 ? lds128(mainp + pan_off(r, q)) : lds128(tailp + pan_off(r, 2 * slot + q - 8));
        reinterpret_cast<float4*>(dst)[i] = v;
    }
}
template <int NT>
__device__ __forceinline__ void copy_out_flat(const void* srcp, float* __restrict__ dst, int n4) {
    for (int i = threadIdx.x; i < n4; i += NT) reinterpret_cast<float4*>(dst)[i] = lds128(reinterpret_cast<const uint8_t*>(srcp) + 16 * i);
}

template <int CQ, int H, int NT, bool SAVE>
__global__ void __launch_bounds__(NT, 1)
mp_fused_kernel(const MpParams p) {
    using G = MpGeom<CQ, H, NT>;
    constexpr int C = G::C, HC = G::HC, NQ = G::NQ, LD = G::LD, WQ = G::WQ;
    // the dynamic shared memory IS the panel arena: every pointer below is derived from this array by plain pointer
    // arithmetic, so the compiler keeps the shared state space (LDS/STS, 32-bit addresses) instead of generic LD/ST
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* XM = sm + G::OFF_XM;  uint8_t* HM = sm + G::OFF_HM;  uint8_t* AT = sm + G::OFF_AT;  uint8_t* REG = sm + G::OFF_REG;
    uint8_t* WN = sm + G::OFF_WN;  uint8_t* WI = sm + G::OFF_WI;  uint8_t* WH = sm + G::OFF_WH;  uint8_t* WT = sm + G::OFF_WT;
    uint8_t* WS = sm + G::OFF_WS;
    float* misc = reinterpret_cast<float*>(sm + G::OFF_MISC);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(misc + G::M_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + G::M_BAR + 2);
    int* rp = reinterpret_cast<int*>(misc + G::M_RP);
    int* rec = reinterpret_cast<int*>(misc + G::M_REC);
    float* alpha_s = misc + G::M_ALPHA;
    float* We = misc + G::M_WE;  float* Ae = misc + G::M_AE;  float* U = misc + G::M_U;
    float* bias_s = misc + G::M_BIAS;  float* gb = misc + G::M_GB;
    float* Wp = misc + G::M_PRE;  float* bp = Wp + C * kMpMaxRaw;
    unsigned int* clk = reinterpret_cast<unsigned int*>(misc + G::M_CLK);
    uint8_t* gid_s = reinterpret_cast<uint8_t*>(misc + G::M_GID);
    float* xp = reinterpret_cast<float*>(REG);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned int clk_start = (unsigned int)clock64();
    const bool mma_thread = tid == NT - 32;                     // lane 0 of the last warp (it has no logit / softmax task) issues every MMA
    const int ntiles = p.meta[0], flags = p.meta[1];

    if (flags != 0) {
        // the batch violates a precondition of this path (graph > 128 nodes / too many edges / edge crossing graphs / edge_attr
        // not one-hot): poison the outputs so the failure cannot go unnoticed (the host checks meta[1] outside graph capture)
        const float nanv = CUDART_NAN_F;
        const int64_t nc = p.N * C, i0 = blockIdx.x * (int64_t)NT + tid, di = (int64_t)gridDim.x * NT;
        if (p.x_out) for (int64_t i = i0; i < (p.keep_all ? (int64_t)p.steps * nc : nc); i += di) p.x_out[i] = nanv;
        if (p.h_out) for (int64_t i = i0; i < nc; i += di) p.h_out[i] = nanv;
        if (p.sX) for (int64_t i = i0; i < (int64_t)(p.steps + 1) * nc; i += di) p.sX[i] = nanv;
        if (p.sHH) for (int64_t i = i0; i < (int64_t)(p.steps + 1) * nc; i += di) p.sHH[i] = nanv;
        return;
    }
    if ((int)blockIdx.x >= ntiles) return;
    if (smem_u32(sm) & 1023u) __trap();                        // SWIZZLE_128B panels need the 1024-byte alignment asked for above

    // ---------------------------------------------------------------- one-time setup: barriers, TMEM, weights
    if (tid == 0) { mbar_init(mma_bar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < (G::OFF_MISC - G::OFF_WN) / 16; i += NT) sts128(WN + 16 * i, z);    // all weight panels
        for (int i = tid; i < kMpPanel / 16; i += NT) sts128(AT + 16 * i, z);                      // tails incl. K padding
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    {
        // Wn: B[n][k] = w_ext[k][n], n < HC
        for (int i = tid; i < HC * CQ; i += NT) {
            const int q = i / HC, n = i - q * HC;                // consecutive lanes: consecutive n (coalesced reads)
            const float* w = p.w_ext + (size_t)(4 * q) * p.ldw + n;
            const float4 v = make_float4(__ldg(w), __ldg(w + p.ldw), __ldg(w + 2 * p.ldw), __ldg(w + 3 * p.ldw));
            sts128(q < 8 ? WN + pan_off(n, q) : WT + pan_off(n, q - 8), v);
        }
        // Wscale: B[n][k] = w_scale[k][n], n < C, k < HC
        for (int i = tid; i < C * NQ; i += NT) {
            const int q = i / C, n = i - q * C;
            const float* w = p.w_scale + (size_t)(4 * q) * C + n;
            const float4 v = make_float4(__ldg(w), __ldg(w + C), __ldg(w + 2 * C), __ldg(w + 3 * C));
            sts128(WS + (q >> 3) * (G::NS * 128) + pan_off(n, q), v);
        }
        if (!p.conv_only) {
            // GRU weights are [3C][C] row-major = B[n][k] as stored
            for (int i = tid; i < 3 * C * CQ; i += NT) {
                const int n = i / CQ, q = i - n * CQ;
                const float4 vi = __ldg(reinterpret_cast<const float4*>(p.w_ih + (size_t)n * C) + q);
                const float4 vh = __ldg(reinterpret_cast<const float4*>(p.w_hh + (size_t)n * C) + q);
                sts128(q < 8 ? WI + pan_off(n, q) : WT + pan_off(n, 2 + q - 8), vi);
                sts128(q < 8 ? WH + pan_off(n, q) : WT + pan_off(n, 4 + q - 8), vh);
            }
            // gate biases per 4-channel chunk j: gb[(4 j + g) 4 + i], g = r (b_ir + b_hr) | z (b_iz + b_hz) | i_n | h_n
            for (int i = tid; i < 4 * C; i += NT) {
                const int j = i >> 4, g = (i >> 2) & 3, c = 4 * j + (i & 3);
                gb[i] = g == 0 ? p.b_ih[c] + p.b_hh[c] : g == 1 ? p.b_ih[C + c] + p.b_hh[C + c] : g == 2 ? p.b_ih[2 * C + c] : p.b_hh[2 * C + c];
            }
        }
        for (int i = tid; i < p.De * HC; i += NT) We[i] = p.w_edge[i];
        for (int i = tid; i < p.De * H; i += NT) Ae[i] = p.att_edge[i];
        // exact logit weights per head and 4-channel chunk q: U[(h CQ + q) 8 + which 4 + i] = u_{i|j},h[4 q + i]
        for (int i = tid; i < C * 2 * H; i += NT) {
            const int h = i / (2 * C), rem = i - h * 2 * C, q = rem >> 3, which = (rem >> 2) & 1, c = 4 * q + (rem & 3);
            U[i] = p.w_ext[(size_t)c * p.ldw + HC + which * H + h];
        }
        for (int i = tid; i < C; i += NT) bias_s[i] = p.bias[i];
        if (p.x_raw) {
            for (int i = tid; i < C * p.raw_dim; i += NT) Wp[i] = p.w_pre[i];
            for (int i = tid; i < C; i += NT) bp[i] = p.b_pre ? p.b_pre[i] : 0.f;
        }
    }
    fence_proxy_async_smem();
    __syncthreads();

    const uint32_t xm_a = smem_u32(XM), hm_a = smem_u32(HM), at_a = smem_u32(AT), reg_a = smem_u32(REG);
    const uint32_t wn_a = smem_u32(WN), wi_a = smem_u32(WI), wh_a = smem_u32(WH), wt_a = smem_u32(WT), ws_a = smem_u32(WS);
    // operand descriptors once per kernel: every operand lives at a fixed shared-memory address, and a K-slice (+32 B) or panel
    // step is a plain add on the encoded start-address field (bytes >> 4) — building a descriptor per MMA cost ~80 cycles each
    // in the issuing thread (29 MMAs per step)
    const uint64_t d_xm = make_smem_desc(xm_a, 16, 1024), d_hm = make_smem_desc(hm_a, 16, 1024), d_at = make_smem_desc(at_a, 16, 1024);
    const uint64_t d_reg = make_smem_desc(reg_a, 16, 1024), d_wn = make_smem_desc(wn_a, 16, 1024), d_wi = make_smem_desc(wi_a, 16, 1024);
    const uint64_t d_wh = make_smem_desc(wh_a, 16, 1024), d_wt = make_smem_desc(wt_a, 16, 1024), d_ws = make_smem_desc(ws_a, 16, 1024);
    auto dsc = [](uint64_t base, uint32_t byte_off) { return base + (uint64_t)(byte_off >> 4); };
    const int q4 = warp & 3, cg = warp >> 2;                     // TMEM lane quarter of this warp / its index among the quarter's warps
    const int row = q4 * 32 + lane;                              // the tile row this thread owns in the epilogues
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    uint32_t ph = 0;
    const int64_t NC = p.N * C;

    // phase clock (profiling aid): thread 0 adds the cycles since its previous tick to counter I — only when a buffer was given
    unsigned int clk_last = (unsigned int)clock64();
    if (p.phase_clock && tid < 16) clk[tid] = tid == 14 ? clk_last - clk_start : 0;      // [14] = set-up (weights, TMEM)
#define MP_TICK(I)                                                              \
    if (p.phase_clock && tid == 0) {                                            \
        const unsigned int now_ = (unsigned int)clock64();                      \
        clk[I] += now_ - clk_last;                                              \
        clk_last = now_;                                                        \
    }
    // ---------------------------------------------------------------- tile loop: the next tile's words travel in registers
    // while the current tile computes (x rows or raw feature words, source / bond-type words, row pointers)
    constexpr int XPF = (kMpM * CQ + NT - 1) / NT;               // float4 of x rows per thread
    constexpr int RPF = (kMpM * kMpMaxRaw + NT - 1) / NT;        // raw feature words per thread (input LinearBlock mode)
    constexpr int EPF = (kMpMaxEdges + NT - 1) / NT;             // edges per thread
    float4 pfx[XPF];
    float pfr[RPF];
    int pfs[EPF], pft[EPF], pfp = 0;
    const bool raw = p.x_raw != nullptr;
#define MP_PREFETCH(TD)                                                                                              \
    {                                                                                                                \
        const int n0_ = (TD).x, nd_ = (TD).y - (TD).x, e0_ = (TD).z, ne_ = (TD).w - (TD).z;                          \
        if (raw) {                                                                                                   \
            const float* xr_ = p.x_raw + (size_t)n0_ * p.raw_dim;                                                    \
            _Pragma("unroll") for (int k = 0; k < RPF; ++k) {                                                        \
                const int i_ = k * NT + tid;                                                                         \
                pfr[k] = i_ < nd_ * p.raw_dim ? __ldg(xr_ + i_) : 0.f;                                               \
            }                                                                                                        \
        } else {                                                                                                     \
            const float4* xr_ = reinterpret_cast<const float4*>(p.x0 + (size_t)n0_ * C);                             \
            _Pragma("unroll") for (int k = 0; k < XPF; ++k) {                                                        \
                const int i_ = k * NT + tid;                                                                         \
                pfx[k] = i_ < nd_ * CQ ? __ldg(xr_ + i_) : make_float4(0.f, 0.f, 0.f, 0.f);                          \
            }                                                                                                        \
        }                                                                                                            \
        _Pragma("unroll") for (int k = 0; k < EPF; ++k) {                                                            \
            const int e_ = k * NT + tid;                                                                             \
            pfs[k] = e_ < ne_ ? __ldg(p.src + e0_ + e_) - n0_ : 0;                                                   \
            pft[k] = e_ < ne_ ? (int)__ldg(p.etype + e0_ + e_) : 0;                                                  \
        }                                                                                                            \
        if (tid <= kMpM) pfp = tid <= nd_ ? __ldg(p.rowptr + n0_ + tid) - e0_ : ne_;                                 \
    }
    int4 td = p.tiles[blockIdx.x];
    int4 tdn = (int)(blockIdx.x + gridDim.x) < ntiles ? p.tiles[blockIdx.x + gridDim.x] : make_int4(0, 0, 0, 0);
    MP_PREFETCH(td)

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n0 = td.x, nd = td.y - td.x, e0 = td.z, ne = td.w - td.z;
        // ------------------------------------------------------------ tile load: registers -> operand panels / index words
        if (raw) {
            // the model's input LinearBlock (src_1gp/model.py:49: Linear(raw_dim -> C) + activation) applied while the tile is
            // loaded: the raw rows are staged (the xp region is free here), then exact fp32 FMAs (what the per-op path does
            // for K % 4 != 0); x0 never exists in HBM
            float* rs = reinterpret_cast<float*>(REG);
#pragma unroll
            for (int k = 0; k < RPF; ++k) rs[k * NT + tid] = pfr[k];
            __syncthreads();
            // a thread owns one 4-channel chunk of FOUR consecutive rows: the chunk's weights are read once per feature for the
            // four rows (8 LDS per 16 FMAs; a thread per (row, chunk) re-read them per row: 2.5x the shared-memory traffic,
            // 10 % of the screening kernel); lanes run over the chunks of a row group, so panel and HBM stores stay contiguous
            constexpr int RG = 4;
            for (int it = tid; it < (kMpM / RG) * CQ; it += NT) {
                const int rg = it / CQ, q = it - rg * CQ;
                const int rd = p.raw_dim;
                const float* w = Wp + 4 * q * rd;
                const float* xr = rs + rg * RG * rd;
                const float4 b4 = lds128(bp + 4 * q);
                float4 v[RG];
#pragma unroll
                for (int j = 0; j < RG; ++j) v[j] = b4;
                for (int k = 0; k < rd; ++k) {
                    const float w0 = w[k], w1 = w[rd + k], w2 = w[2 * rd + k], w3 = w[3 * rd + k];
#pragma unroll
                    for (int j = 0; j < RG; ++j) {
                        const float xv = xr[j * rd + k];
                        v[j].x = fmaf(xv, w0, v[j].x); v[j].y = fmaf(xv, w1, v[j].y);
                        v[j].z = fmaf(xv, w2, v[j].z); v[j].w = fmaf(xv, w3, v[j].w);
                    }
                }
#pragma unroll
                for (int j = 0; j < RG; ++j) {
                    const int r = rg * RG + j;
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nd) {
                        o.x = act_fwd(v[j].x, p.pre_act, p.pre_act_param); o.y = act_fwd(v[j].y, p.pre_act, p.pre_act_param);
                        o.z = act_fwd(v[j].z, p.pre_act, p.pre_act_param); o.w = act_fwd(v[j].w, p.pre_act, p.pre_act_param);
                        if (SAVE && !p.conv_only) {
                            reinterpret_cast<float4*>(p.sX + (size_t)(n0 + r) * C)[q] = o;
                            if (!p.h0) reinterpret_cast<float4*>(p.sHH + (size_t)(n0 + r) * C)[q] = o;
                        }
                    }
                    sts128(q < 8 ? XM + pan_off(r, q) : AT + pan_off(r, q - 8), o);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < XPF; ++k) {
                const int i = k * NT + tid;
                if (i < kMpM * CQ) {
                    const int r = i / CQ, q = i - r * CQ;
                    const float4 v = pfx[k];                     // zeros beyond the tile's rows
                    if (SAVE && !p.conv_only && r < nd) {
                        reinterpret_cast<float4*>(p.sX + (size_t)(n0 + r) * C)[q] = v;
                        if (!p.h0) reinterpret_cast<float4*>(p.sHH + (size_t)(n0 + r) * C)[q] = v;
                    }
                    sts128(q < 8 ? XM + pan_off(r, q) : AT + pan_off(r, q - 8), v);
                }
            }
        }
        if (p.h0)
            for (int i = tid; i < kMpM * CQ; i += NT) {
                const int r = i / CQ, q = i - r * CQ;
                float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nd) {
                    hv = __ldg(reinterpret_cast<const float4*>(p.h0 + (size_t)(n0 + r) * C) + q);
                    if (SAVE && !p.conv_only) reinterpret_cast<float4*>(p.sHH + (size_t)(n0 + r) * C)[q] = hv;
                }
                sts128(q < 8 ? HM + pan_off(r, q) : AT + pan_off(r, 2 + q - 8), hv);
            }
        if (tid <= kMpM) rp[tid] = pfp;
#pragma unroll
        for (int k = 0; k < EPF; ++k) {
            const int e = k * NT + tid;
            if (e < ne) rec[e] = pfs[k] | (pft[k] << 8);
        }
        if (p.pn_batch) {                                        // PairNorm: local graph id of every tile row (0 .. 127)
            const int64_t gfirst = p.pn_batch[n0];
            for (int i = tid; i < kMpM; i += NT) gid_s[i] = i < nd ? (uint8_t)(p.pn_batch[n0 + i] - gfirst) : (uint8_t)255;
        }
        fence_proxy_async_smem();
        __syncthreads();
        MP_TICK(0)
        {
            const int tnn = t + 2 * (int)gridDim.x;
            const int4 tdnn = tnn < ntiles ? p.tiles[tnn] : make_int4(0, 0, 0, 0);
            if (t + (int)gridDim.x < ntiles) MP_PREFETCH(tdn)
            td = tdn; tdn = tdnn;                                // (the current tile's n0 / nd / e0 / ne were taken above)
        }

        for (int s = 0; s < p.steps; ++s) {
            const bool pn = p.pn_batch != nullptr;
            const bool h_is_x = (s == 0 && p.h0 == nullptr && !pn); // first step: h = x (layer.py:253-254); PairNorm copies x into the h panels
            if (pn) {
                // ---------------------------------------------------- P0: PairNorm of the block input per graph (PyG PairNorm(scale=1) @1.7.2
                // behind _PairNorm, src_1gp/layer.py:179-185,255): xc = x - mean_g(x), y = xc / sqrt(eps + mean_g(sum_c xc^2)).
                // The tile holds whole graphs, so the statistics are tile-local; the pre-norm rows are parked in tensor memory
                // for the residual (layer.py:264) and, on the first step, copied into the h panels (h = x BEFORE the norm, :253)
                float* mu = reinterpret_cast<float*>(REG);           // [graphs of the tile <= 128][C]   (the region is free here)
                float* rsq = mu + kMpM * C;                          // [128] squared deviation of a row
                float* isg = rsq + kMpM;                             // [128] 1 / sqrt(eps + ...) per graph
                int* gst = reinterpret_cast<int*>(isg + kMpM);       // [129] first row of every graph
                const int ng = nd > 0 ? (int)gid_s[nd - 1] + 1 : 0;
                for (int r = tid; r < nd; r += NT)
                    if (r == 0 || gid_s[r] != gid_s[r - 1]) gst[gid_s[r]] = r;
                if (tid == 0) gst[ng] = nd;
                __syncthreads();
                for (int i = tid; i < ng * C; i += NT) {             // mean per (graph, channel), rows in order
                    const int k = i / C, c = i - k * C, q = c >> 2;
                    float acc = 0.f;
                    for (int r = gst[k]; r < gst[k + 1]; ++r)
                        acc += *reinterpret_cast<const float*>((q < 8 ? XM + pan_off(r, q) : AT + pan_off(r, q - 8)) + 4 * (c & 3));
                    mu[i] = acc / (float)(gst[k + 1] - gst[k]);
                }
                __syncthreads();
                for (int r = tid; r < nd; r += NT) {                 // squared deviation of every row
                    const float4* m4 = reinterpret_cast<const float4*>(mu + gid_s[r] * C);
                    float acc = 0.f;
#pragma unroll
                    for (int q = 0; q < CQ; ++q) {
                        const float4 v = q < 8 ? lds128(XM + pan_off(r, q)) : lds128(AT + pan_off(r, q - 8)), m = m4[q];
                        const float dx = v.x - m.x, dy = v.y - m.y, dz = v.z - m.z, dw = v.w - m.w;
                        acc = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, acc))));
                    }
                    rsq[r] = acc;
                }
                __syncthreads();
                for (int k = tid; k < ng; k += NT) {
                    float acc = 0.f;
                    for (int r = gst[k]; r < gst[k + 1]; ++r) acc += rsq[r];
                    isg[k] = 1.f / sqrtf(p.pn_eps + acc / (float)(gst[k + 1] - gst[k]));
                }
                __syncthreads();
#pragma unroll
                for (int jj = 0; jj < G::JPW; ++jj) {
                    const int j = cg + WQ * jj;
                    if (j < CQ) {                                    // warp-uniform
                        uint8_t* xd = j < 8 ? XM + pan_off(row, j) : AT + pan_off(row, j - 8);
                        const float4 v = lds128(xd);
                        tmem_st4(lane_base + G::TM_XRES + 4 * j, v);
                        if (s == 0 && p.h0 == nullptr) sts128(j < 8 ? HM + pan_off(row, j) : AT + pan_off(row, 2 + j - 8), v);
                        if (row < nd) {
                            const int k = gid_s[row];
                            const float4 m = lds128(mu + k * C + 4 * j);
                            const float sc = isg[k];
                            sts128(xd, make_float4((v.x - m.x) * sc, (v.y - m.y) * sc, (v.z - m.z) * sc, (v.w - m.w) * sc));
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before_sync();
                fence_proxy_async_smem();
                __syncthreads();
            }
            // -------------------------------------------------------- P1: xp = x Wn on the tensor core (and, right behind it,
            // gh = h W_hh^T, which depends on nothing this step computes); exact logit columns on the CUDA cores meanwhile
            if (mma_thread) {
                const unsigned int ci0 = (unsigned int)clock64();
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, G::NXP, 0, 0);
#pragma unroll
                for (int ks = 0; ks < G::KSX; ++ks) {
                    const uint64_t a = ks < 4 ? dsc(d_xm, ks * 32) : d_at;
                    const uint64_t b = ks < 4 ? dsc(d_wn, ks * 32) : d_wt;
                    mma_tf32_ss(tmem_base + G::TM_XP, a, b, idesc, ks > 0 ? 1u : 0u);
                }
                mma_commit(mma_bar);
                if (!p.conv_only) {
                    const uint32_t idesc_g = make_idesc_tf32(kMpM, G::NG, 0, 0);
#pragma unroll
                    for (int ks = 0; ks < G::KSX; ++ks) {
                        const uint64_t a = ks < 4 ? dsc(h_is_x ? d_xm : d_hm, ks * 32) : dsc(d_at, h_is_x ? 0 : 32);
                        const uint64_t b = ks < 4 ? dsc(d_wh, ks * 32) : dsc(d_wt, 64);
                        mma_tf32_ss(tmem_base + G::TM_GH, a, b, idesc_g, ks > 0 ? 1u : 0u);
                    }                                            // completes before the P6 / P8 commits do (in-order)
                }
                if (p.phase_clock) clk[13] += (unsigned int)clock64() - ci0;
            }
            // a thread per (row, head): s_i and s_j of that head from one pass over the row, two channels per FFMA2
            for (int task = tid; task < kMpM * H; task += NT) {
                const int r = task & (kMpM - 1), h = task >> 7;
                float2 ai = f2(0.f), aj = f2(0.f), bi = f2(0.f), bj = f2(0.f);
                const float4* u = reinterpret_cast<const float4*>(U + h * 2 * C);
#pragma unroll
                for (int q = 0; q < CQ; ++q) {
                    const float4 v = q < 8 ? lds128(XM + pan_off(r, q)) : lds128(AT + pan_off(r, q - 8));
                    const float4 ui = u[2 * q], uj = u[2 * q + 1];
                    ai = fma2(f2(v.x, v.y), f2(ui.x, ui.y), ai); bi = fma2(f2(v.z, v.w), f2(ui.z, ui.w), bi);
                    aj = fma2(f2(v.x, v.y), f2(uj.x, uj.y), aj); bj = fma2(f2(v.z, v.w), f2(uj.z, uj.w), bj);
                }
                ai = add2(ai, bi); aj = add2(aj, bj);
                xp[r * LD + HC + h] = ai.x + ai.y;
                xp[r * LD + HC + H + h] = aj.x + aj.y;
            }
            if (LD > HC + 2 * H)
                for (int i = tid; i < kMpM * (LD - HC - 2 * H); i += NT) {
                    const int r = i / (LD - HC - 2 * H), k = i - r * (LD - HC - 2 * H);
                    xp[r * LD + HC + 2 * H + k] = 0.f;
                }
            __syncthreads();
            MP_TICK(1)
            // -------------------------------------------------------- P3: segment softmax per (destination, head) — needs only
            // the logit columns, so it runs while the projection is still on the tensor core
            for (int task = tid; task < nd * H; task += NT) {
                const int d = task / H, h = task - d * H;
                const int beg = rp[d], end = rp[d + 1], deg = end - beg;
                const float si = xp[d * LD + HC + h];
                if (deg <= 4) {
                    // molecular graphs: valence-bounded in-degree — registers only, no shared-memory round trips
                    float l[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        l[k] = -INFINITY;
                        if (k < deg) {
                            const int rc = rec[beg + k];
                            const float v = si + Ae[(rc >> 8) * H + h] + xp[(rc & 0xff) * LD + HC + H + h];
                            l[k] = v > 0.f ? v : p.slope * v;
                        }
                    }
                    const float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
                    float sum = 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                // same summation order as the loop below
                        l[k] = k < deg ? ex2_approx((l[k] - mx) * 1.4426950408889634f) : 0.f;
                        if (k < deg) sum += l[k];
                    }
                    const float inv = 1.f / (sum + 1e-16f);      // PyG: exp(a - max) / (sum + 1e-16)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < deg) alpha_s[(beg + k) * H + h] = l[k] * inv;
                    continue;
                }
                float mx = -INFINITY, sum = 0.f;
                for (int e = beg; e < end; ++e) {
                    const int rc = rec[e];
                    float l = si + Ae[(rc >> 8) * H + h] + xp[(rc & 0xff) * LD + HC + H + h];
                    l = l > 0.f ? l : p.slope * l;
                    alpha_s[e * H + h] = l;
                    mx = fmaxf(mx, l);
                }
                for (int e = beg; e < end; ++e) {
                    const float x = ex2_approx((alpha_s[e * H + h] - mx) * 1.4426950408889634f);
                    alpha_s[e * H + h] = x;
                    sum += x;
                }
                const float inv = 1.f / (sum + 1e-16f);
                for (int e = beg; e < end; ++e) alpha_s[e * H + h] = alpha_s[e * H + h] * inv;
            }
            MP_TICK(2)
            mbar_wait_guarded(mma_bar, ph); ph ^= 1u;
            MP_TICK(3)
            tc_fence_after_sync();
            // -------------------------------------------------------- P2: TMEM -> xp tile (row-major, pitch LD)
            for (int c0 = 16 * cg; c0 < HC; c0 += 16 * WQ) {
                float v[16];
                tmem_ld16(lane_base + G::TM_XP + c0, v);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c0 + 4 * i < HC) sts128(xp + row * LD + c0 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
            tc_fence_before_sync();
            __syncthreads();
            MP_TICK(4)
            if (SAVE) {
                copy_out_flat<NT>(xp, p.sXPE + ((size_t)s * p.N + n0) * LD, nd * LD / 4);
                float* ad = p.sALPHA + ((size_t)s * p.E + e0) * H;
                for (int i = tid; i < ne * H; i += NT) ad[i] = alpha_s[i];
            }
            // -------------------------------------------------------- P4: aggregate: a warp takes RPW consecutive destination
            // rows one after the other, lane = 16-byte chunk of the row (NQ of 32 lanes): the edge loop is warp-uniform (no
            // divergence over in-degrees), x_j rows are read as contiguous runs, the record / alpha words are broadcasts and the
            // lane's slice of weight_edge stays in registers (one float4 per bond type)
            constexpr int NW = NT / 32, RPW = (kMpM + NW - 1) / NW;
            float4 acc[RPW];
            {
                const int lq = lane < NQ ? lane : 0, hq = lq / CQ;
                float4 wt[kMpMaxDe];
#pragma unroll
                for (int ty = 0; ty < kMpMaxDe; ++ty) wt[ty] = ty < p.De ? lds128(We + ty * HC + 4 * lq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    const int d = warp * RPW + k;
                    float2 a0 = f2(0.f), a1 = f2(0.f);
                    if (d < kMpM) {
                        const int beg = rp[d], end = rp[d + 1];
                        for (int e = beg; e < end; ++e) {
                            const int rc = rec[e], ty = rc >> 8;
                            const float c = alpha_s[e * H + hq];
                            const float4 xj = lds128(xp + (rc & 0xff) * LD + 4 * lq);
                            const float4 w = ty == 0 ? wt[0] : ty == 1 ? wt[1] : ty == 2 ? wt[2] : wt[3];
                            a0 = fma2(f2(c), mul2(f2(xj.x, xj.y), f2(w.x, w.y)), a0);
                            a1 = fma2(f2(c), mul2(f2(xj.z, xj.w), f2(w.z, w.w)), a1);
                        }
                    }
                    acc[k] = make_float4(a0.x, a0.y, a1.x, a1.y);
                }
            }
            __syncthreads();                                     // every read of the xp tile is done: its bytes become the agg panels
            MP_TICK(5)
            // -------------------------------------------------------- P5: registers -> agg operand panels
            if (lane < G::KAGG / 4) {                            // chunks in [NQ, KAGG/4) are the K padding: zeros
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    const int d = warp * RPW + k;
                    if (d < kMpM) sts128(REG + (lane >> 3) * kMpPanel + pan_off(d, lane), lane < NQ ? acc[k] : make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
            fence_proxy_async_smem();
            __syncthreads();
            MP_TICK(6)
            // -------------------------------------------------------- P6: pre = agg Wscale
            if (mma_thread) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, G::NS, 0, 0);
#pragma unroll
                for (int ks = 0; ks < G::KSA; ++ks) {
                    const uint32_t pan = ks >> 2, within = (ks & 3) * 32;
                    mma_tf32_ss(tmem_base + G::TM_PRE, dsc(d_reg, pan * kMpPanel + within), dsc(d_ws, pan * (G::NS * 128) + within), idesc,
                                ks > 0 ? 1u : 0u);
                }
                mma_commit(mma_bar);
            }
            if (SAVE) {
                float* ag = p.sAGG + ((size_t)s * p.N + n0) * HC;
                for (int i = tid; i < nd * NQ; i += NT) {
                    const int r = i / NQ, q = i - r * NQ;
                    reinterpret_cast<float4*>(ag)[i] = lds128(REG + (q >> 3) * kMpPanel + pan_off(r, q));
                }
            }
            mbar_wait_guarded(mma_bar, ph); ph ^= 1u;
            MP_TICK(7)
            tc_fence_after_sync();
            if (SAVE) __syncthreads();                           // agg copy-out done before m overwrites panel 0
            // -------------------------------------------------------- P7: epilogue: + bias, CELU -> m operand panels (or the conv output)
            float* ostage = reinterpret_cast<float*>(REG + 2 * kMpPanel);        // conv-only: row-major [128][C] output staging
#pragma unroll
            for (int jj = 0; jj < G::JPW; ++jj) {
                const int cq = cg + WQ * jj;
                if (cq < CQ) {                                   // warp-uniform
                    float v[4];
                    tmem_ld4(lane_base + G::TM_PRE + 4 * cq, v);
                    const float4 b = lds128(bias_s + 4 * cq);
                    float2 m01 = add2(f2(v[0], v[1]), f2(b.x, b.y)), m23 = add2(f2(v[2], v[3]), f2(b.z, b.w));
                    if (p.conv_only) {
                        sts128(ostage + row * C + 4 * cq, make_float4(m01.x, m01.y, m23.x, m23.y));
                    } else {
                        m01 = celu2(m01); m23 = celu2(m23);
                        sts128(cq < 8 ? REG + pan_off(row, cq) : AT + pan_off(row, 4 + cq - 8), make_float4(m01.x, m01.y, m23.x, m23.y));
                        if (SAVE && p.sGT && row < nd)
                            reinterpret_cast<float4*>(p.sGT + ((size_t)s * p.N + n0) * 7 * C)[(6 * CQ + cq) * nd + row] = make_float4(m01.x, m01.y, m23.x, m23.y);
                    }
                }
            }
            tc_fence_before_sync();
            fence_proxy_async_smem();
            __syncthreads();
            MP_TICK(8)
            if (p.conv_only) {
                copy_out_flat<NT>(ostage, p.x_out + (size_t)n0 * C, nd * CQ);
                __syncthreads();
                continue;
            }
            // -------------------------------------------------------- P8: gi = m W_ih^T (gh was issued with the projection)
            if (mma_thread) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, G::NG, 0, 0);
#pragma unroll
                for (int ks = 0; ks < G::KSX; ++ks) {
                    const uint64_t a = ks < 4 ? dsc(d_reg, ks * 32) : dsc(d_at, 64);
                    const uint64_t b = ks < 4 ? dsc(d_wi, ks * 32) : dsc(d_wt, 32);
                    mma_tf32_ss(tmem_base + G::TM_GI, a, b, idesc, ks > 0 ? 1u : 0u);
                }
                mma_commit(mma_bar);
            }
            if (SAVE && p.sMH) {
                // m | h_in | 1: the row operand of the merged GRU weight-gradient contraction [m h 1]^T [g_r g_z g_n g_n r] (the
                // constant column yields both bias gradients as a row of the product)
                float4* mh = reinterpret_cast<float4*>(p.sMH + ((size_t)s * p.N + n0) * (2 * C + 4));
                const uint8_t* hp = h_is_x ? XM : HM;
                const int hslot = h_is_x ? 0 : 1;
                for (int i = tid; i < nd * (2 * CQ + 1); i += NT) {
                    const int r = i / (2 * CQ + 1), q = i - r * (2 * CQ + 1);
                    float4 v = make_float4(1.f, 0.f, 0.f, 0.f);
                    if (q < CQ) v = q < 8 ? lds128(REG + pan_off(r, q)) : lds128(AT + pan_off(r, 4 + q - 8));
                    else if (q < 2 * CQ) { const int qh = q - CQ; v = qh < 8 ? lds128(hp + pan_off(r, qh)) : lds128(AT + pan_off(r, 2 * hslot + qh - 8)); }
                    mh[i] = v;
                }
            }
            if (SAVE && p.sM) copy_out_panels<CQ, NT>(REG, AT, 2, p.sM + ((size_t)s * p.N + n0) * C, nd);
            mbar_wait_guarded(mma_bar, ph); ph ^= 1u;
            MP_TICK(9)
            tc_fence_after_sync();
            if (SAVE) __syncthreads();                           // m copy-out done before the r|z|n staging overwrites the region
            // -------------------------------------------------------- P9: gates on channel pairs; h' and x' in place
            float* rstage = reinterpret_cast<float*>(REG);       // save mode: r|z|n rows, pitch 3C
            float4 ghn[G::JPW];
#pragma unroll
            for (int jj = 0; jj < G::JPW; ++jj) {
                const int j = cg + WQ * jj;
                ghn[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < CQ) {                                    // warp-uniform
                    float v[24];
                    const uint32_t gi = lane_base + G::TM_GI + 4 * j, gh = lane_base + G::TM_GH + 4 * j;
                    tmem_ld4x6(gi, gi + C, gi + 2 * C, gh, gh + C, gh + 2 * C, v);
                    const uint8_t* hsrc = h_is_x ? (j < 8 ? XM + pan_off(row, j) : AT + pan_off(row, j - 8))
                                                 : (j < 8 ? HM + pan_off(row, j) : AT + pan_off(row, 2 + j - 8));
                    uint8_t* xdst = j < 8 ? XM + pan_off(row, j) : AT + pan_off(row, j - 8);
                    uint8_t* hdst = j < 8 ? HM + pan_off(row, j) : AT + pan_off(row, 2 + j - 8);
                    const float4 hv = lds128(hsrc);
                    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.res) { if (pn) { float t4[4]; tmem_ld4(lane_base + G::TM_XRES + 4 * j, t4); xv = make_float4(t4[0], t4[1], t4[2], t4[3]); } else xv = lds128(xdst); }
                    const float4 b_r = lds128(gb + 16 * j), b_z = lds128(gb + 16 * j + 4), b_n = lds128(gb + 16 * j + 8), b_h = lds128(gb + 16 * j + 12);
                    float2 r[2], z[2], nn[2], gn[2], hw[2], xo[2];
#define MP_GATE2(k, i, BR, BZ, BN, BH, HV, XV)                                                                    \
    {                                                                                                              \
        gn[k] = add2(f2(v[20 + i], v[21 + i]), BH);                                                                \
        r[k] = sigmoid2(add2(add2(f2(v[i], v[i + 1]), f2(v[12 + i], v[13 + i])), BR));                             \
        z[k] = sigmoid2(add2(add2(f2(v[4 + i], v[5 + i]), f2(v[16 + i], v[17 + i])), BZ));                         \
        nn[k] = tanh2(fma2(r[k], gn[k], add2(f2(v[8 + i], v[9 + i]), BN)));                                        \
        hw[k] = fma2(fma2(z[k], f2(-1.f), f2(1.f)), nn[k], mul2(z[k], HV));                                        \
        xo[k] = add2(hw[k], XV);                                                                                   \
    }
                    MP_GATE2(0, 0, f2(b_r.x, b_r.y), f2(b_z.x, b_z.y), f2(b_n.x, b_n.y), f2(b_h.x, b_h.y), f2(hv.x, hv.y), f2(xv.x, xv.y))
                    MP_GATE2(1, 2, f2(b_r.z, b_r.w), f2(b_z.z, b_z.w), f2(b_n.z, b_n.w), f2(b_h.z, b_h.w), f2(hv.z, hv.w), f2(xv.z, xv.w))
#undef MP_GATE2
                    if (p.act == ACT_CELU) { xo[0] = celu2(xo[0]); xo[1] = celu2(xo[1]); }
                    else if (p.act != ACT_NONE) {                // relu = leaky with slope 0
                        const float sl = p.act == ACT_LEAKY ? p.act_param : 0.f;
                        xo[0].x = xo[0].x > 0.f ? xo[0].x : sl * xo[0].x; xo[0].y = xo[0].y > 0.f ? xo[0].y : sl * xo[0].y;
                        xo[1].x = xo[1].x > 0.f ? xo[1].x : sl * xo[1].x; xo[1].y = xo[1].y > 0.f ? xo[1].y : sl * xo[1].y;
                    }
                    sts128(hdst, make_float4(hw[0].x, hw[0].y, hw[1].x, hw[1].y));
                    sts128(xdst, make_float4(xo[0].x, xo[0].y, xo[1].x, xo[1].y));
                    if (SAVE && p.sGT) {
                        // tile-blocked save for the one-launch backward: slot-major inside the tile ([r z n gh_n h x' m] x CQ chunks,
                        // then the tile's rows), so a warp's 32 rows of one chunk are 512 contiguous bytes here AND in its gate phase
                        if (row < nd) {
                            float4* gt = reinterpret_cast<float4*>(p.sGT + ((size_t)s * p.N + n0) * 7 * C);
                            gt[j * nd + row] = make_float4(r[0].x, r[0].y, r[1].x, r[1].y);
                            gt[(CQ + j) * nd + row] = make_float4(z[0].x, z[0].y, z[1].x, z[1].y);
                            gt[(2 * CQ + j) * nd + row] = make_float4(nn[0].x, nn[0].y, nn[1].x, nn[1].y);
                            gt[(3 * CQ + j) * nd + row] = make_float4(gn[0].x, gn[0].y, gn[1].x, gn[1].y);
                            gt[(4 * CQ + j) * nd + row] = hv;
                            gt[(5 * CQ + j) * nd + row] = make_float4(xo[0].x, xo[0].y, xo[1].x, xo[1].y);
                        }
                    } else if (SAVE) {
                        sts128(rstage + row * 3 * C + 4 * j, make_float4(r[0].x, r[0].y, r[1].x, r[1].y));
                        sts128(rstage + row * 3 * C + C + 4 * j, make_float4(z[0].x, z[0].y, z[1].x, z[1].y));
                        sts128(rstage + row * 3 * C + 2 * C + 4 * j, make_float4(nn[0].x, nn[0].y, nn[1].x, nn[1].y));
                        ghn[jj] = make_float4(gn[0].x, gn[0].y, gn[1].x, gn[1].y);
                    }
                }
            }
            tc_fence_before_sync();
            fence_proxy_async_smem();
            __syncthreads();
            MP_TICK(10)
            // -------------------------------------------------------- outputs of the step
            if (SAVE) {
                copy_out_panels<CQ, NT>(XM, AT, 0, p.sX + (size_t)(s + 1) * NC + (size_t)n0 * C, nd);
                // with the tile-blocked gate save the backward reads h_in from GT and the weight gradient from MH: only the final
                // state (an output of the node) is written row-major then
                if (!p.sGT || s + 1 == p.steps) copy_out_panels<CQ, NT>(HM, AT, 1, p.sHH + (size_t)(s + 1) * NC + (size_t)n0 * C, nd);
                if (!p.sGT) {
                    copy_out_flat<NT>(rstage, p.sRZN + ((size_t)s * p.N + n0) * 3 * C, nd * 3 * CQ);
                    __syncthreads();
                    float* gstage = reinterpret_cast<float*>(REG);   // gh_n rows, pitch C
#pragma unroll
                    for (int jj = 0; jj < G::JPW; ++jj) {
                        const int j = cg + WQ * jj;
                        if (j < CQ) sts128(gstage + row * C + 4 * j, ghn[jj]);
                    }
                    __syncthreads();
                    copy_out_flat<NT>(gstage, p.sGH + ((size_t)s * p.N + n0) * C, nd * CQ);
                }
                __syncthreads();                                 // before the next step's P1 writes the logit columns into the region / the panels
            } else {
                const bool last = s == p.steps - 1;
                if (p.keep_all || last)
                    copy_out_panels<CQ, NT>(XM, AT, 0, p.x_out + (p.keep_all ? (size_t)s * NC : (size_t)0) + (size_t)n0 * C, nd);
                if (last && p.h_out) copy_out_panels<CQ, NT>(HM, AT, 1, p.h_out + (size_t)n0 * C, nd);
            }
            MP_TICK(11)
        }
        __syncthreads();                                         // all reads of the tile (copy-outs) done before the next tile load
        MP_TICK(12)
    }
#undef MP_PREFETCH
#undef MP_TICK
    if (p.phase_clock && tid < 16) p.phase_clock[blockIdx.x * 32 + tid] = clk[tid];
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

constexpr int kMpThreads = 512;                                 // 384 / 768 / 1024 threads measured slower

template <int CQ, int H>
int mp_launch(const MpParams& p, bool save, cudaStream_t stream) {
    using G = MpGeom<CQ, H, kMpThreads>;
    auto go = [&](auto kernel) -> int {
        cudaError_t e = ensure_dyn_smem((const void*)kernel, (size_t)(G::SMEM));
        if (e != cudaSuccess) { set_error("glam_message_stack_fwd: cudaFuncSetAttribute(%d bytes): %s", G::SMEM, cudaGetErrorString(e)); return (int)e; }
        kernel<<<kNumSMs, kMpThreads, G::SMEM, stream>>>(p);
        return 0;
    };
    return save ? go(mp_fused_kernel<CQ, H, kMpThreads, true>) : go(mp_fused_kernel<CQ, H, kMpThreads, false>);
}

}  // namespace
}  // namespace glam

using namespace glam;

extern "C" int glam_graph_tile_caps(int* max_nodes, int* max_edges) {
    if (max_nodes) *max_nodes = kMpM;
    if (max_edges) *max_edges = kMpMaxEdges;
    return 0;
}

extern "C" size_t glam_graph_tiles_workspace_bytes(int64_t num_graphs) {
    const size_t chunks = (size_t)((num_graphs + kTileChunk - 1) / kTileChunk);
    return chunks * kTileChunk * sizeof(int4) + (chunks + 1) * sizeof(int32_t) + 16;
}

extern "C" int glam_build_graph_tiles(const int32_t* graph_ptr, int64_t num_graphs, const int32_t* dst_rowptr, const int32_t* dst_src,
                                      int64_t num_nodes, int64_t num_edges, int32_t* tiles, int32_t* meta, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && num_edges >= 0, "glam_build_graph_tiles: bad sizes");
    GLAM_REQUIRE(graph_ptr && dst_rowptr && tiles && meta, "glam_build_graph_tiles: null pointer");
    GLAM_REQUIRE(al16(tiles), "glam_build_graph_tiles: tiles must be 16-byte aligned");
    GLAM_REQUIRE(num_graphs < ((int64_t)1 << 30), "glam_build_graph_tiles: too many graphs");
    cudaStream_t stream = (cudaStream_t)stream_;
    GLAM_REQUIRE(workspace && workspace_bytes >= glam_graph_tiles_workspace_bytes(num_graphs) && al16(workspace),
                 "glam_build_graph_tiles: workspace too small or not 16-byte aligned");
    const int64_t chunks = (num_graphs + kTileChunk - 1) / kTileChunk;
    if (chunks == 0) return 0;                                   // meta[0] stays 0
    int4* scratch = reinterpret_cast<int4*>(workspace);
    int32_t* counts = reinterpret_cast<int32_t*>(scratch + chunks * kTileChunk);
    const unsigned grid = (unsigned)((chunks + kTileWarps - 1) / kTileWarps);
    graph_tiles_pack_kernel<<<(unsigned)chunks, 256, 0, stream>>>(graph_ptr, num_graphs, dst_rowptr, kMpM, kMpMaxEdges, scratch, counts, meta);
    GLAM_CHECK_LAUNCH();
    graph_tiles_scan_kernel<<<1, 1024, 0, stream>>>(counts, chunks, meta);
    GLAM_CHECK_LAUNCH();
    graph_tiles_emit_kernel<<<grid, kTileWarps * 32, 0, stream>>>(scratch, counts, chunks, meta, reinterpret_cast<int4*>(tiles));
    GLAM_CHECK_LAUNCH();
    if (num_graphs > 0 && num_edges > 0 && dst_src) {
        const int64_t warps = num_graphs;
        graph_tiles_check_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const int4*>(tiles), dst_src, meta, num_graphs);
        GLAM_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int glam_edge_types(const float* edge_attr_sorted, const int32_t* perm, int64_t num_edges, int edge_dim, uint8_t* etype,
                               int32_t* meta, void* stream_) {
    GLAM_REQUIRE(num_edges >= 0 && edge_dim > 0, "glam_edge_types: bad sizes");
    if (num_edges == 0) return 0;
    GLAM_REQUIRE(edge_attr_sorted && etype && meta, "glam_edge_types: null pointer");
    edge_types_kernel<<<(unsigned)((num_edges + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(edge_attr_sorted, perm, num_edges, edge_dim, etype, meta);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_message_stack_phase_clock(unsigned long long* cycles) {
    g_mp_phase_clock = cycles;
    return 0;
}

extern "C" int glam_message_stack_supported(int channels, int heads, int edge_dim) {
    if (g_math_mode_get() == 0) return 0;                        // exact-fp32 mode keeps the CUDA-core projections
    if (heads != 3 || edge_dim < 1 || edge_dim > kMpMaxDe) return 0;
    return (channels == 32 || channels == 36 || channels == 40) ? 1 : 0;
}

extern "C" int glam_message_stack_fwd(const float* x0, const float* h0, const float* x_raw, int raw_dim, const float* w_pre,
                                      const float* b_pre, int pre_act, float pre_act_param, const float* w_ext, int64_t ldw, const float* w_edge,
                                      const float* att_edge, const float* w_scale, const float* bias, const float* w_ih,
                                      const float* w_hh, const float* b_ih, const float* b_hh, const int32_t* tiles,
                                      const int32_t* tile_meta, const int32_t* dst_rowptr, const int32_t* dst_src,
                                      const uint8_t* etype, int64_t num_nodes, int64_t num_edges, int channels, int heads,
                                      int edge_dim, int steps, float negative_slope, int act, float act_param, int res,
                                      int conv_only, int keep_all, float* x_out, float* h_out, float* save_x, float* save_h,
                                      float* save_xpe, float* save_agg, float* save_alpha, float* save_m, float* save_rzn,
                                      float* save_gh, float* save_gt, float* save_mh, const int64_t* pn_batch, float pn_eps, void* stream_) {
    GLAM_REQUIRE(glam_message_stack_supported(channels, heads, edge_dim),
                 "glam_message_stack_fwd: unsupported (channels=%d heads=%d edge_dim=%d math mode %d); use the per-op calls", channels,
                 heads, edge_dim, g_math_mode_get());
    GLAM_REQUIRE(num_nodes >= 0 && num_edges >= 0 && steps >= 1, "glam_message_stack_fwd: bad sizes");
    if (num_nodes == 0) return 0;
    const bool save = save_xpe != nullptr;              // training: also write what MessageStackFn.backward / TripletConvFn.backward read
    GLAM_REQUIRE(!pn_batch || (!save && !conv_only && !h0), "glam_message_stack_fwd: PairNorm inside the kernel is an evaluation-mode path (no saves, no conv_only, h0 == NULL)");
    GLAM_REQUIRE((x0 || x_raw) && w_ext && w_edge && att_edge && w_scale && bias && tiles && tile_meta && dst_rowptr && (num_edges == 0 || (dst_src && etype)),
                 "glam_message_stack_fwd: null pointer");
    GLAM_REQUIRE(!x_raw || (w_pre && raw_dim >= 1 && raw_dim <= kMpMaxRaw), "glam_message_stack_fwd: input LinearBlock needs weights and raw_dim <= %d", kMpMaxRaw);
    GLAM_REQUIRE(conv_only ? steps == 1 : (w_ih && w_hh && b_ih && b_hh), "glam_message_stack_fwd: GRU weights missing / conv-only takes one step");
    GLAM_REQUIRE(save ? (save_agg && save_alpha && (conv_only ? x_out != nullptr : (save_x && save_h && (save_m || save_mh) && (save_gt || (save_rzn && save_gh)))))
                      : (x_out != nullptr),
                 "glam_message_stack_fwd: output pointers missing");
    const int HC = heads * channels, ld = (HC + 2 * heads + 3) / 4 * 4;
    GLAM_REQUIRE(ldw == ld, "glam_message_stack_fwd: w_ext pitch %lld, expected %d", (long long)ldw, ld);
    GLAM_REQUIRE(al16(x0) && al16(h0) && al16(w_ih) && al16(w_hh) && al16(x_out) && al16(h_out) && al16(save_x) && al16(save_h) &&
                 al16(save_xpe) && al16(save_agg) && al16(save_m) && al16(save_rzn) && al16(save_gh) && al16(save_gt) && al16(save_mh) && al16(tiles),
                 "glam_message_stack_fwd: pointers must be 16-byte aligned");
    GLAM_REQUIRE(num_nodes < ((int64_t)1 << 31) && num_edges < ((int64_t)1 << 31), "glam_message_stack_fwd: too large");
    MpParams p;
    p.x0 = x0; p.h0 = h0; p.x_raw = x_raw; p.raw_dim = raw_dim; p.w_pre = w_pre; p.b_pre = b_pre; p.pre_act = pre_act; p.pre_act_param = pre_act_param;
    p.w_ext = w_ext; p.ldw = (int)ldw; p.w_edge = w_edge; p.att_edge = att_edge; p.w_scale = w_scale; p.bias = bias;
    p.w_ih = w_ih; p.w_hh = w_hh; p.b_ih = b_ih; p.b_hh = b_hh; p.tiles = reinterpret_cast<const int4*>(tiles); p.meta = tile_meta;
    p.rowptr = dst_rowptr; p.src = dst_src; p.etype = etype; p.N = num_nodes; p.E = num_edges; p.De = edge_dim; p.steps = steps;
    p.act = act; p.res = res; p.conv_only = conv_only; p.keep_all = keep_all; p.slope = negative_slope; p.act_param = act_param;
    p.x_out = x_out; p.h_out = h_out; p.sX = save_x; p.sHH = save_h; p.sXPE = save_xpe; p.sAGG = save_agg; p.sALPHA = save_alpha;
    p.sM = save_m; p.sRZN = save_rzn; p.sGH = save_gh; p.sGT = save_gt; p.sMH = save_mh; p.pn_batch = pn_batch; p.pn_eps = pn_eps; p.phase_clock = g_mp_phase_clock;
    int rc = 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    switch (channels) {
        case 32: rc = mp_launch<8, 3>(p, save, stream); break;
        case 36: rc = mp_launch<9, 3>(p, save, stream); break;
        default: rc = mp_launch<10, 3>(p, save, stream); break;
    }
    if (rc) return rc;
    GLAM_CHECK_LAUNCH();
    return 0;
}
