// Set2Set readout rounds (PyG Set2Set(in_channels=C, processing_steps=S) @1.7.2, constructed at src_1gp/model.py:41).
//
// The reference runs S sequential rounds of { LSTM cell on q* [B,2C] -> q ; e = <x, q[batch]> ; segment softmax ;
// r = segment sum a*x ; q* = [q | r] }: ~15 tiny kernels per round forward, more backward.  Here a round is TWO
// launches: the gate pre-activations of all graphs are one tensor-core GEMM  U [B,3C] x [W_ih | W_hh]^T  (the weights
// are shared by every graph, so that is where the reuse is), and everything else of the round — gate non-linearities,
// cell update, attention logits, segment softmax, pooled read r and the next round's GEMM operand U' = [h | r | h] —
// is one kernel with one warp per graph working on a private shared-memory copy of the graph's node rows.  Backward
// mirrors it: one kernel per round (attention + cell backward, g_x accumulated in place, gate gradients G [B,4C]
// out), one GEMM G x [W_ih | W_hh] for the gradient of U, and ONE A^T B over the stacked rounds for the weights.
// All sums run in node order: deterministic.
//
// (A first version kept all S rounds of a graph inside one kernel with the LSTM mat-vec done per warp from shared
// memory; at 4096 graphs it was 0.2 ms forward / 0.4 ms backward, latency-bound on 15.5k serial MACs per graph per
// round with no reuse of the weights across graphs.  Splitting at the GEMM is ~5x faster.)
#include "common.cuh"

namespace glam {

constexpr int kS2SWarps = 8;
constexpr int kS2STileFloats = 2400;           // per-warp node tile: e.g. 64 nodes x (36+1); larger graphs stream from global

// dot product with four independent accumulators (breaks the FMA dependency chain)
__device__ __forceinline__ float dot_ilp(const float* __restrict__ a, const float* __restrict__ b, int n, float init) {
    float s0 = init, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        s0 = fmaf(a[k], b[k], s0); s1 = fmaf(a[k + 1], b[k + 1], s1);
        s2 = fmaf(a[k + 2], b[k + 2], s2); s3 = fmaf(a[k + 3], b[k + 3], s3);
    }
    for (; k < n; ++k) s0 = fmaf(a[k], b[k], s0);
    return (s0 + s1) + (s2 + s3);
}

// x block of a graph -> private tile [n][C+1] when it fits (the spare column is per-node scratch), else global rows
__device__ __forceinline__ const float* s2s_stage_x(const float* __restrict__ x, int64_t ldx, int C, float* tile, int n0, int n,
                                                    int lane, int* xld) {
    if (n * (C + 1) > kS2STileFloats) { *xld = (int)ldx; return x + (int64_t)n0 * ldx; }
    for (int idx = lane; idx < n * C; idx += 32) {
        const int i = idx / C, k = idx - i * C;
        tile[i * (C + 1) + k] = x[(int64_t)(n0 + i) * ldx + k];
    }
    __syncwarp();
    *xld = C + 1;
    return tile;
}

struct S2SRoundFwd {
    const float* x; int64_t ldx; const int32_t* gptr; int64_t B; int C;
    float* gates;            // [B,4C] in: pre-activations (bias included); out: activated i,f,g,o
    const float* c_prev;     // [B,C]
    float* c_new;            // [B,C]
    float* att;              // [N] attention weights of this round
    float* u_next;           // [B,3C] = [h | r | h]
    float* q_star;           // [B,2C] = [h | r] or null
};

__global__ void __launch_bounds__(kS2SWarps * 32)
set2set_round_fwd_kernel(const S2SRoundFwd p) {
    extern __shared__ float smem[];
    const int C = p.C;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* tile = smem + (size_t)wid * (kS2STileFloats + 2 * C);
    float* h = tile + kS2STileFloats;
    float* r = h + C;
    const int64_t g = (int64_t)blockIdx.x * kS2SWarps + wid;
    if (g >= p.B) return;
    const int n0 = p.gptr[g], n = p.gptr[g + 1] - n0;
    // LSTM cell (torch.nn.LSTM gate order i, f, g, o)
    float* gr = p.gates + g * 4 * C;
    for (int k = lane; k < C; k += 32) {
        const float i_ = sigmoidf_(gr[k]), f_ = sigmoidf_(gr[C + k]), g_ = tanhf(gr[2 * C + k]), o_ = sigmoidf_(gr[3 * C + k]);
        const float cn = f_ * p.c_prev[g * C + k] + i_ * g_;
        gr[k] = i_; gr[C + k] = f_; gr[2 * C + k] = g_; gr[3 * C + k] = o_;
        p.c_new[g * C + k] = cn;
        h[k] = o_ * tanhf(cn);
    }
    int xld;
    const float* xt = s2s_stage_x(p.x, p.ldx, C, tile, n0, n, lane, &xld);   // ends with __syncwarp
    __syncwarp();
    // attention: a = softmax_n <x[n], h>  (PyG softmax: exp(e - max) / (sum + 1e-16))
    float* att_s = p.att + n0;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) {
        const float s = dot_ilp(xt + (size_t)i * xld, h, C, 0.f);
        att_s[i] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float e = expf(att_s[i] - mx);
        att_s[i] = e;
        sum += e;
    }
    sum = warp_sum(sum) + 1e-16f;
    for (int i = lane; i < n; i += 32) att_s[i] = att_s[i] / sum;
    __syncwarp();
    for (int k = lane; k < C; k += 32) {
        float a0 = 0.f, a1 = 0.f;
        int i = 0;
        for (; i + 2 <= n; i += 2) {
            a0 = fmaf(att_s[i], xt[(size_t)i * xld + k], a0);
            a1 = fmaf(att_s[i + 1], xt[(size_t)(i + 1) * xld + k], a1);
        }
        if (i < n) a0 = fmaf(att_s[i], xt[(size_t)i * xld + k], a0);
        r[k] = a0 + a1;
    }
    __syncwarp();
    float* un = p.u_next + g * 3 * C;
    for (int k = lane; k < C; k += 32) {
        const float hv = h[k], rv = r[k];
        un[k] = hv; un[C + k] = rv; un[2 * C + k] = hv;
        if (p.q_star) { p.q_star[g * 2 * C + k] = hv; p.q_star[g * 2 * C + C + k] = rv; }
    }
}

struct S2SRoundBwd {
    const float* x; int64_t ldx; const int32_t* gptr; int64_t B; int C;
    const float* gates;      // [B,4C] activated gates of this round
    const float* c_prev;     // [B,C]
    const float* c_new;      // [B,C]
    const float* att;        // [N]
    const float* g_u; int64_t ldgu; int gu_cols;   // gradient of this round's output: [h | r] (2C) or [h | r | h] (3C)
    float* g_c;              // [B,C] in/out (gradient of c_new in, of c_prev out)
    float* g_x; int accumulate;
    float* G;                // [B,4C] gate pre-activation gradients out
};

__global__ void __launch_bounds__(kS2SWarps * 32)
set2set_round_bwd_kernel(const S2SRoundBwd p) {
    extern __shared__ float smem[];
    const int C = p.C;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* tile = smem + (size_t)wid * (kS2STileFloats + 3 * C);
    float* h = tile + kS2STileFloats;
    float* g_h = h + C;
    float* g_r = g_h + C;
    const int64_t g = (int64_t)blockIdx.x * kS2SWarps + wid;
    if (g >= p.B) return;
    const int n0 = p.gptr[g], n = p.gptr[g + 1] - n0;
    const float* gs = p.gates + g * 4 * C;
    const float* gu = p.g_u + g * p.ldgu;
    for (int k = lane; k < C; k += 32) {
        h[k] = gs[3 * C + k] * tanhf(p.c_new[g * C + k]);
        g_h[k] = gu[k] + (p.gu_cols == 3 * C ? gu[2 * C + k] : 0.f);
        g_r[k] = gu[C + k];
    }
    int xld;
    const float* xt = s2s_stage_x(p.x, p.ldx, C, tile, n0, n, lane, &xld);
    __syncwarp();
    const bool local = xld == C + 1;
    const float* att_s = p.att + n0;
    // r = sum a x, a = softmax(<x, h>):  g_a = <x, g_r>;  g_e = a (g_a - sum a g_a);  g_x += a g_r + g_e h;  g_h += sum g_e x
    float dot = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float ga = dot_ilp(xt + (size_t)i * xld, g_r, C, 0.f);
        dot = fmaf(att_s[i], ga, dot);
        if (local) tile[(size_t)i * xld + C] = ga;
    }
    dot = warp_sum(dot);
    __syncwarp();
    float* gx = p.g_x + (int64_t)n0 * C;
    if (local) {
        for (int i = lane; i < n; i += 32) tile[(size_t)i * xld + C] = att_s[i] * (tile[(size_t)i * xld + C] - dot);
        __syncwarp();
        // coalesced g_x update: consecutive lanes -> consecutive channels
        for (int idx = lane; idx < n * C; idx += 32) {
            const int i = idx / C, k = idx - i * C;
            const float v = fmaf(att_s[i], g_r[k], tile[(size_t)i * xld + C] * h[k]);
            gx[idx] = p.accumulate ? gx[idx] + v : v;
        }
        for (int k = lane; k < C; k += 32) {
            float a0 = 0.f, a1 = 0.f;
            int i = 0;
            for (; i + 2 <= n; i += 2) {
                a0 = fmaf(tile[(size_t)i * xld + C], xt[(size_t)i * xld + k], a0);
                a1 = fmaf(tile[(size_t)(i + 1) * xld + C], xt[(size_t)(i + 1) * xld + k], a1);
            }
            if (i < n) a0 = fmaf(tile[(size_t)i * xld + C], xt[(size_t)i * xld + k], a0);
            g_h[k] += a0 + a1;
        }
    } else {
        // large graph: stream rows from global; g_e recomputed per row, g_h partials reduced across lanes through smem
        float* acc = tile;                       // [C] accumulators for sum g_e x (tile is free on this path)
        for (int k = lane; k < C; k += 32) acc[k] = 0.f;
        __syncwarp();
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            float ge = 0.f;
            if (i < n) {
                const float ga = dot_ilp(xt + (size_t)i * xld, g_r, C, 0.f);
                ge = att_s[i] * (ga - dot);
            }
            const int cnt = min(32, n - i0);
            for (int j = 0; j < cnt; ++j) {
                const float gej = __shfl_sync(0xffffffffu, ge, j);
                const float aj = att_s[i0 + j];
                const float* row = xt + (size_t)(i0 + j) * xld;
                float* gxr = gx + (size_t)(i0 + j) * C;
                for (int k = lane; k < C; k += 32) {
                    const float v = fmaf(aj, g_r[k], gej * h[k]);
                    gxr[k] = p.accumulate ? gxr[k] + v : v;
                    acc[k] = fmaf(gej, row[k], acc[k]);
                }
            }
        }
        __syncwarp();
        for (int k = lane; k < C; k += 32) g_h[k] += acc[k];
    }
    __syncwarp();
    // LSTM cell backward
    float* Gr = p.G + g * 4 * C;
    for (int k = lane; k < C; k += 32) {
        const float i_ = gs[k], f_ = gs[C + k], gg = gs[2 * C + k], o_ = gs[3 * C + k];
        const float tc = tanhf(p.c_new[g * C + k]);
        const float gh = g_h[k];
        const float gc = p.g_c[g * C + k] + gh * o_ * (1.f - tc * tc);
        Gr[k] = gc * gg * i_ * (1.f - i_);
        Gr[C + k] = gc * p.c_prev[g * C + k] * f_ * (1.f - f_);
        Gr[2 * C + k] = gc * i_ * (1.f - gg * gg);
        Gr[3 * C + k] = gh * tc * o_ * (1.f - o_);
        p.g_c[g * C + k] = gc * f_;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Row-per-lane variants (C % 4 == 0, C <= 64, 16-byte aligned rows): a lane keeps its node's row in registers
// (C/4 float4 loads), so a 32-node chunk of the graph is read once from global and never staged element-wise; the
// cross-lane sums (r = sum_n a_n x_n, g_h += sum_n g_e,n x_n) go through a [32][C+1] shared tile read column-wise.
// Graphs larger than 32 nodes loop over chunks (rows re-read from L1/L2 in the second pass).
// ---------------------------------------------------------------------------------------------------------------
template <int C4>
__device__ __forceinline__ void load_row(float4 (&xr)[C4], const float* __restrict__ row) {
#pragma unroll
    for (int q = 0; q < C4; ++q) xr[q] = __ldg(reinterpret_cast<const float4*>(row) + q);
}
template <int C4>
__device__ __forceinline__ float dot_row(const float4 (&xr)[C4], const float* __restrict__ v_smem) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int q = 0; q < C4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(v_smem + 4 * q);
        s0 = fmaf(xr[q].x, v.x, s0); s1 = fmaf(xr[q].y, v.y, s1); s2 = fmaf(xr[q].z, v.z, s2); s3 = fmaf(xr[q].w, v.w, s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// tile[lane][:] = w * xr  (rows of a chunk that do not exist are written as zeros)
template <int C4>
__device__ __forceinline__ void scatter_scaled_row(float* tile, int lane, const float4 (&xr)[C4], float w) {
    float* tr = tile + lane * (4 * C4 + 1);
#pragma unroll
    for (int q = 0; q < C4; ++q) {
        tr[4 * q] = w * xr[q].x; tr[4 * q + 1] = w * xr[q].y; tr[4 * q + 2] = w * xr[q].z; tr[4 * q + 3] = w * xr[q].w;
    }
}
// column sums of the tile: lane owns channels lane and lane + 32
template <int C4>
__device__ __forceinline__ void column_sums(const float* tile, int lane, int rows, float& acc0, float& acc1) {
    constexpr int C = 4 * C4, ld = C + 1;
    if (lane < C) {
        float a0 = 0.f, a1 = 0.f;
        int i = 0;
        for (; i + 2 <= rows; i += 2) { a0 += tile[i * ld + lane]; a1 += tile[(i + 1) * ld + lane]; }
        if (i < rows) a0 += tile[i * ld + lane];
        acc0 += a0 + a1;
    }
    if (C > 32 && lane + 32 < C) {
        float a0 = 0.f, a1 = 0.f;
        int i = 0;
        for (; i + 2 <= rows; i += 2) { a0 += tile[i * ld + lane + 32]; a1 += tile[(i + 1) * ld + lane + 32]; }
        if (i < rows) a0 += tile[i * ld + lane + 32];
        acc1 += a0 + a1;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Chunk-per-lane path for graphs of <= 32 nodes (every molecule): a lane owns one 16-byte chunk of the row, so a warp
// load covers RP = 32 / (C/4) consecutive rows of the graph as ONE contiguous span (4 lines instead of 32 per request),
// the rows stay in registers for both passes, row dots are segmented shuffles over the C/4 lanes of a row, and the
// per-graph sums (r, g_h) are register accumulations reduced over the RP row groups — no shared-memory tile at all.
// ---------------------------------------------------------------------------------------------------------------
// SFU forms for this path (ex2.approx / rcp.approx: relative error ~2^-22, the same the one-launch message kernels use):
// the accurate expf / tanhf / divisions were 40 % of the warp's instructions, and the kernel is issue-latency bound
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, fast_sigmoid(2.f * x), -1.f); }
template <int CH>
__device__ __forceinline__ float seg_sum(float v, int q, int seg_base) {
    constexpr int P2 = CH >= 16 ? 16 : CH >= 8 ? 8 : CH >= 4 ? 4 : CH >= 2 ? 2 : 1;
    if (CH > P2) { const float t = __shfl_down_sync(0xffffffffu, v, P2); if (q + P2 < CH) v += t; }
#pragma unroll
    for (int off = P2 / 2; off >= 1; off >>= 1) { const float t = __shfl_down_sync(0xffffffffu, v, off); if (q < off) v += t; }
    return __shfl_sync(0xffffffffu, v, seg_base);
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// sum of a per-row-group float4 over the RP groups, valid in the lanes of group 0 (fixed order)
template <int CH, int RP>
__device__ __forceinline__ float4 group_sum4(float4 v, int rg) {
#pragma unroll
    for (int r = 1; r < RP; ++r) {
        const float tx = __shfl_down_sync(0xffffffffu, v.x, r * CH), ty = __shfl_down_sync(0xffffffffu, v.y, r * CH);
        const float tz = __shfl_down_sync(0xffffffffu, v.z, r * CH), tw = __shfl_down_sync(0xffffffffu, v.w, r * CH);
        if (rg == 0) { v.x += tx; v.y += ty; v.z += tz; v.w += tw; }
    }
    return v;
}

template <int C4>
__device__ __forceinline__ void set2set_fwd_small(const S2SRoundFwd& p, int64_t g, int n0, int n, int lane, float* h) {
    constexpr int C = 4 * C4, CH = C4, RP = 32 / CH, MAXP = (32 + RP - 1) / RP;
    const int q = lane % CH, rg = lane / CH, seg_base = rg * CH;
    const bool live = rg < RP;
    float4 xq[MAXP];
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        const int i = pp * RP + rg;
        xq[pp] = (live && i < n) ? __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)(n0 + i) * p.ldx) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* gr = p.gates + g * 4 * C;
    float hv[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            const float i_ = fast_sigmoid(gr[k]), f_ = fast_sigmoid(gr[C + k]), g_ = fast_tanh(gr[2 * C + k]), o_ = fast_sigmoid(gr[3 * C + k]);
            const float cn = f_ * p.c_prev[g * C + k] + i_ * g_;
            gr[k] = i_; gr[C + k] = f_; gr[2 * C + k] = g_; gr[3 * C + k] = o_;
            p.c_new[g * C + k] = cn;
            hv[j] = o_ * fast_tanh(cn);
            h[k] = hv[j];
        }
    }
    __syncwarp();
    const float4 h4 = live ? *reinterpret_cast<const float4*>(h + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float e[MAXP], mx = -INFINITY;
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        e[pp] = 0.f;
        if (pp * RP < n) {                                   // warp-uniform: passes beyond the graph's last row are skipped
            e[pp] = seg_sum<CH>(dot4(xq[pp], h4), q, seg_base);
            if (live && pp * RP + rg < n) mx = fmaxf(mx, e[pp]);
        }
    }
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < RP; ++r) m = fmaxf(m, __shfl_sync(0xffffffffu, mx, r * CH));
    float sl = 0.f;
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        e[pp] = (live && pp * RP + rg < n) ? __expf(e[pp] - m) : 0.f;
        sl += e[pp];
    }
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < RP; ++r) sum += __shfl_sync(0xffffffffu, sl, r * CH);
    sum += 1e-16f;
    const float inv_sum = __fdividef(1.f, sum);
    float* att_s = p.att + n0;
    float4 racc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        if (pp * RP >= n) break;
        const float a = e[pp] * inv_sum;
        const int i = pp * RP + rg;
        if (live && q == 0 && i < n) att_s[i] = a;
        racc.x = fmaf(a, xq[pp].x, racc.x); racc.y = fmaf(a, xq[pp].y, racc.y);
        racc.z = fmaf(a, xq[pp].z, racc.z); racc.w = fmaf(a, xq[pp].w, racc.w);
    }
    racc = group_sum4<CH, RP>(racc, rg);
    float* un = p.u_next + g * 3 * C;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            un[k] = hv[j]; un[2 * C + k] = hv[j];
            if (p.q_star) p.q_star[g * 2 * C + k] = hv[j];
        }
    }
    if (rg == 0) {
        const float rv[4] = {racc.x, racc.y, racc.z, racc.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            un[C + 4 * q + u] = rv[u];
            if (p.q_star) p.q_star[g * 2 * C + C + 4 * q + u] = rv[u];
        }
    }
}

template <int C4>
__device__ __forceinline__ void set2set_bwd_small(const S2SRoundBwd& p, int64_t g, int n0, int n, int lane, float* h, float* g_r) {
    constexpr int C = 4 * C4, CH = C4, RP = 32 / CH, MAXP = (32 + RP - 1) / RP;
    const int q = lane % CH, rg = lane / CH, seg_base = rg * CH;
    const bool live = rg < RP;
    float4 xq[MAXP];
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        const int i = pp * RP + rg;
        xq[pp] = (live && i < n) ? __ldg(reinterpret_cast<const float4*>(p.x + (int64_t)(n0 + i) * p.ldx) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float a_mine = lane < n ? p.att[n0 + lane] : 0.f;
    const float* gs = p.gates + g * 4 * C;
    const float* gu = p.g_u + g * p.ldgu;
    float gh[2] = {0.f, 0.f}, tc[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            tc[j] = fast_tanh(p.c_new[g * C + k]);
            h[k] = gs[3 * C + k] * tc[j];
            gh[j] = gu[k] + (p.gu_cols == 3 * C ? gu[2 * C + k] : 0.f);
            g_r[k] = gu[C + k];
        }
    }
    __syncwarp();
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 h4 = live ? *reinterpret_cast<const float4*>(h + 4 * q) : zero4;
    const float4 r4 = live ? *reinterpret_cast<const float4*>(g_r + 4 * q) : zero4;
    float ga[MAXP], a[MAXP], part = 0.f;
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        ga[pp] = 0.f; a[pp] = 0.f;
        if (pp * RP < n) {                                   // warp-uniform
            ga[pp] = seg_sum<CH>(dot4(xq[pp], r4), q, seg_base);
            const int i = pp * RP + rg;
            a[pp] = __shfl_sync(0xffffffffu, a_mine, i < 32 ? i : 31);
            if (!(live && i < n)) a[pp] = 0.f;
            part = fmaf(a[pp], ga[pp], part);
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int r = 0; r < RP; ++r) dot += __shfl_sync(0xffffffffu, part, r * CH);
    float4 hacc = zero4;
#pragma unroll
    for (int pp = 0; pp < MAXP; ++pp) {
        if (pp * RP >= n) break;
        const float ge = a[pp] * (ga[pp] - dot);
        const int i = pp * RP + rg;
        if (live && i < n) {
            float4 v = make_float4(fmaf(a[pp], r4.x, ge * h4.x), fmaf(a[pp], r4.y, ge * h4.y), fmaf(a[pp], r4.z, ge * h4.z), fmaf(a[pp], r4.w, ge * h4.w));
            float4* dst = reinterpret_cast<float4*>(p.g_x + (int64_t)(n0 + i) * C) + q;
            // earlier rounds' g_x: a 16-byte reduction at L2, no read back (one adder per element and launch: deterministic)
            if (p.accumulate) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            else *dst = v;
        }
        hacc.x = fmaf(ge, xq[pp].x, hacc.x); hacc.y = fmaf(ge, xq[pp].y, hacc.y);
        hacc.z = fmaf(ge, xq[pp].z, hacc.z); hacc.w = fmaf(ge, xq[pp].w, hacc.w);
    }
    hacc = group_sum4<CH, RP>(hacc, rg);
    __syncwarp();                                          // every lane holds its g_r chunk in registers: the slot is free
    if (rg == 0) *reinterpret_cast<float4*>(g_r + 4 * q) = hacc;
    __syncwarp();
    float* Gr = p.G + g * 4 * C;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            const float ghk = gh[j] + g_r[k];
            const float i_ = gs[k], f_ = gs[C + k], gg = gs[2 * C + k], o_ = gs[3 * C + k];
            const float gc = p.g_c[g * C + k] + ghk * o_ * (1.f - tc[j] * tc[j]);
            Gr[k] = gc * gg * i_ * (1.f - i_);
            Gr[C + k] = gc * p.c_prev[g * C + k] * f_ * (1.f - f_);
            Gr[2 * C + k] = gc * i_ * (1.f - gg * gg);
            Gr[3 * C + k] = ghk * tc[j] * o_ * (1.f - o_);
            p.g_c[g * C + k] = gc * f_;
        }
    }
}

template <int C4>
__global__ void __launch_bounds__(kS2SWarps * 32, C4 <= 9 ? 3 : 1)
set2set_round_fwd_rows_kernel(const S2SRoundFwd p) {
    constexpr int C = 4 * C4;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* tile = smem + (size_t)wid * (32 * (C + 1) + C + 4);
    float* h = tile + 32 * (C + 1) + ((32 * (C + 1)) & 3 ? 4 - ((32 * (C + 1)) & 3) : 0);   // 16-byte aligned
    const int64_t g = (int64_t)blockIdx.x * kS2SWarps + wid;
    if (g >= p.B) return;
    const int n0 = p.gptr[g], n = p.gptr[g + 1] - n0;
    if (n <= 32) { set2set_fwd_small<C4>(p, g, n0, n, lane, h); return; }
    const int nch = (n + 31) >> 5;
    float4 xr[C4];
    if (lane < n) load_row<C4>(xr, p.x + (int64_t)(n0 + lane) * p.ldx);          // in flight while the cell is computed
    float* gr = p.gates + g * 4 * C;
    float hv[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            const float i_ = sigmoidf_(gr[k]), f_ = sigmoidf_(gr[C + k]), g_ = tanhf(gr[2 * C + k]), o_ = sigmoidf_(gr[3 * C + k]);
            const float cn = f_ * p.c_prev[g * C + k] + i_ * g_;
            gr[k] = i_; gr[C + k] = f_; gr[2 * C + k] = g_; gr[3 * C + k] = o_;
            p.c_new[g * C + k] = cn;
            hv[j] = o_ * tanhf(cn);
            h[k] = hv[j];
        }
    }
    __syncwarp();
    float* att_s = p.att + n0;
    // pass 1: logits, max
    float e = -INFINITY, mx = -INFINITY;
    for (int ch = 0; ch < nch; ++ch) {
        const int i = ch * 32 + lane;
        if (i < n) {
            if (ch > 0) load_row<C4>(xr, p.x + (int64_t)(n0 + i) * p.ldx);
            e = dot_row<C4>(xr, h);
            if (nch > 1) att_s[i] = e;
            mx = fmaxf(mx, e);
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    if (nch == 1) { e = lane < n ? expf(e - mx) : 0.f; sum = e; }
    else {
        __syncwarp();
        for (int i = lane; i < n; i += 32) { const float ex = expf(att_s[i] - mx); att_s[i] = ex; sum += ex; }
    }
    sum = warp_sum(sum) + 1e-16f;
    // pass 2: a = e / sum, r = sum a x
    float r0 = 0.f, r1 = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
        const int i = ch * 32 + lane;
        float a = 0.f;
        if (i < n) {
            if (nch > 1) { load_row<C4>(xr, p.x + (int64_t)(n0 + i) * p.ldx); a = att_s[i] / sum; }
            else a = e / sum;
            att_s[i] = a;
            scatter_scaled_row<C4>(tile, lane, xr, a);
        }
        __syncwarp();
        column_sums<C4>(tile, lane, min(32, n - ch * 32), r0, r1);
        __syncwarp();
    }
    float* un = p.u_next + g * 3 * C;
    const float rv[2] = {r0, r1};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            un[k] = hv[j]; un[C + k] = rv[j]; un[2 * C + k] = hv[j];
            if (p.q_star) { p.q_star[g * 2 * C + k] = hv[j]; p.q_star[g * 2 * C + C + k] = rv[j]; }
        }
    }
}

template <int C4>
__global__ void __launch_bounds__(kS2SWarps * 32, C4 <= 9 ? 2 : 1)
set2set_round_bwd_rows_kernel(const S2SRoundBwd p) {
    constexpr int C = 4 * C4;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int kTile = (32 * (C + 1) + 3) / 4 * 4;
    float* tile = smem + (size_t)wid * (kTile + 2 * C);
    float* h = tile + kTile;
    float* g_r = h + C;
    const int64_t g = (int64_t)blockIdx.x * kS2SWarps + wid;
    if (g >= p.B) return;
    const int n0 = p.gptr[g], n = p.gptr[g + 1] - n0;
    if (n <= 32) { set2set_bwd_small<C4>(p, g, n0, n, lane, h, g_r); return; }
    const int nch = (n + 31) >> 5;
    float4 xr[C4], old[C4];
    if (lane < n) load_row<C4>(xr, p.x + (int64_t)(n0 + lane) * p.ldx);
    if (p.accumulate && lane < n) {              // previous rounds' g_x row: in flight while the logits are computed
#pragma unroll
        for (int q = 0; q < C4; ++q) old[q] = *(reinterpret_cast<const float4*>(p.g_x + (int64_t)(n0 + lane) * C) + q);
    }
    const float* gs = p.gates + g * 4 * C;
    const float* gu = p.g_u + g * p.ldgu;
    float gh[2] = {0.f, 0.f}, tc[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            tc[j] = tanhf(p.c_new[g * C + k]);
            h[k] = gs[3 * C + k] * tc[j];
            gh[j] = gu[k] + (p.gu_cols == 3 * C ? gu[2 * C + k] : 0.f);
            g_r[k] = gu[C + k];
        }
    }
    __syncwarp();
    const float* att_s = p.att + n0;
    // pass A: dot = sum_n a_n <x_n, g_r>
    float ga = 0.f, a = 0.f, part = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
        const int i = ch * 32 + lane;
        if (i < n) {
            if (ch > 0) load_row<C4>(xr, p.x + (int64_t)(n0 + i) * p.ldx);
            ga = dot_row<C4>(xr, g_r);
            a = att_s[i];
            part = fmaf(a, ga, part);
        }
    }
    const float dot = warp_sum(part);
    // pass B: g_e = a (g_a - dot);  g_x_n (+)= a g_r + g_e h;  g_h += sum_n g_e x_n
    for (int ch = 0; ch < nch; ++ch) {
        const int i = ch * 32 + lane;
        float ge = 0.f;
        if (i < n) {
            if (nch > 1) { load_row<C4>(xr, p.x + (int64_t)(n0 + i) * p.ldx); ga = dot_row<C4>(xr, g_r); a = att_s[i]; }
            ge = a * (ga - dot);
            float4* gxr = reinterpret_cast<float4*>(p.g_x + (int64_t)(n0 + i) * C);
#pragma unroll
            for (int q = 0; q < C4; ++q) {
                const float4 r4 = *reinterpret_cast<const float4*>(g_r + 4 * q), h4 = *reinterpret_cast<const float4*>(h + 4 * q);
                float4 v = make_float4(fmaf(a, r4.x, ge * h4.x), fmaf(a, r4.y, ge * h4.y), fmaf(a, r4.z, ge * h4.z), fmaf(a, r4.w, ge * h4.w));
                if (p.accumulate) { const float4 o = ch == 0 ? old[q] : gxr[q]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                gxr[q] = v;
            }
            scatter_scaled_row<C4>(tile, lane, xr, ge);
        }
        __syncwarp();
        column_sums<C4>(tile, lane, min(32, n - ch * 32), gh[0], gh[1]);
        __syncwarp();
    }
    // LSTM cell backward
    float* Gr = p.G + g * 4 * C;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int k = lane + 32 * j;
        if (k < C) {
            const float i_ = gs[k], f_ = gs[C + k], gg = gs[2 * C + k], o_ = gs[3 * C + k];
            const float gc = p.g_c[g * C + k] + gh[j] * o_ * (1.f - tc[j] * tc[j]);
            Gr[k] = gc * gg * i_ * (1.f - i_);
            Gr[C + k] = gc * p.c_prev[g * C + k] * f_ * (1.f - f_);
            Gr[2 * C + k] = gc * i_ * (1.f - gg * gg);
            Gr[3 * C + k] = gh[j] * tc[j] * o_ * (1.f - o_);
            p.g_c[g * C + k] = gc * f_;
        }
    }
}

static bool s2s_rows_ok(const float* x, int64_t ldx, int C) {
    return (C == 32 || C == 36 || C == 48 || C == 64) && (ldx & 3) == 0 && ((uintptr_t)x & 15) == 0;
}
template <int C4> static size_t s2s_rows_smem(bool bwd) {
    constexpr int C = 4 * C4;
    return sizeof(float) * kS2SWarps * (size_t)(bwd ? (32 * (C + 1) + 3) / 4 * 4 + 2 * C : 32 * (C + 1) + C + 4);
}
#define S2S_ROWS_DISPATCH(C, ...)                                \
    switch (C) {                                                  \
        case 32: { constexpr int C4 = 8; __VA_ARGS__; } break;    \
        case 36: { constexpr int C4 = 9; __VA_ARGS__; } break;    \
        case 48: { constexpr int C4 = 12; __VA_ARGS__; } break;   \
        default: { constexpr int C4 = 16; __VA_ARGS__; } break;   \
    }

}  // namespace glam

using namespace glam;

extern "C" int glam_set2set_round_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels,
                                      float* gates, const float* c_prev, float* c_new, float* att, float* u_next,
                                      float* q_star, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && channels > 0 && channels <= 512 && ldx >= channels, "glam_set2set_round_fwd: bad shape");
    if (num_graphs == 0) return 0;
    GLAM_REQUIRE(x && graph_ptr && gates && c_prev && c_new && att && u_next, "glam_set2set_round_fwd: null pointer");
    S2SRoundFwd p{x, ldx, graph_ptr, num_graphs, channels, gates, c_prev, c_new, att, u_next, q_star};
    if (s2s_rows_ok(x, ldx, channels)) {
        const unsigned grid_r = (unsigned)((num_graphs + kS2SWarps - 1) / kS2SWarps);
        S2S_ROWS_DISPATCH(channels, {
            const size_t sm = s2s_rows_smem<C4>(false);
            ensure_dyn_smem((const void*)set2set_round_fwd_rows_kernel<C4>, (size_t)((int)sm));
            set2set_round_fwd_rows_kernel<C4><<<grid_r, kS2SWarps * 32, sm, (cudaStream_t)stream_>>>(p);
        })
        GLAM_CHECK_LAUNCH();
        return 0;
    }
    const size_t smem = sizeof(float) * kS2SWarps * (size_t)(kS2STileFloats + 2 * channels);
    ensure_dyn_smem((const void*)set2set_round_fwd_kernel, (size_t)((int)smem));
    const unsigned grid = (unsigned)((num_graphs + kS2SWarps - 1) / kS2SWarps);
    set2set_round_fwd_kernel<<<grid, kS2SWarps * 32, smem, (cudaStream_t)stream_>>>(p);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_set2set_round_bwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int64_t num_graphs, int channels,
                                      const float* gates, const float* c_prev, const float* c_new, const float* att,
                                      const float* g_u, int64_t ldgu, int gu_cols, float* g_c, float* g_x, int accumulate,
                                      float* G, void* stream_) {
    GLAM_REQUIRE(num_graphs >= 0 && channels > 0 && channels <= 512 && ldx >= channels, "glam_set2set_round_bwd: bad shape");
    GLAM_REQUIRE(gu_cols == 2 * channels || gu_cols == 3 * channels, "glam_set2set_round_bwd: g_u must have 2C or 3C columns");
    if (num_graphs == 0) return 0;
    GLAM_REQUIRE(x && graph_ptr && gates && c_prev && c_new && att && g_u && g_c && g_x && G && ldgu >= gu_cols,
                 "glam_set2set_round_bwd: null pointer");
    S2SRoundBwd p{x, ldx, graph_ptr, num_graphs, channels, gates, c_prev, c_new, att, g_u, ldgu, gu_cols, g_c, g_x, accumulate, G};
    if (s2s_rows_ok(x, ldx, channels) && ((uintptr_t)g_x & 15) == 0) {
        const unsigned grid_r = (unsigned)((num_graphs + kS2SWarps - 1) / kS2SWarps);
        S2S_ROWS_DISPATCH(channels, {
            const size_t sm = s2s_rows_smem<C4>(true);
            ensure_dyn_smem((const void*)set2set_round_bwd_rows_kernel<C4>, (size_t)((int)sm));
            set2set_round_bwd_rows_kernel<C4><<<grid_r, kS2SWarps * 32, sm, (cudaStream_t)stream_>>>(p);
        })
        GLAM_CHECK_LAUNCH();
        return 0;
    }
    const size_t smem = sizeof(float) * kS2SWarps * (size_t)(kS2STileFloats + 3 * channels);
    ensure_dyn_smem((const void*)set2set_round_bwd_kernel, (size_t)((int)smem));
    const unsigned grid = (unsigned)((num_graphs + kS2SWarps - 1) / kS2SWarps);
    set2set_round_bwd_kernel<<<grid, kS2SWarps * 32, smem, (cudaStream_t)stream_>>>(p);
    GLAM_CHECK_LAUNCH();
    return 0;
}
