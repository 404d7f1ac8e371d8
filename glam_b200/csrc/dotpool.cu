// Cross-graph interaction pool dot_and_global_pool2 (see include/glam_b200.h (6)).
//
// The reference loops over pairs on the host with 4 .item() syncs per pair (src_2gi_ddi/layer.py:275-282);
// here one CTA owns one pair, tiles S = Xa Xb^T through shared memory 32x32 at a time and keeps only the
// running max/argmax.  mean(S) needs no product at all: <sum_a Xa, sum_b Xb> / (na*nb).
// Backward of the full-reduction max routes to the first arg-max in row-major order (ties have measure zero
// for real activations; the oracle test pins this choice).
#include "common.cuh"
#include "tc_common.cuh"
#include "mp_common.cuh"

namespace glam {

int g_math_mode_get();

constexpr int kPairThreads = 128;
constexpr int kPairTile = 32;

__global__ void __launch_bounds__(kPairThreads)
pair_dot_pool_fwd_kernel(const float* __restrict__ xa, const float* __restrict__ xb, const int32_t* __restrict__ ptr_a,
                         const int32_t* __restrict__ ptr_b, const int32_t* __restrict__ idx_b, int C, float* __restrict__ out,
                         int32_t* __restrict__ argmax, float* __restrict__ sum_a, float* __restrict__ sum_b) {
    extern __shared__ float smem[];
    const int ld = C | 1;                               // odd stride: conflict-free row reads
    float* As = smem;                                   // [32][ld]
    float* Bs = smem + kPairTile * ld;                  // [32][ld]
    __shared__ float red_v[kPairThreads / 32];
    __shared__ long long red_i[kPairThreads / 32];
    __shared__ float red_dot[kPairThreads / 32];
    const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int gb = idx_b ? idx_b[g] : g;                // shared second operand: pair g reads graph idx_b[g] of the b side
    const int a0 = ptr_a[g], a1 = ptr_a[g + 1], b0 = ptr_b[gb], b1 = ptr_b[gb + 1];
    const int na = a1 - a0, nb = b1 - b0;

    // column sums (fixed row order) and the mean.  Pairs of two small graphs (drug-drug: one tile each) take the sums from the
    // staged tiles below instead of a second, latency-bound pass over global memory (36 threads x 50 dependent-rate loads per CTA
    // were most of the kernel's 68 us at 4096 pairs); same values in the same order either way.
    const bool small = na > 0 && nb > 0 && na <= kPairTile && nb <= kPairTile;
    float dotp = 0.f;
    if (!small) {
        for (int k = t; k < C; k += kPairThreads) {
            float sa = 0.f, sb = 0.f;
            for (int a = a0; a < a1; ++a) sa += xa[(int64_t)a * C + k];
            for (int b = b0; b < b1; ++b) sb += xb[(int64_t)b * C + k];
            sum_a[(int64_t)g * C + k] = sa;
            sum_b[(int64_t)g * C + k] = sb;
            dotp = fmaf(sa, sb, dotp);
        }
    }

    float best = -INFINITY;
    long long best_i = 0x7fffffffffffffffLL;
    const int bi = t & 31;
    for (int at = 0; at < na; at += kPairTile) {
        __syncthreads();
        {                                               // (row, column) stepped along with idx: no division per element
            int r = t / C, k = t - r * C;
            const int dr = kPairThreads / C, dk = kPairThreads - dr * C;
            while (r < kPairTile) {
                As[r * ld + k] = (at + r < na) ? xa[(int64_t)(a0 + at + r) * C + k] : 0.f;
                r += dr; k += dk;
                if (k >= C) { k -= C; ++r; }
            }
        }
        for (int bt = 0; bt < nb; bt += kPairTile) {
            __syncthreads();
            {
                int r = t / C, k = t - r * C;
                const int dr = kPairThreads / C, dk = kPairThreads - dr * C;
                while (r < kPairTile) {
                    Bs[r * ld + k] = (bt + r < nb) ? xb[(int64_t)(b0 + bt + r) * C + k] : 0.f;
                    r += dr; k += dk;
                    if (k >= C) { k -= C; ++r; }
                }
            }
            __syncthreads();
            if (small) {                                // the only tile pair: column sums out of shared memory, rows in order
                for (int k = t; k < C; k += kPairThreads) {
                    float sa = 0.f, sb = 0.f;
                    for (int a = 0; a < na; ++a) sa += As[a * ld + k];
                    for (int b = 0; b < nb; ++b) sb += Bs[b * ld + k];
                    sum_a[(int64_t)g * C + k] = sa;
                    sum_b[(int64_t)g * C + k] = sb;
                    dotp = fmaf(sa, sb, dotp);
                }
            }
            float acc[kPairTile / 4];
#pragma unroll
            for (int i = 0; i < kPairTile / 4; ++i) acc[i] = 0.f;
            for (int k = 0; k < C; ++k) {
                const float bv = Bs[bi * ld + k];
#pragma unroll
                for (int i = 0; i < kPairTile / 4; ++i) acc[i] = fmaf(As[(wid + 4 * i) * ld + k], bv, acc[i]);
            }
#pragma unroll
            for (int i = 0; i < kPairTile / 4; ++i) {
                const int a = at + wid + 4 * i, b = bt + bi;
                if (a < na && b < nb) {
                    const long long lin = (long long)a * nb + b;
                    if (acc[i] > best || (acc[i] == best && lin < best_i)) { best = acc[i]; best_i = lin; }
                }
            }
        }
    }
    dotp = warp_sum(dotp);
    if (lane == 0) red_dot[wid] = dotp;
    // block arg-max (value desc, linear index asc)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = best_i; }
    __syncthreads();
    if (t == 0) {
        float d = 0.f;
        for (int w = 0; w < kPairThreads / 32; ++w) {
            d += red_dot[w];
            if (red_v[w] > best || (red_v[w] == best && red_i[w] < best_i)) { best = red_v[w]; best_i = red_i[w]; }
        }
        if (na > 0 && nb > 0) {
            out[2 * (int64_t)g] = best;
            out[2 * (int64_t)g + 1] = d / ((float)na * (float)nb);
            argmax[2 * (int64_t)g] = a0 + (int)(best_i / nb);
            argmax[2 * (int64_t)g + 1] = b0 + (int)(best_i % nb);
        } else {
            out[2 * (int64_t)g] = 0.f;
            out[2 * (int64_t)g + 1] = 0.f;
            argmax[2 * (int64_t)g] = -1;
            argmax[2 * (int64_t)g + 1] = -1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Small pairs (drug-drug: 25 x 25 atoms): one WARP per pair, eight pairs per CTA, no block barrier.  Lane a loads row a of
// the first graph and stores it transposed ([k][a], pitch 36: a broadcast LDS.128 then delivers four rows' values of one
// channel), lane b keeps row b of the second graph in REGISTERS and owns column b of S = Xa Xb^T: 32 accumulators, the
// channel loop fully unrolled.  The CTA-per-pair kernel above spent its time in barriers and in the start-up latency of
// 4096 CTAs (59 us per 4096 pairs, 0.06 of HBM); same values (k-ascending fmaf chains), same arg-max rule.  Column sums come
// out of the transposed tiles in row order.  Graphs over 32 rows loop over 32-row tiles (correct, not the intended use).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPairWarps = 8;
constexpr int kPairTP = 36;                              // pitch of the transposed tiles: 16-byte aligned rows of 32 (+4 pad)
template <int C4>
__global__ void __launch_bounds__(kPairWarps * 32)
pair_dot_pool_fwd_warp_kernel(const float* __restrict__ xa, const float* __restrict__ xb, const int32_t* __restrict__ ptr_a,
                              const int32_t* __restrict__ ptr_b, const int32_t* __restrict__ idx_b, int64_t num_pairs,
                              float* __restrict__ out, int32_t* __restrict__ argmax, float* __restrict__ sum_a,
                              float* __restrict__ sum_b) {
    constexpr int C = 4 * C4;
    extern __shared__ __align__(16) float pw_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* At = pw_smem + (size_t)wid * (2 * C * kPairTP);       // [C][36]: At[k][a]
    float* Bt = At + C * kPairTP;                                // [C][36]: Bt[k][b] (column sums only)
    const int64_t g = (int64_t)blockIdx.x * kPairWarps + wid;
    if (g >= num_pairs) return;
    const int gb = idx_b ? idx_b[g] : (int)g;
    const int a0 = ptr_a[g], a1 = ptr_a[g + 1], b0 = ptr_b[gb], b1 = ptr_b[gb + 1];
    const int na = a1 - a0, nb = b1 - b0;
    const bool single = na <= 32 && nb <= 32;
    float dotp = 0.f;
    if (!single) {                                       // column sums by a pass over global memory (fixed row order)
        for (int k = lane; k < C; k += 32) {
            float sa = 0.f, sb = 0.f;
            for (int a = a0; a < a1; ++a) sa += xa[(int64_t)a * C + k];
            for (int b = b0; b < b1; ++b) sb += xb[(int64_t)b * C + k];
            sum_a[g * C + k] = sa; sum_b[g * C + k] = sb;
            dotp = fmaf(sa, sb, dotp);
        }
    }
    float best = -INFINITY;
    long long best_i = 0x7fffffffffffffffLL;
    for (int at = 0; at < na || (at == 0 && single); at += 32) {
        __syncwarp();
        {                                                // row at+lane of A -> At[.][lane]
            const bool ok = at + lane < na;
            const float4* row = reinterpret_cast<const float4*>(xa + (int64_t)(a0 + at + (ok ? lane : 0)) * C);
#pragma unroll
            for (int q = 0; q < C4; ++q) {
                const float4 v = ok ? __ldg(row + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                At[(4 * q + 0) * kPairTP + lane] = v.x; At[(4 * q + 1) * kPairTP + lane] = v.y;
                At[(4 * q + 2) * kPairTP + lane] = v.z; At[(4 * q + 3) * kPairTP + lane] = v.w;
            }
        }
        for (int bt = 0; bt < nb || (bt == 0 && single); bt += 32) {
            const bool okb = bt + lane < nb;
            float4 br[C4];
            {
                const float4* row = reinterpret_cast<const float4*>(xb + (int64_t)(b0 + bt + (okb ? lane : 0)) * C);
#pragma unroll
                for (int q = 0; q < C4; ++q) br[q] = okb ? __ldg(row + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (single) {
                __syncwarp();
#pragma unroll
                for (int q = 0; q < C4; ++q) {
                    Bt[(4 * q + 0) * kPairTP + lane] = br[q].x; Bt[(4 * q + 1) * kPairTP + lane] = br[q].y;
                    Bt[(4 * q + 2) * kPairTP + lane] = br[q].z; Bt[(4 * q + 3) * kPairTP + lane] = br[q].w;
                }
            }
            __syncwarp();
            if (single) {                                // column sums out of the transposed tiles, rows in order
                for (int k = lane; k < C; k += 32) {
                    float sa = 0.f, sb = 0.f;
                    for (int a = 0; a < na; ++a) sa += At[k * kPairTP + a];
                    for (int b = 0; b < nb; ++b) sb += Bt[k * kPairTP + b];
                    sum_a[g * C + k] = sa; sum_b[g * C + k] = sb;
                    dotp = fmaf(sa, sb, dotp);
                }
            }
            const int rows = min(32, na - at);
            const int groups = (rows + 3) >> 2;          // warp-uniform
            float acc[32];
#pragma unroll
            for (int a = 0; a < 32; ++a) acc[a] = 0.f;
#pragma unroll
            for (int q = 0; q < C4; ++q) {
                const float bq[4] = {br[q].x, br[q].y, br[q].z, br[q].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* arow = At + (4 * q + i) * kPairTP;
#pragma unroll
                    for (int a4 = 0; a4 < 8; ++a4) {
                        if (a4 < groups) {
                            const float4 av = *reinterpret_cast<const float4*>(arow + 4 * a4);
                            acc[4 * a4 + 0] = fmaf(av.x, bq[i], acc[4 * a4 + 0]); acc[4 * a4 + 1] = fmaf(av.y, bq[i], acc[4 * a4 + 1]);
                            acc[4 * a4 + 2] = fmaf(av.z, bq[i], acc[4 * a4 + 2]); acc[4 * a4 + 3] = fmaf(av.w, bq[i], acc[4 * a4 + 3]);
                        }
                    }
                }
            }
            if (okb) {
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    if (a < rows) {
                        const long long lin = (long long)(at + a) * nb + (bt + lane);
                        if (acc[a] > best || (acc[a] == best && lin < best_i)) { best = acc[a]; best_i = lin; }
                    }
                }
            }
        }
    }
    dotp = warp_sum(dotp);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) {
        if (na > 0 && nb > 0) {
            out[2 * g] = best;
            out[2 * g + 1] = dotp / ((float)na * (float)nb);
            argmax[2 * g] = a0 + (int)(best_i / nb);
            argmax[2 * g + 1] = b0 + (int)(best_i % nb);
        } else {
            out[2 * g] = 0.f; out[2 * g + 1] = 0.f;
            argmax[2 * g] = -1; argmax[2 * g + 1] = -1;
        }
    }
}

__global__ void __launch_bounds__(kPairThreads)
pair_dot_pool_bwd_kernel(const float* __restrict__ xa, const float* __restrict__ xb, const int32_t* __restrict__ ptr_a,
                         const int32_t* __restrict__ ptr_b, const float* __restrict__ g_out,
                         const int32_t* __restrict__ argmax, const float* __restrict__ sum_a,
                         const float* __restrict__ sum_b, int C, float* __restrict__ g_xa, float* __restrict__ g_xb) {
    const int g = blockIdx.x, t = threadIdx.x;
    const int a0 = ptr_a[g], a1 = ptr_a[g + 1], b0 = ptr_b[g], b1 = ptr_b[g + 1];
    const int na = a1 - a0, nb = b1 - b0;
    const bool ok = na > 0 && nb > 0;
    const float gmax = ok ? g_out[2 * (int64_t)g] : 0.f;
    const float gmean = ok ? g_out[2 * (int64_t)g + 1] / ((float)na * (float)nb) : 0.f;
    const int ia = argmax[2 * (int64_t)g], ib = argmax[2 * (int64_t)g + 1];
    for (int idx = t; idx < na * C; idx += kPairThreads) {
        int r = idx / C, k = idx - r * C;
        float v = gmean * sum_b[(int64_t)g * C + k];
        if (a0 + r == ia) v = fmaf(gmax, xb[(int64_t)ib * C + k], v);
        g_xa[(int64_t)(a0 + r) * C + k] = v;
    }
    for (int idx = t; idx < nb * C; idx += kPairThreads) {
        int r = idx / C, k = idx - r * C;
        float v = gmean * sum_a[(int64_t)g * C + k];
        if (b0 + r == ib) v = fmaf(gmax, xa[(int64_t)ia * C + k], v);
        g_xb[(int64_t)(b0 + r) * C + k] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core variant (tf32 math mode, C % 4 == 0, C <= 64): S = Xa Xb^T is a real GEMM for drug-target pairs (25 x ~500 x C),
// and the CUDA-core kernel above spends 9 shared-memory loads per 8 FMAs on it (204 us for 256 pairs, 1.5 % of the HBM
// roofline).  One CTA per pair: the ligand rows (<= 128 per pass) and 256 protein rows at a time go into K-major SWIZZLE_128B
// operand panels (both are row-major [rows][C]: rows = M resp. N, channels = K), one lane issues tcgen05.mma.kind::tf32 into a
// [128 x 256] TMEM accumulator, and the four warps scan their lane quarter (thread = ligand row) for the running max / first
// arg-max.  The mean and the column sums that backward needs stay exact fp32 (accumulated in a fixed order while the rows are
// staged).  The max is TF32-accurate (operands rounded to 10 mantissa bits, fp32 accumulation), like every other projection
// of this math mode.
constexpr int kPairTcThreads = 128;
constexpr int kPairTcN = 256;

template <int KP>                                       // operand panels of 32 channels: 1 (C <= 32) or 2 (C <= 64)
__global__ void __launch_bounds__(kPairTcThreads)
pair_dot_pool_tc_kernel(const float* __restrict__ xa, const float* __restrict__ xb, const int32_t* __restrict__ ptr_a,
                        const int32_t* __restrict__ ptr_b, const int32_t* __restrict__ idx_b, int C, float* __restrict__ out,
                        int32_t* __restrict__ argmax, float* __restrict__ sum_a, float* __restrict__ sum_b) {
    using namespace tc;
    using namespace mp;
    extern __shared__ uint8_t dsm_raw[];
    uint8_t* dsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B panels
    uint8_t* AP = dsm;                                  // [KP][128 rows][128 B]
    uint8_t* BP = dsm + KP * kMpPanel;                  // [KP][256 rows][128 B]
    float* part = reinterpret_cast<float*>(BP + KP * kPairTcN * 128);      // [row lanes][64] column-sum partials
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float red_v[4];
    __shared__ long long red_i[4];
    const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int gb = idx_b ? idx_b[g] : g;
    const int a0 = ptr_a[g], na = ptr_a[g + 1] - a0, b0 = ptr_b[gb], nb = ptr_b[gb + 1] - b0;
    const int CQ = C >> 2, KC = (CQ + 1) & ~1;          // 16-byte chunks per row; chunks of the K extent (k-steps of 2 chunks)
    if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (wid == 1) tmem_alloc(&tmem_slot, (uint32_t)kPairTcN);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot, lane_base = tmem_base + ((uint32_t)(wid * 32) << 16);
    const uint64_t d_a = make_smem_desc(smem_u32(AP), 16, 1024), d_b = make_smem_desc(smem_u32(BP), 16, 1024);
    // staging map: thread -> (16-byte chunk q, row lane rl); rows rl, rl + RL, ... : a warp request covers ~3.5 contiguous rows
    const int RL = kPairTcThreads / CQ, q = t % CQ, rl = t / CQ;
    const bool stager = rl < RL;
    uint32_t ph = 0;
    float best = -INFINITY;
    long long best_i = 0x7fffffffffffffffLL;
    float4 csum_b = make_float4(0.f, 0.f, 0.f, 0.f);    // this thread's share of the protein column sums (chunk q)
    for (int at = 0; at < na; at += kMpM) {
        const int ma = min(kMpM, na - at);
        // ---- ligand rows -> A panels (zero rows beyond ma, zero K padding); column sums of the ligand on the way
        float4 csum_a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (stager)
            for (int r0 = rl; r0 < kMpM; r0 += 5 * RL) {
                float4 v[5];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int r = r0 + u * RL;
                    v[u] = r < ma ? __ldg(reinterpret_cast<const float4*>(xa + (int64_t)(a0 + at + r) * C) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int r = r0 + u * RL;
                    if (r < kMpM) {
                        csum_a.x += v[u].x; csum_a.y += v[u].y; csum_a.z += v[u].z; csum_a.w += v[u].w;
                        sts128(AP + (q >> 3) * kMpPanel + pan_off(r, q), v[u]);
                    }
                }
            }
        if (KC > CQ && t < kMpM) sts128(AP + (CQ >> 3) * kMpPanel + pan_off(t, CQ), make_float4(0.f, 0.f, 0.f, 0.f));
        if (stager) *reinterpret_cast<float4*>(part + (rl * 16 + q) * 4) = csum_a;
        __syncthreads();
        if (t < C) {                                    // fixed order over the row lanes, accumulated over the a passes
            float sa = at == 0 ? 0.f : sum_a[(int64_t)g * C + t];
            for (int k = 0; k < RL; ++k) sa += part[(k * 16 + (t >> 2)) * 4 + (t & 3)];
            sum_a[(int64_t)g * C + t] = sa;
        }
        for (int bt = 0; bt < nb; bt += kPairTcN) {
            const int mb = min(kPairTcN, nb - bt);
            __syncthreads();                            // previous epilogue / partial reads done
            if (stager)
                for (int r0 = rl; r0 < kPairTcN; r0 += 6 * RL) {                 // six rows in flight per thread
                    float4 v[6];
#pragma unroll
                    for (int u = 0; u < 6; ++u) {
                        const int r = r0 + u * RL;
                        v[u] = r < mb ? __ldg(reinterpret_cast<const float4*>(xb + (int64_t)(b0 + bt + r) * C) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 6; ++u) {
                        const int r = r0 + u * RL;
                        if (r < kPairTcN) {
                            if (at == 0) { csum_b.x += v[u].x; csum_b.y += v[u].y; csum_b.z += v[u].z; csum_b.w += v[u].w; }
                            sts128(BP + (q >> 3) * (kPairTcN * 128) + pan_off(r, q), v[u]);
                        }
                    }
                }
            if (KC > CQ)
                for (int r = t; r < kPairTcN; r += kPairTcThreads) sts128(BP + (CQ >> 3) * (kPairTcN * 128) + pan_off(r, CQ), make_float4(0.f, 0.f, 0.f, 0.f));
            fence_proxy_async_smem();
            __syncthreads();
            if (t == 0) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_tf32(kMpM, kPairTcN, 0, 0);
                for (int i = 0; i < KC / 2; ++i) {
                    const int qq = 2 * i;
                    mma_tf32_ss(tmem_base, d_a + (uint64_t)((((qq >> 3) * kMpPanel) + (qq & 7) * 16) >> 4),
                                d_b + (uint64_t)((((qq >> 3) * (kPairTcN * 128)) + (qq & 7) * 16) >> 4), idesc, i > 0 ? 1u : 0u);
                }
                mma_commit(&bar);
            }
            mbar_wait_guarded(&bar, ph); ph ^= 1u;
            tc_fence_after_sync();
            // ---- scan: thread = ligand row (TMEM lane); first arg-max in row-major order (strict > inside the row)
            const int row = wid * 32 + lane;
            const bool live = row < ma;                 // (tcgen05.ld is warp-aligned: every lane loads, live rows update)
            for (int c0 = 0; c0 < mb; c0 += 16) {
                float v[16];
                tmem_ld16(lane_base + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (live && c0 + i < mb && v[i] > best) { best = v[i]; best_i = (long long)(at + row) * nb + (bt + c0 + i); }
            }
            tc_fence_before_sync();
        }
    }
    // ---- protein column sums, mean, block arg-max (value desc, linear index asc)
    __syncthreads();
    if (stager) *reinterpret_cast<float4*>(part + (rl * 16 + q) * 4) = csum_b;
    __syncthreads();
    float dotp = 0.f;
    if (t < C) {
        float sb = 0.f;
        for (int k = 0; k < RL; ++k) sb += part[(k * 16 + (t >> 2)) * 4 + (t & 3)];
        sum_b[(int64_t)g * C + t] = sb;
        dotp = sum_a[(int64_t)g * C + t] * sb;
    }
    __shared__ float red_dot[4];
    dotp = warp_sum(dotp);
    if (lane == 0) red_dot[wid] = dotp;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = best_i; }
    __syncthreads();
    if (t == 0) {
        float d = 0.f;
        for (int w = 0; w < 4; ++w) {
            d += red_dot[w];
            if (red_v[w] > best || (red_v[w] == best && red_i[w] < best_i)) { best = red_v[w]; best_i = red_i[w]; }
        }
        if (na > 0 && nb > 0) {
            out[2 * (int64_t)g] = best;
            out[2 * (int64_t)g + 1] = d / ((float)na * (float)nb);
            argmax[2 * (int64_t)g] = a0 + (int)(best_i / nb);
            argmax[2 * (int64_t)g + 1] = b0 + (int)(best_i % nb);
        } else {
            out[2 * (int64_t)g] = 0.f; out[2 * (int64_t)g + 1] = 0.f;
            argmax[2 * (int64_t)g] = -1; argmax[2 * (int64_t)g + 1] = -1;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (wid == 1) tmem_dealloc(tmem_base, (uint32_t)kPairTcN);
}

static bool pair_tc_ok(const float* xa, const float* xb, int C) {
    return g_math_mode_get() != 0 && (C & 3) == 0 && C >= 32 && C <= 64 && (((uintptr_t)xa | (uintptr_t)xb) & 15) == 0;
}
static int pair_tc_launch(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b, const int32_t* idx_b,
                          int64_t num_pairs, int C, float* out, int32_t* argmax, float* sum_a, float* sum_b, cudaStream_t stream) {
    const int KP = C <= 32 ? 1 : 2;
    const size_t smem = (size_t)KP * (mp::kMpPanel + kPairTcN * 128) + sizeof(float) * 16 * 16 * 4 + 1024;
    auto go = [&](auto kern) {
        ensure_dyn_smem((const void*)kern, (size_t)((int)smem));
        kern<<<(unsigned)num_pairs, kPairTcThreads, smem, stream>>>(xa, xb, ptr_a, ptr_b, idx_b, C, out, argmax, sum_a, sum_b);
    };
    if (KP == 1) go(pair_dot_pool_tc_kernel<1>); else go(pair_dot_pool_tc_kernel<2>);
    return 0;
}

}  // namespace glam

using namespace glam;

extern "C" int glam_pair_dot_pool_fwd(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                                      int64_t num_pairs, int C, float* out, int32_t* argmax, float* sum_a, float* sum_b,
                                      void* stream_) {
    GLAM_REQUIRE(num_pairs >= 0 && C > 0 && C <= 512, "glam_pair_dot_pool_fwd: bad shape");
    if (num_pairs == 0) return 0;
    GLAM_REQUIRE(xa && xb && ptr_a && ptr_b && out && argmax && sum_a && sum_b, "glam_pair_dot_pool_fwd: null pointer");
    GLAM_REQUIRE(num_pairs < (int64_t)1 << 31, "glam_pair_dot_pool_fwd: too many pairs");
    const size_t smem = sizeof(float) * 2 * kPairTile * (C | 1);
    if (smem > 48 * 1024)
        ensure_dyn_smem((const void*)pair_dot_pool_fwd_kernel, (size_t)((int)smem));
    pair_dot_pool_fwd_kernel<<<(unsigned)num_pairs, kPairThreads, smem, (cudaStream_t)stream_>>>(xa, xb, ptr_a, ptr_b, nullptr, C, out,
                                                                                               argmax, sum_a, sum_b);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_pair_dot_pool_fwd_idx(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                                          const int32_t* idx_b, int64_t num_pairs, int C, float* out, int32_t* argmax,
                                          float* sum_a, float* sum_b, void* stream_) {
    GLAM_REQUIRE(num_pairs >= 0 && C > 0 && C <= 512, "glam_pair_dot_pool_fwd_idx: bad shape");
    if (num_pairs == 0) return 0;
    GLAM_REQUIRE(xa && xb && ptr_a && ptr_b && idx_b && out && argmax && sum_a && sum_b, "glam_pair_dot_pool_fwd_idx: null pointer");
    GLAM_REQUIRE(num_pairs < (int64_t)1 << 31, "glam_pair_dot_pool_fwd_idx: too many pairs");
    const size_t smem = sizeof(float) * 2 * kPairTile * (C | 1);
    if (smem > 48 * 1024)
        ensure_dyn_smem((const void*)pair_dot_pool_fwd_kernel, (size_t)((int)smem));
    pair_dot_pool_fwd_kernel<<<(unsigned)num_pairs, kPairThreads, smem, (cudaStream_t)stream_>>>(xa, xb, ptr_a, ptr_b, idx_b, C, out,
                                                                                               argmax, sum_a, sum_b);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_pair_dot_pool_small_supported(int channels) {
    return (channels == 32 || channels == 36 || channels == 48 || channels == 64) ? 1 : 0;
}

extern "C" int glam_pair_dot_pool_fwd_small(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                                            const int32_t* idx_b, int64_t num_pairs, int C, float* out, int32_t* argmax,
                                            float* sum_a, float* sum_b, void* stream_) {
    GLAM_REQUIRE(num_pairs >= 0 && num_pairs < (int64_t)1 << 31, "glam_pair_dot_pool_fwd_small: bad shape");
    if (num_pairs == 0) return 0;
    GLAM_REQUIRE(xa && xb && ptr_a && ptr_b && out && argmax && sum_a && sum_b, "glam_pair_dot_pool_fwd_small: null pointer");
    GLAM_REQUIRE(glam_pair_dot_pool_small_supported(C) && (((uintptr_t)xa | (uintptr_t)xb) & 15) == 0,
                 "glam_pair_dot_pool_fwd_small: channels must be 32, 36, 48 or 64 and the rows 16-byte aligned (channels=%d)", C);
    const size_t smem = sizeof(float) * kPairWarps * 2 * (size_t)C * kPairTP;
    const unsigned grid = (unsigned)((num_pairs + kPairWarps - 1) / kPairWarps);
    auto go = [&](auto kern) {
        ensure_dyn_smem((const void*)kern, smem);
        kern<<<grid, kPairWarps * 32, smem, (cudaStream_t)stream_>>>(xa, xb, ptr_a, ptr_b, idx_b, num_pairs, out, argmax, sum_a, sum_b);
    };
    switch (C) {
        case 32: go(pair_dot_pool_fwd_warp_kernel<8>); break;
        case 36: go(pair_dot_pool_fwd_warp_kernel<9>); break;
        case 48: go(pair_dot_pool_fwd_warp_kernel<12>); break;
        default: go(pair_dot_pool_fwd_warp_kernel<16>); break;
    }
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_pair_dot_pool_tc_supported(int channels) {
    return (g_math_mode_get() != 0 && (channels & 3) == 0 && channels >= 32 && channels <= 64) ? 1 : 0;
}

extern "C" int glam_pair_dot_pool_fwd_tc(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                                         const int32_t* idx_b, int64_t num_pairs, int C, float* out, int32_t* argmax,
                                         float* sum_a, float* sum_b, void* stream_) {
    GLAM_REQUIRE(num_pairs >= 0 && num_pairs < (int64_t)1 << 31, "glam_pair_dot_pool_fwd_tc: bad shape");
    if (num_pairs == 0) return 0;
    GLAM_REQUIRE(xa && xb && ptr_a && ptr_b && out && argmax && sum_a && sum_b, "glam_pair_dot_pool_fwd_tc: null pointer");
    GLAM_REQUIRE(pair_tc_ok(xa, xb, C), "glam_pair_dot_pool_fwd_tc: needs tf32 math mode, channels %% 4 == 0 in [32, 64] and 16-byte aligned rows "
                 "(channels=%d, math mode %d)", C, g_math_mode_get());
    pair_tc_launch(xa, xb, ptr_a, ptr_b, idx_b, num_pairs, C, out, argmax, sum_a, sum_b, (cudaStream_t)stream_);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" int glam_pair_dot_pool_bwd(const float* xa, const float* xb, const int32_t* ptr_a, const int32_t* ptr_b,
                                      const float* g_out, const int32_t* argmax, const float* sum_a, const float* sum_b,
                                      int64_t num_pairs, int C, float* g_xa, float* g_xb, void* stream_) {
    GLAM_REQUIRE(num_pairs >= 0 && C > 0, "glam_pair_dot_pool_bwd: bad shape");
    if (num_pairs == 0) return 0;
    GLAM_REQUIRE(xa && xb && ptr_a && ptr_b && g_out && argmax && sum_a && sum_b && g_xa && g_xb,
                 "glam_pair_dot_pool_bwd: null pointer");
    pair_dot_pool_bwd_kernel<<<(unsigned)num_pairs, kPairThreads, 0, (cudaStream_t)stream_>>>(xa, xb, ptr_a, ptr_b, g_out, argmax, sum_a,
                                                                                            sum_b, C, g_xa, g_xb);
    GLAM_CHECK_LAUNCH();
    return 0;
}
