// Small dense contractions on the fp32 pipe (see include/glam_b200.h (2)).
//
// The projections of the hot path have a huge M (nodes) and tiny K/N (36..276), so they are streaming,
// HBM-bound-if-fast-enough contractions.  This file is the exact-fp32 path used for parity; the tcgen05
// path for the same contractions lives in proj_tc.cu.
#include "common.cuh"

namespace glam {

bool tc_gemm_eligible(const float* X, int64_t ldx, const float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K);
int tc_gemm_launch(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                   const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, int epi,
                   int exact_begin, int exact_end, cudaStream_t stream);

bool tc_gemm_tn_eligible(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb, int want_colsum);
size_t tc_gemm_tn_workspace(int64_t M, int64_t Ka, int64_t Kb, int want_colsum);
int tc_gemm_tn_launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb, float* out,
                      int64_t ldo, int transpose_out, float* colsum_b, void* workspace, cudaStream_t stream);

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kGemmThreads = (BM / TM) * (BN / TN);  // 256

enum Epi { EPI_NONE = 0, EPI_CELU = 1, EPI_MUL_CELU_GRAD = 2, EPI_ACCUM = 3 };

__global__ void __launch_bounds__(kGemmThreads)
gemm_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t w_sk, int64_t w_sn,
            const float* __restrict__ bias, const float* __restrict__ aux, int64_t ldaux, float* __restrict__ Y,
            int64_t ldy, int64_t M, int N, int K, int epi) {
    __shared__ __align__(16) float Xs[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int t = threadIdx.x;
    const int tx = t % (BN / TN), ty = t / (BN / TN);
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int it = 0; it < BM * BK / kGemmThreads; ++it) {
            int idx = t + it * kGemmThreads;
            int r = idx / BK, kk = idx % BK;
            int64_t m = m0 + r;
            int k = k0 + kk;
            Xs[kk][r] = (m < M && k < K) ? X[m * ldx + k] : 0.f;
        }
        if (w_sn == 1) {
#pragma unroll
            for (int it = 0; it < BN * BK / kGemmThreads; ++it) {
                int idx = t + it * kGemmThreads;
                int kk = idx / BN, n = idx % BN;
                int k = k0 + kk, nn = n0 + n;
                Ws[kk][n] = (k < K && nn < N) ? W[(int64_t)k * w_sk + nn] : 0.f;
            }
        } else {
#pragma unroll
            for (int it = 0; it < BN * BK / kGemmThreads; ++it) {
                int idx = t + it * kGemmThreads;
                int n = idx / BK, kk = idx % BK;
                int k = k0 + kk, nn = n0 + n;
                Ws[kk][n] = (k < K && nn < N) ? W[(int64_t)k * w_sk + (int64_t)nn * w_sn] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * TM]);
            float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * TN]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (epi == EPI_CELU) v = celu1(v);
            else if (epi == EPI_MUL_CELU_GRAD) { float y = aux[m * ldaux + n]; v *= (y > 0.f ? 1.f : y + 1.f); }
            else if (epi == EPI_ACCUM) v += Y[m * ldy + n];
            Y[m * ldy + n] = v;
        }
    }
}

// out_partial[s][Ka][Kb] = sum over the s-th row chunk of A[m,ka]*B[m,kb]
__global__ void __launch_bounds__(kGemmThreads)
gemm_tn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int64_t M, int Ka,
               int Kb, int tiles_b, int64_t rows_per_split, float* __restrict__ partial) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const int tx = t % (BN / TN), ty = t / (BN / TN);
    const int a0 = (blockIdx.x / tiles_b) * BM, b0 = (blockIdx.x % tiles_b) * BN;
    const int64_t mbeg = (int64_t)blockIdx.y * rows_per_split;
    int64_t mend = mbeg + rows_per_split;
    if (mend > M) mend = M;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int64_t m0 = mbeg; m0 < mend; m0 += BK) {
#pragma unroll
        for (int it = 0; it < BM * BK / kGemmThreads; ++it) {
            int idx = t + it * kGemmThreads;
            int mm = idx / BM, c = idx % BM;
            int64_t m = m0 + mm;
            As[mm][c] = (m < mend && a0 + c < Ka) ? A[m * lda + a0 + c] : 0.f;
            Bs[mm][c] = (m < mend && b0 + c < Kb) ? B[m * ldb + b0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* P = partial + (int64_t)blockIdx.y * Ka * Kb;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int a = a0 + ty * TM + i;
        if (a >= Ka) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int b = b0 + tx * TN + j;
            if (b < Kb) P[(int64_t)a * Kb + b] = acc[i][j];
        }
    }
}

// out[r*ldo + c] = sum_{s<S} partial[s][r*cols + c], s ascending (fixed order => reproducible)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int S, int64_t rows, int cols, float* __restrict__ out,
                                       int64_t ldo, int transpose_out = 0) {
    int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < S; ++k) s += partial[(int64_t)k * total + i];
        if (transpose_out) out[(i % cols) * ldo + (i / cols)] = s; else out[(i / cols) * ldo + (i % cols)] = s;
    }
}

__global__ void colsum_partial_kernel(const float* __restrict__ G, int64_t ldg, int64_t M, int N, int64_t rows_per_split,
                                      float* __restrict__ partial) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t mbeg = (int64_t)blockIdx.x * rows_per_split;
    int64_t mend = mbeg + rows_per_split;
    if (mend > M) mend = M;
    for (int c0 = 0; c0 < N; c0 += 32) {
        int c = c0 + tx;
        float s = 0.f;
        if (c < N)
            for (int64_t m = mbeg + ty; m < mend; m += 8) s += G[m * ldg + c];
        red[ty][tx] = s;
        __syncthreads();
        if (ty == 0 && c < N) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += red[k][tx];
            partial[(int64_t)blockIdx.x * N + c] = v;
        }
        __syncthreads();
    }
}

// Skinny weight gradients in exact fp32: out[Wp, Wq] = sum_m P[m, Wp] * Q[m, Wq] with Wq <= 8 (attention-logit
// columns s_i/s_j, edge-attention weights).  These sums cancel almost completely (softmax gradients are zero-sum per
// destination), so they are kept out of the TF32 path; they are also far too narrow for a tensor-core tile.
// One warp per row (lanes over the wide operand's features), 8 warps x S CTAs, fixed-order reductions.
constexpr int kSkinnyWarps = 8;
// QW = 8 or 16: register width of the narrow operand (16 covers the 9 / 15-feature input projections of the models, whose
// weight gradients otherwise fell to the scalar fp32 A^T B kernel: 81 us for 18 MB of operands at the bench shape)
// colsum: 0 none, 1 column sums of P, 2 column sums of Q ride along (the bias gradient of a LinearBlock: no second pass over B);
// they follow the [Wp][Wq] block of the CTA's partial.
template <int KPL, int QW = 8>
__global__ void __launch_bounds__(kSkinnyWarps * 32)
skinny_tn_kernel(const float* __restrict__ P, int64_t ldp, int Wp, const float* __restrict__ Q, int64_t ldq, int Wq, int64_t M,
                 int64_t rows_per_cta, float* __restrict__ partial, int colsum) {
    extern __shared__ float red[];                     // [kSkinnyWarps][Wp][QW]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t mbeg = (int64_t)blockIdx.x * rows_per_cta;
    int64_t mend = mbeg + rows_per_cta;
    if (mend > M) mend = M;
    float acc[KPL][QW], csp[KPL], csq[QW];
#pragma unroll
    for (int t = 0; t < KPL; ++t) {
        csp[t] = 0.f;
#pragma unroll
        for (int j = 0; j < QW; ++j) acc[t][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < QW; ++j) csq[j] = 0.f;
    // 4 rows per iteration: 4 independent load groups in flight per warp
    for (int64_t m0 = mbeg + warp; m0 < mend; m0 += 4 * kSkinnyWarps) {
        float q[4][QW], a[4][KPL];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t m = m0 + (int64_t)u * kSkinnyWarps;
            const bool ok = m < mend;
#pragma unroll
            for (int j = 0; j < QW; ++j) q[u][j] = (ok && j < Wq) ? Q[m * ldq + j] : 0.f;
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                const int f = lane + 32 * t;
                a[u][t] = (ok && f < Wp) ? P[m * ldp + f] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int t = 0; t < KPL; ++t) {
                csp[t] += a[u][t];
#pragma unroll
                for (int j = 0; j < QW; ++j) acc[t][j] = fmaf(a[u][t], q[u][j], acc[t][j]);
            }
#pragma unroll
            for (int j = 0; j < QW; ++j) csq[j] += q[u][j];
        }
    }
#pragma unroll
    for (int t = 0; t < KPL; ++t) {
        const int f = lane + 32 * t;
        if (f < Wp)
#pragma unroll
            for (int j = 0; j < QW; ++j) red[((size_t)warp * Wp + f) * QW + j] = acc[t][j];
    }
    __syncthreads();
    const int extra = colsum == 1 ? Wp : colsum == 2 ? Wq : 0;
    float* out = partial + (int64_t)blockIdx.x * (Wp * Wq + extra);
    for (int idx = threadIdx.x; idx < Wp * Wq; idx += blockDim.x) {
        const int f = idx / Wq, j = idx - f * Wq;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kSkinnyWarps; ++w) v += red[((size_t)w * Wp + f) * QW + j];
        out[idx] = v;
    }
    if (colsum) {                                      // second use of the staging area: per-warp column sums, fixed order over warps
        __syncthreads();
        if (colsum == 1) {
#pragma unroll
            for (int t = 0; t < KPL; ++t) { const int f = lane + 32 * t; if (f < Wp) red[warp * Wp + f] = csp[t]; }
        } else if (lane == 0) {
#pragma unroll
            for (int j = 0; j < QW; ++j) if (j < Wq) red[warp * Wq + j] = csq[j];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < extra; idx += blockDim.x) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kSkinnyWarps; ++w) v += red[w * extra + idx];
            out[Wp * Wq + idx] = v;
        }
    }
}

// The same product for Wp <= 64, Wq <= 16 (every skinny product of the models), staged through shared memory.  The warp-per-row
// kernel above keeps ~11 warps per SM each waiting on one row's loads (41 us for 18 MB at the bench shape, 6 % of HBM rate).
// Here 96-row tiles of both operands arrive by cp.async through a 3-deep ring (2 CTAs per SM, 64 KB in flight per SM).
// A lane owns four features (one 16-byte chunk of the P row) of one of RS = 2..4 rows, so a warp pass covers RS rows with
// ONE LDS.128 of P and QW/4 LDS.128 of Q (shared-memory cycles, not FMAs, bounded the first version: a broadcast LDS.128
// still costs four).  Column Wp of the P tile is a column of ones, so colsum(Q) is one more feature; colsum(P) is one FADD
// per value.  Every reduction runs in a fixed order.
constexpr int kStRows = 96, kStPitch = 68, kStThreads = 256, kStStages = 3;
constexpr int kStSmem = kStStages * (kStRows * kStPitch + kStRows * 16) * 4;
__device__ __forceinline__ void st_cp4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void st_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int QW>
__global__ void __launch_bounds__(kStThreads)
skinny_tile_kernel(const float* __restrict__ P, int64_t ldp, int Wp, const float* __restrict__ Q, int64_t ldq, int Wq, int64_t M,
                   int tiles_per_cta, float* __restrict__ partial, int colsum, int p_vec) {
    extern __shared__ __align__(16) float st_smem[];
    float* const Qs = st_smem;                                             // [stages][96][16]
    float* const Ps = st_smem + kStStages * kStRows * 16;                  // [stages][96][68]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int FG = (Wp + (colsum == 2 ? 1 : 0) + 3) / 4;                   // feature groups of 4 (<= 16)
    const int RS = FG <= 8 ? 4 : FG <= 10 ? 3 : 2;                         // rows per warp pass
    const int fg = lane % FG, rsub = lane / FG;
    const bool live = rsub < RS;
    const int64_t ntiles = (M + kStRows - 1) / kStRows;
    const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
    const int n = (int)((tile0 + tiles_per_cta <= ntiles ? tile0 + tiles_per_cta : ntiles) - tile0);
    const int WpQ = Wp >> 2;
    auto issue = [&](int buf, int64_t tile) {
        const int64_t row0 = tile * kStRows;
        const int rows = (int)(M - row0 < kStRows ? M - row0 : kStRows);
        float* ps = Ps + buf * kStRows * kStPitch; float* qs = Qs + buf * kStRows * 16;
        if (p_vec) {                                                        // rows are 16-byte aligned: one copy per 4 features
            int r = t / WpQ, c = t - r * WpQ;
            const int dr = kStThreads / WpQ, dc = kStThreads - dr * WpQ;
            for (; r < kStRows; ) {
                if (r < rows) st_cp16(ps + r * kStPitch + 4 * c, P + (row0 + r) * ldp + 4 * c);
                else *reinterpret_cast<float4*>(ps + r * kStPitch + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
                r += dr; c += dc;
                if (c >= WpQ) { c -= WpQ; ++r; }
            }
        } else {
            int r = t / Wp, c = t - r * Wp;
            const int dr = kStThreads / Wp, dc = kStThreads - dr * Wp;
            for (; r < kStRows; ) {
                if (r < rows) st_cp4(ps + r * kStPitch + c, P + (row0 + r) * ldp + c); else ps[r * kStPitch + c] = 0.f;
                r += dr; c += dc;
                if (c >= Wp) { c -= Wp; ++r; }
            }
        }
        for (int idx = t; idx < kStRows * QW; idx += kStThreads) {
            const int r = idx / QW, c = idx % QW;
            if (r < rows && c < Wq) st_cp4(qs + r * 16 + c, Q + (row0 + r) * ldq + c); else qs[r * 16 + c] = 0.f;
        }
        if (t < kStRows) {                                                  // ones column, and zeros up to the end of its group
            const float one = t < rows ? 1.f : 0.f;
            for (int c = Wp; c < 4 * FG; ++c) ps[t * kStPitch + c] = c == Wp ? one : 0.f;
        }
    };
    float acc[4][QW], csp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < QW; ++j) acc[i][j] = 0.f;
    for (int i = 0; i < kStStages - 1; ++i) {                               // one commit per slot, empty when there is no tile
        if (i < n) issue(i, tile0 + i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int rows_per_warp = kStRows / (kStThreads / 32);                  // 12
    for (int i = 0; i < n; ++i) {
        if (i + kStStages - 1 < n) issue((i + kStStages - 1) % kStStages, tile0 + i + kStStages - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(kStStages - 1) : "memory");
        __syncthreads();
        const float* ps = Ps + (i % kStStages) * kStRows * kStPitch + 4 * fg;
        const float* qs = Qs + (i % kStStages) * kStRows * 16;
        if (live) {
            for (int rr = rsub; rr < rows_per_warp; rr += RS) {
                const int r = warp * rows_per_warp + rr;
                const float4 pv = *reinterpret_cast<const float4*>(ps + r * kStPitch);
                csp[0] += pv.x; csp[1] += pv.y; csp[2] += pv.z; csp[3] += pv.w;
#pragma unroll
                for (int j4 = 0; j4 < QW / 4; ++j4) {
                    const float4 q = *reinterpret_cast<const float4*>(qs + r * 16 + 4 * j4);
                    const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        acc[0][4 * j4 + jj] = fmaf(pv.x, qq[jj], acc[0][4 * j4 + jj]);
                        acc[1][4 * j4 + jj] = fmaf(pv.y, qq[jj], acc[1][4 * j4 + jj]);
                        acc[2][4 * j4 + jj] = fmaf(pv.z, qq[jj], acc[2][4 * j4 + jj]);
                        acc[3][4 * j4 + jj] = fmaf(pv.w, qq[jj], acc[3][4 * j4 + jj]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // every (warp, row slot) leaves its block in the idle ring; the output threads add them in a fixed order
    constexpr int RW = QW + 1;
    float* red = st_smem;                                                   // [8 warps][RS][4 FG][RW]
    if (live) {
        float* rp = red + (((warp * RS + rsub) * FG + fg) * 4) * RW;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < QW; ++j) rp[i * RW + j] = acc[i][j];
            rp[i * RW + QW] = csp[i];
        }
    }
    __syncthreads();
    const int extra = colsum == 1 ? Wp : colsum == 2 ? Wq : 0;
    float* out = partial + (int64_t)blockIdx.x * (Wp * Wq + extra);
    const int slots = (kStThreads / 32) * RS, slot_stride = FG * 4 * RW;
    for (int idx = t; idx < Wp * Wq + extra; idx += kStThreads) {
        int ff, j;
        if (idx < Wp * Wq) { ff = idx / Wq; j = idx - ff * Wq; }
        else if (colsum == 1) { ff = idx - Wp * Wq; j = QW; }
        else { ff = Wp; j = idx - Wp * Wq; }
        const float* rp = red + ff * RW + j;
        float v = 0.f;
        for (int s2 = 0; s2 < slots; ++s2) v += rp[s2 * slot_stride];
        out[idx] = v;
    }
}
static int skinny_tile_grid(int64_t M, int* tiles_per_cta) {
    const int64_t ntiles = (M + kStRows - 1) / kStRows;
    int64_t tpc = (ntiles + 2 * kNumSMs - 1) / (2 * kNumSMs);
    if (tpc < 1) tpc = 1;
    if (tiles_per_cta) *tiles_per_cta = (int)tpc;
    const int64_t g = (ntiles + tpc - 1) / tpc;
    return (int)(g < 1 ? 1 : g);
}

// out[r, c] (or transposed) = sum_s partial[s][r*cols + c]; 4 thread rows split S, combined in a fixed order
__global__ void __launch_bounds__(256)
reduce4_kernel(const float* __restrict__ partial, int S, int rows, int cols, float* __restrict__ out, int64_t ldo, int transpose_out) {
    __shared__ float red4[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int total = rows * cols, i = blockIdx.x * 64 + tx;
    float s = 0.f;
    if (i < total) {
        const int per = (S + 3) / 4, k0 = ty * per, k1 = min(S, k0 + per);
        for (int k = k0; k < k1; ++k) s += partial[(int64_t)k * total + i];
    }
    red4[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < total) {
        const float v = ((red4[0][tx] + red4[1][tx]) + red4[2][tx]) + red4[3][tx];
        const int r = i / cols, c = i - r * cols;
        if (transpose_out) out[(int64_t)c * ldo + r] = v; else out[(int64_t)r * ldo + c] = v;
    }
}

static int skinny_grid(int64_t M) {
    int64_t g = (M + 511) / 512;
    if (g > 2 * kNumSMs) g = 2 * kNumSMs;
    return (int)(g < 1 ? 1 : g);
}

static int tn_splits(int64_t M, int64_t Ka, int64_t Kb) {
    int64_t tiles = ((Ka + BM - 1) / BM) * ((Kb + BN - 1) / BN);
    int64_t s = (2 * kNumSMs + tiles - 1) / tiles;
    int64_t maxs = (M + 255) / 256;
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    return (int)s;
}
static int colsum_splits(int64_t M) {
    int64_t s = (M + 63) / 64;
    if (s > 2 * kNumSMs) s = 2 * kNumSMs;
    if (s < 1) s = 1;
    return (int)s;
}

}  // namespace glam

using namespace glam;

extern "C" int glam_gemm_ex(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                            const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                            int epilogue, int exact_col_begin, int exact_col_end, void* stream_);

extern "C" int glam_gemm(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                         const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                         int epilogue, void* stream_) {
    return glam_gemm_ex(X, ldx, W, w_sk, w_sn, bias, aux, ldaux, Y, ldy, M, N, K, epilogue, 0, 0, stream_);
}

extern "C" int glam_gemm_ex(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                            const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                            int epilogue, int exact_col_begin, int exact_col_end, void* stream_) {
    GLAM_REQUIRE(M >= 0 && N > 0 && K > 0, "glam_gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if (M == 0) return 0;
    GLAM_REQUIRE(X && W && Y, "glam_gemm: null pointer");
    GLAM_REQUIRE(ldx >= K && ldy >= N, "glam_gemm: leading dimension too small");
    GLAM_REQUIRE(epilogue >= 0 && epilogue <= 3, "glam_gemm: unknown epilogue %d", epilogue);
    GLAM_REQUIRE(epilogue != EPI_MUL_CELU_GRAD || (aux && ldaux >= N), "glam_gemm: epilogue 2 needs aux");
    GLAM_REQUIRE(N <= 65535 * BN && K < (1 << 30), "glam_gemm: N/K too large");
    GLAM_REQUIRE(exact_col_begin >= 0 && exact_col_begin <= exact_col_end && exact_col_end <= N, "glam_gemm_ex: bad exact column range");
    if (tc_gemm_eligible(X, ldx, Y, ldy, M, N, K))
        return tc_gemm_launch(X, ldx, W, w_sk, w_sn, bias, aux, ldaux, Y, ldy, M, N, K, epilogue, exact_col_begin,
                              exact_col_end, (cudaStream_t)stream_);
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
    gemm_kernel<<<grid, kGemmThreads, 0, (cudaStream_t)stream_>>>(X, ldx, W, w_sk, w_sn, bias, aux, ldaux, Y, ldy, M, (int)N,
                                                                 (int)K, epilogue);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t glam_gemm_tn_workspace_bytes(int64_t M, int64_t Ka, int64_t Kb) {
    return sizeof(float) * (size_t)tn_splits(M, Ka, Kb) * (size_t)Ka * (size_t)Kb;
}

extern "C" int glam_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb,
                            float* out, int64_t ldo, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLAM_REQUIRE(M >= 0 && Ka > 0 && Kb > 0 && out && ldo >= Kb, "glam_gemm_tn: bad arguments");
    GLAM_REQUIRE(Ka < 65536 && Kb < 65536, "glam_gemm_tn: Ka/Kb too large");
    if (M == 0) {
        cudaMemset2DAsync(out, ldo * sizeof(float), 0, Kb * sizeof(float), Ka, stream);
        return 0;
    }
    GLAM_REQUIRE(A && B && lda >= Ka && ldb >= Kb, "glam_gemm_tn: bad inputs");
    const int S = tn_splits(M, Ka, Kb);
    GLAM_REQUIRE(workspace && workspace_bytes >= sizeof(float) * (size_t)S * Ka * Kb, "glam_gemm_tn: workspace too small");
    const int tiles_a = (int)((Ka + BM - 1) / BM), tiles_b = (int)((Kb + BN - 1) / BN);
    int64_t rps = (M + S - 1) / S;
    rps = (rps + BK - 1) / BK * BK;
    gemm_tn_kernel<<<dim3(tiles_a * tiles_b, S), kGemmThreads, 0, stream>>>(A, lda, B, ldb, M, (int)Ka, (int)Kb, tiles_b, rps,
                                                                          (float*)workspace);
    GLAM_CHECK_LAUNCH();
    int64_t total = Ka * Kb;
    reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const float*)workspace, S, Ka, (int)Kb, out, ldo);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t glam_colsum_workspace_bytes(int64_t M, int64_t N) {
    return sizeof(float) * (size_t)colsum_splits(M) * (size_t)N;
}

extern "C" int glam_colsum(const float* G, int64_t ldg, int64_t M, int64_t N, float* out, void* workspace,
                           size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLAM_REQUIRE(M >= 0 && N > 0 && out, "glam_colsum: bad arguments");
    if (M == 0) {
        cudaMemsetAsync(out, 0, sizeof(float) * N, stream);
        return 0;
    }
    GLAM_REQUIRE(G && ldg >= N, "glam_colsum: bad input");
    const int S = colsum_splits(M);
    GLAM_REQUIRE(workspace && workspace_bytes >= sizeof(float) * (size_t)S * N, "glam_colsum: workspace too small");
    int64_t rps = (M + S - 1) / S;
    colsum_partial_kernel<<<S, dim3(32, 8), 0, stream>>>(G, ldg, M, (int)N, rps, (float*)workspace);
    GLAM_CHECK_LAUNCH();
    reduce_partials_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>((const float*)workspace, S, 1, (int)N, out, N);
    GLAM_CHECK_LAUNCH();
    return 0;
}

extern "C" size_t glam_gemm_tn_ex_workspace_bytes(int64_t M, int64_t Ka, int64_t Kb, int want_colsum) {
    size_t a = glam_gemm_tn_workspace_bytes(M, Ka, Kb), b = want_colsum ? glam_colsum_workspace_bytes(M, Kb) : 0;
    size_t c = tc_gemm_tn_workspace(M, Ka, Kb, want_colsum);
    const int sg = skinny_grid(M), tg = skinny_tile_grid(M, nullptr);
    size_t d = sizeof(float) * (size_t)(sg > tg ? sg : tg) * ((size_t)Ka * (size_t)Kb + (size_t)(Ka > Kb ? Ka : Kb));
    size_t m = a > b ? a : b;
    m = m > c ? m : c;
    return m > d ? m : d;
}

extern "C" int glam_gemm_tn_ex(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb,
                               float* out, int64_t ldo, int transpose_out, float* colsum_b, void* workspace,
                               size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLAM_REQUIRE(M >= 0 && Ka > 0 && Kb > 0 && out && ldo >= (transpose_out ? Ka : Kb), "glam_gemm_tn_ex: bad arguments");
    GLAM_REQUIRE(Ka < 65536 && Kb < 65536, "glam_gemm_tn_ex: Ka/Kb too large");
    if (M == 0) {
        if (transpose_out) cudaMemset2DAsync(out, ldo * sizeof(float), 0, Ka * sizeof(float), Kb, stream);
        else cudaMemset2DAsync(out, ldo * sizeof(float), 0, Kb * sizeof(float), Ka, stream);
        if (colsum_b) cudaMemsetAsync(colsum_b, 0, sizeof(float) * Kb, stream);
        return 0;
    }
    GLAM_REQUIRE(A && B && lda >= Ka && ldb >= Kb, "glam_gemm_tn_ex: bad inputs");
    GLAM_REQUIRE(workspace && workspace_bytes >= glam_gemm_tn_ex_workspace_bytes(M, Ka, Kb, colsum_b != nullptr),
                 "glam_gemm_tn_ex: workspace too small");
    const int64_t knarrow = Ka < Kb ? Ka : Kb;
    if ((knarrow <= 8 || (knarrow <= 16 && Ka <= 64 && Kb <= 64)) && Ka <= 288 && Kb <= 288) {
        // skinny product: exact fp32, result partial is [Wp][Wq] with the wide operand first
        const bool a_wide = Kb <= Ka;
        const float* P = a_wide ? A : B; const float* Q = a_wide ? B : A;
        const int64_t ldp = a_wide ? lda : ldb, ldq = a_wide ? ldb : lda;
        const int Wp = (int)(a_wide ? Ka : Kb), Wq = (int)(a_wide ? Kb : Ka);
        const int cs_mode = !colsum_b ? 0 : (a_wide ? 2 : 1);               // colsum(B): B is Q when A is the wide operand, else P
        const int tr = (a_wide ? 0 : 1) ^ (transpose_out ? 1 : 0);
        if (Wp + (cs_mode == 2 ? 1 : 0) <= 64 && Wq <= 16) {
            int tpc = 1;
            const int S = skinny_tile_grid(M, &tpc);
            auto go = [&](auto fn) {
                ensure_dyn_smem((const void*)fn, (size_t)(kStSmem));
                const int p_vec = (Wp % 4 == 0 && ldp % 4 == 0 && ((uintptr_t)P & 15) == 0) ? 1 : 0;
                fn<<<S, kStThreads, kStSmem, stream>>>(P, ldp, Wp, Q, ldq, Wq, M, tpc, (float*)workspace, cs_mode, p_vec);
            };
            if (Wq <= 8) go(skinny_tile_kernel<8>); else go(skinny_tile_kernel<16>);
            GLAM_CHECK_LAUNCH();
            return launch_reduce_partials((const float*)workspace, S, Wp, Wq, colsum_b ? (int)Kb : 0, out, ldo, tr, colsum_b, stream);
        }
        const int S = skinny_grid(M);
        const int64_t rpc = (M + S - 1) / S;
        const int qw = Wq <= 8 ? 8 : 16;
        const size_t smem = sizeof(float) * kSkinnyWarps * Wp * qw;
        const int kpl = (Wp + 31) / 32;
        auto launch = [&](auto fn) {
            if (smem > 48 * 1024) ensure_dyn_smem((const void*)fn, (size_t)((int)smem));
            fn<<<S, kSkinnyWarps * 32, smem, stream>>>(P, ldp, Wp, Q, ldq, Wq, M, rpc, (float*)workspace, cs_mode);
        };
        if (qw == 16) launch(skinny_tn_kernel<2, 16>);
        else if (kpl <= 2) launch(skinny_tn_kernel<2>); else if (kpl <= 4) launch(skinny_tn_kernel<4>);
        else if (kpl <= 6) launch(skinny_tn_kernel<6>); else launch(skinny_tn_kernel<9>);
        GLAM_CHECK_LAUNCH();
        // partial holds [Wp][Wq]; out is [Ka][Kb]: transposed w.r.t. the partial exactly when A is the narrow operand
        // colsum(B) rode along in the same pass: it follows the product in every partial and leaves through the reduction's out2
        return launch_reduce_partials((const float*)workspace, S, Wp, Wq, colsum_b ? (int)Kb : 0, out, ldo, tr, colsum_b, stream);
    }
    if (tc_gemm_tn_eligible(A, lda, B, ldb, M, Ka, Kb, colsum_b != nullptr))
        return tc_gemm_tn_launch(A, lda, B, ldb, M, Ka, Kb, out, ldo, transpose_out, colsum_b, workspace, stream);
    if (colsum_b && tc_gemm_tn_eligible(A, lda, B, ldb, M, Ka, Kb, 0)) {
        // B is too wide for the M side: product on the tensor cores, column sums by the exact reduction
        if (int rc = tc_gemm_tn_launch(A, lda, B, ldb, M, Ka, Kb, out, ldo, transpose_out, nullptr, workspace, stream)) return rc;
        const int S2 = colsum_splits(M);
        int64_t rps2 = (M + S2 - 1) / S2;
        colsum_partial_kernel<<<S2, dim3(32, 8), 0, stream>>>(B, ldb, M, (int)Kb, rps2, (float*)workspace);
        GLAM_CHECK_LAUNCH();
        return launch_reduce_partials((const float*)workspace, S2, 1, (int)Kb, 0, colsum_b, Kb, 0, nullptr, stream);
    }
    const int S = tn_splits(M, Ka, Kb);
    const int tiles_a = (int)((Ka + BM - 1) / BM), tiles_b = (int)((Kb + BN - 1) / BN);
    int64_t rps = (M + S - 1) / S;
    rps = (rps + BK - 1) / BK * BK;
    gemm_tn_kernel<<<dim3(tiles_a * tiles_b, S), kGemmThreads, 0, stream>>>(A, lda, B, ldb, M, (int)Ka, (int)Kb, tiles_b, rps,
                                                                          (float*)workspace);
    GLAM_CHECK_LAUNCH();
    int64_t total = Ka * Kb;
    reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const float*)workspace, S, Ka, (int)Kb, out, ldo,
                                                                               transpose_out);
    GLAM_CHECK_LAUNCH();
    if (colsum_b) {
        const int S2 = colsum_splits(M);
        int64_t rps2 = (M + S2 - 1) / S2;
        colsum_partial_kernel<<<S2, dim3(32, 8), 0, stream>>>(B, ldb, M, (int)Kb, rps2, (float*)workspace);
        GLAM_CHECK_LAUNCH();
        reduce_partials_kernel<<<(unsigned)((Kb + 255) / 256), 256, 0, stream>>>((const float*)workspace, S2, 1, (int)Kb, colsum_b, Kb);
        GLAM_CHECK_LAUNCH();
    }
    return 0;
}
