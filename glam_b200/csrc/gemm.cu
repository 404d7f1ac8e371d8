// Small dense contractions on the fp32 pipe (see include/glam_b200.h (2)).
//
// The projections of the hot path have a huge M (nodes) and tiny K/N (36..276), so they are streaming,
// HBM-bound-if-fast-enough contractions.  This file is the exact-fp32 path used for parity; the tcgen05
// path for the same contractions lives in proj_tc.cu.
#include "common.cuh"

namespace glam {

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
                                       int64_t ldo) {
    int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < S; ++k) s += partial[(int64_t)k * total + i];
        out[(i / cols) * ldo + (i % cols)] = s;
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

static int tn_splits(int64_t M, int64_t Ka, int64_t Kb) {
    int64_t tiles = ((Ka + BM - 1) / BM) * ((Kb + BN - 1) / BN);
    int64_t s = (2 * kNumSMs + tiles - 1) / tiles;
    int64_t maxs = (M + 255) / 256;
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    return (int)s;
}
static int colsum_splits(int64_t M) {
    int64_t s = (M + 511) / 512;
    if (s > 2 * kNumSMs) s = 2 * kNumSMs;
    if (s < 1) s = 1;
    return (int)s;
}

}  // namespace glam

using namespace glam;

extern "C" int glam_gemm(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                         const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                         int epilogue, void* stream_) {
    GLAM_REQUIRE(M >= 0 && N > 0 && K > 0, "glam_gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if (M == 0) return 0;
    GLAM_REQUIRE(X && W && Y, "glam_gemm: null pointer");
    GLAM_REQUIRE(ldx >= K && ldy >= N, "glam_gemm: leading dimension too small");
    GLAM_REQUIRE(epilogue >= 0 && epilogue <= 3, "glam_gemm: unknown epilogue %d", epilogue);
    GLAM_REQUIRE(epilogue != EPI_MUL_CELU_GRAD || (aux && ldaux >= N), "glam_gemm: epilogue 2 needs aux");
    GLAM_REQUIRE(N <= 65535 * BN && K < (1 << 30), "glam_gemm: N/K too large");
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
