// Weight gradients on the tensor cores:  D[Ka,Kb] = sum_m A[m,Ka] * B[m,Kb]  (+ optional colsum(B))
//
// The reduction runs over the node/edge rows (10^5..10^8) while Ka,Kb are 3..288, so each CTA streams a
// contiguous row range, stages 64-row chunks of A and B as panelised SWIZZLE_128B_BASE32B images — the
// MN-major TF32 operand layout, rows being the MMA K dimension (tc_common.cuh) — and accumulates
// D (128 TMEM lanes x N columns) with 8 tcgen05.mma (K = 8 rows each) per chunk.  Two shared-memory stages
// with one mbarrier each let the loads of chunk i+1 overlap the MMAs of chunk i.  Every CTA writes one fp32
// partial; a second kernel adds the partials in a fixed order, so results are bitwise reproducible.
//
// Bias gradients come for free: feature index Ka of the A image is set to 1.0, so row Ka of D is colsum(B).
#include "common.cuh"
#include "tc_common.cuh"

namespace glam {

using namespace tc;

constexpr int kTnThreads = 128;
constexpr int kTnRows = 64;                 // rows (MMA K) per stage: 8 MMAs
constexpr int kTnStages = 2;

struct TcTnParams {
    const float* P; int64_t ldp; int Wp;    // M-side operand (Wp features (+1 if ones), <= 128)
    const float* Q; int64_t ldq; int Wq;    // N-side operand (Wq features (+1 if ones), <= 256)
    int ones_on_p, ones_on_q;               // append a constant-1 feature (bias gradient)
    int64_t M, rows_per_cta;
    int Npad, q_panels, tmem_cols;
    float* partial;                         // [grid][out_rows][out_cols]
    int out_rows, out_cols, swap;           // swap: D holds the transposed result
};

__device__ __forceinline__ void stage_rows(uint8_t* dst, const float* __restrict__ src, int64_t ld, int width, int ones_at,
                                           int panels, int64_t row0, int64_t row_end, int t) {
    // image: [panels][kTnRows][128 B]; zero rows past row_end (K padding must be zero in both operands)
    const int chunks = panels * 8;
    const int wq = width >> 2;              // full 16-byte chunks available in a source row
    for (int idx = t; idx < kTnRows * chunks; idx += kTnThreads) {
        const int r = idx / chunks, q = idx - r * chunks;
        const int64_t m = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < row_end) {
            const float* s = src + m * ld + 4 * q;
            if (q < wq) v = *reinterpret_cast<const float4*>(s);
            else if (4 * q < width) {       // ragged tail of the row
                v.x = s[0];
                if (4 * q + 1 < width) v.y = s[1];
                if (4 * q + 2 < width) v.z = s[2];
            }
            if (ones_at >= 0 && (ones_at >> 2) == q) {
                const int e = ones_at & 3;
                if (e == 0) v.x = 1.f; else if (e == 1) v.y = 1.f; else if (e == 2) v.z = 1.f; else v.w = 1.f;
            }
        }
        *reinterpret_cast<float4*>(dst + mn32_chunk_offset(r, q, kTnRows)) = v;
    }
}

__global__ void __launch_bounds__(kTnThreads)
tc_gemm_tn_kernel(const TcTnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kTnStages];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int kPPanels = 4;                                  // M = 128 = 4 MN-blocks of 32 features
    const int p_bytes = kPPanels * kTnRows * kPanelRowBytes;     // 32 KB
    const int q_bytes = p.q_panels * kTnRows * kPanelRowBytes;
    const int t = threadIdx.x, warp = t >> 5;

    if (t == 0) {
        for (int s = 0; s < kTnStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, p.Npad, 1, 1);   // both operands MN-major
    const int64_t mbeg = (int64_t)blockIdx.x * p.rows_per_cta;
    int64_t mend = mbeg + p.rows_per_cta;
    if (mend > p.M) mend = p.M;
    const int p_src_panels = (p.Wp + p.ones_on_p + 31) / 32;     // panels that actually carry data
    const int nchunks = mend > mbeg ? (int)((mend - mbeg + kTnRows - 1) / kTnRows) : 0;
    uint32_t phase[kTnStages] = {0, 0};
    int issued[kTnStages] = {0, 0};

    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        uint8_t* Ps = smem + (size_t)s * (p_bytes + q_bytes);
        uint8_t* Qs = Ps + p_bytes;
        if (issued[s]) {                                         // MMAs that read this stage must be done
            mbar_wait(&bars[s], phase[s]);
            phase[s] ^= 1;
        }
        const int64_t row0 = mbeg + (int64_t)c * kTnRows;
        stage_rows(Ps, p.P, p.ldp, p.Wp, p.ones_on_p ? p.Wp : -1, p_src_panels, row0, mend, t);
        stage_rows(Qs, p.Q, p.ldq, p.Wq, p.ones_on_q ? p.Wq : -1, p.q_panels, row0, mend, t);
        fence_proxy_async_smem();
        __syncthreads();
        if (t == 0) {
            tc_fence_after_sync();
            const uint32_t pa = smem_u32(Ps), qa = smem_u32(Qs);
#pragma unroll
            for (int g = 0; g < kTnRows / 8; ++g) {
                // MN-major (SW128_32B): LBO = bytes between 32-feature blocks (panel stride), SBO = bytes between 4-row K groups
                const uint64_t da = make_smem_desc(pa + g * 1024, kTnRows * kPanelRowBytes, 512, kLayoutSw128Base32);
                const uint64_t db = make_smem_desc(qa + g * 1024, kTnRows * kPanelRowBytes, 512, kLayoutSw128Base32);
                mma_tf32_ss(tmem_base, da, db, idesc, (c > 0 || g > 0) ? 1u : 0u);
            }
            mma_commit(&bars[s]);
        }
        issued[s] = 1;
    }
    // drain: the last commit covers every earlier MMA (commits complete in issue order)
    if (nchunks > 0) {
        const int s = (nchunks - 1) & 1;
        mbar_wait(&bars[s], phase[s]);
    }
    tc_fence_after_sync();
    // ---- epilogue: thread = D row (M-side feature) ; write this CTA's partial
    float* part = p.partial + (int64_t)blockIdx.x * p.out_rows * p.out_cols;
    const int prow = t;
    const int p_rows = p.Wp + p.ones_on_p, q_cols = p.Wq + p.ones_on_q;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.Npad; c0 += 16) {
        float v[16];
        if (nchunks > 0) tmem_ld16(lane_base + (uint32_t)c0, v);
        else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (prow >= p_rows) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int qc = c0 + j;
            if (qc < q_cols) {
                if (p.swap) part[(int64_t)qc * p.out_cols + prow] = v[j];
                else part[(int64_t)prow * p.out_cols + qc] = v[j];
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// out[...] = sum over S partials in ascending order; each output element is owned by one thread column, the S
// range is split over 4 thread rows that are combined in a fixed order.
__global__ void __launch_bounds__(256)
tn_reduce_kernel(const float* __restrict__ partial, int S, int rows, int cols, float* __restrict__ out, int64_t ldo,
                 int transpose_out, int split_row, float* __restrict__ out2) {
    __shared__ float red[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int total = rows * cols;
    const int i = blockIdx.x * 64 + tx;
    float s = 0.f;
    if (i < total) {
        const int per = (S + 3) / 4;
        const int k0 = ty * per, k1 = min(S, k0 + per);
        for (int k = k0; k < k1; ++k) s += partial[(int64_t)k * total + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < total) {
        const float v = ((red[0][tx] + red[1][tx]) + red[2][tx]) + red[3][tx];
        const int r = i / cols, c = i - r * cols;
        if (split_row >= 0 && r == split_row) { out2[c] = v; return; }      // the appended ones-row = column sums
        if (transpose_out == 1) out[(int64_t)c * ldo + r] = v; else out[(int64_t)r * ldo + c] = v;
    }
}

static int g_tn_grid(int64_t M) {
    int64_t g = (M + 255) / 256;
    if (g > 2 * kNumSMs) g = 2 * kNumSMs;
    return (int)(g < 1 ? 1 : g);
}

int g_math_mode_get();

bool tc_gemm_tn_eligible(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb, int want_colsum) {
    if (g_math_mode_get() == 0 || M < 1) return false;
    if ((lda & 3) || (ldb & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return false;
    const int64_t ka = Ka + (want_colsum ? 1 : 0);
    // A on the M side (<=128 features) or on the N side (<=256) with B on the M side
    if (ka <= 128 && Kb <= 256) return true;
    if (Kb <= 128 && ka <= 256) return true;
    return false;
}

size_t tc_gemm_tn_workspace(int64_t M, int64_t Ka, int64_t Kb, int want_colsum) {
    return sizeof(float) * (size_t)g_tn_grid(M) * (size_t)(Ka + (want_colsum ? 1 : 0)) * (size_t)Kb;
}

int tc_gemm_tn_launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb, float* out,
                      int64_t ldo, int transpose_out, float* colsum_b, void* workspace, cudaStream_t stream) {
    const int ones = colsum_b ? 1 : 0;
    TcTnParams p;
    const bool a_on_m = (Ka + ones) <= 128 && Kb <= 256;
    if (a_on_m) { p.P = A; p.ldp = lda; p.Wp = (int)Ka; p.ones_on_p = ones; p.Q = B; p.ldq = ldb; p.Wq = (int)Kb; p.ones_on_q = 0; p.swap = 0; }
    else        { p.P = B; p.ldp = ldb; p.Wp = (int)Kb; p.ones_on_p = 0; p.Q = A; p.ldq = lda; p.Wq = (int)Ka; p.ones_on_q = ones; p.swap = 1; }
    p.M = M;
    const int grid = g_tn_grid(M);
    p.rows_per_cta = ((M + grid - 1) / grid + kTnRows - 1) / kTnRows * kTnRows;
    p.Npad = (p.Wq + p.ones_on_q + 31) / 32 * 32;     // whole 32-feature MN blocks
    p.q_panels = (p.Npad + 31) / 32;
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)p.Npad);
    p.partial = (float*)workspace;
    p.out_rows = (int)Ka + ones;             // partial is always stored as [Ka(+1)][Kb]
    p.out_cols = (int)Kb;
    const size_t smem = (size_t)kTnStages * (4 + p.q_panels) * kTnRows * kPanelRowBytes + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("tc_gemm_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = smem;
    }
    tc_gemm_tn_kernel<<<grid, kTnThreads, smem, stream>>>(p);
    GLAM_CHECK_LAUNCH();
    const int total = p.out_rows * p.out_cols;
    tn_reduce_kernel<<<(total + 63) / 64, 256, 0, stream>>>((const float*)workspace, grid, p.out_rows, p.out_cols, out, ldo,
                                                           transpose_out ? 1 : 0, ones ? (int)Ka : -1, colsum_b);
    GLAM_CHECK_LAUNCH();
    return 0;
}

}  // namespace glam
