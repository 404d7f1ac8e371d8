// Weight gradients on the tensor cores:  D[Ka,Kb] = sum_m A[m,Ka] * B[m,Kb]  (+ optional colsum(B))
//
// The reduction runs over the node/edge rows (10^5..10^8) while Ka,Kb are 3..288: a pure streaming read.  One
// persistent CTA per SM owns a contiguous row range; a TMA producer lane streams 64-row chunks of both operands
// into a 2..4-stage ring as 32-feature panels in the SWIZZLE_128B_ATOM_32B layout — the only legal swizzled
// MN-major TF32 operand layout (rows are the MMA K dimension) — an MMA lane issues 8 tcgen05.mma (K = 8 rows
// each) per chunk into one TMEM accumulator that lives for the whole kernel, and tcgen05.commit hands the stage
// back to the producer.  Bias gradients ride along: a second accumulator multiplies the M-side operand with a
// constant "ones" B tile, so column sums cost one extra N=16 MMA per K step and no extra memory traffic.
// Every CTA writes one fp32 partial; a second kernel adds the partials in a fixed order (bitwise reproducible).
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace glam {

using namespace tc;

int make_tmap_rows(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int swizzle);
int g_math_mode_get();

constexpr int kTnThreads = 192;
constexpr int kTnRows = 64;                           // rows (MMA K) per stage: 8 MMAs
constexpr int kTnMaxStages = 4;
constexpr int kTnPanelBytes = kTnRows * kPanelRowBytes;   // 8 KB
constexpr int kPPanels = 4;                           // M = 128 = 4 MN blocks of 32 features

struct TcTnParams {
    int Wp, Wq;                  // features on the M side (<= 128) / N side (<= 256)
    int p_is_b;                  // 1: P = B, Q = A (D holds the transposed result)
    int want_colsum;             // column sums of the M-side operand (must be B)
    int64_t M, rows_per_cta;
    int Npad, p_panels, q_panels, tmem_cols, stages;
    float* partial;              // [grid][Ka*Kb (+ Kb)]
    int Ka, Kb;
};

__global__ void __launch_bounds__(kTnThreads, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_q, const TcTnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kTnMaxStages], empty_bar[kTnMaxStages], done_bar, ones_bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = (kPPanels + p.q_panels) * kTnPanelBytes;
    uint8_t* ring = smem;                                            // [stages][P: 4 panels | Q: q_panels]
    uint8_t* ones = smem + (size_t)p.stages * stage_bytes;           // [64][128 B]: feature 0 == 1
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) {
        for (int s = 0; s < kTnMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        mbar_init(&ones_bar, 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t mbeg = (int64_t)blockIdx.x * p.rows_per_cta;
    int64_t mend = mbeg + p.rows_per_cta;
    if (mend > p.M) mend = p.M;
    const int nchunks = mend > mbeg ? (int)((mend - mbeg + kTnRows - 1) / kTnRows) : 0;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            prefetch_tensormap(&tmap_p);
            prefetch_tensormap(&tmap_q);
            const uint32_t bytes = (uint32_t)((p.p_panels + p.q_panels) * kTnPanelBytes);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % p.stages;
                const uint32_t ph = (uint32_t)(c / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], bytes);
                uint8_t* Ps = ring + (size_t)s * stage_bytes;
                uint8_t* Qs = Ps + kPPanels * kTnPanelBytes;
                const int row0 = (int)(mbeg + (int64_t)c * kTnRows);
                for (int pn = 0; pn < p.p_panels; ++pn) tma_load_2d(Ps + pn * kTnPanelBytes, &tmap_p, pn * 32, row0, &full_bar[s]);
                for (int pn = 0; pn < p.q_panels; ++pn) tma_load_2d(Qs + pn * kTnPanelBytes, &tmap_q, pn * 32, row0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0 && nchunks > 0) {
            const uint32_t idesc = make_idesc_tf32(128, p.Npad, 1, 1);      // both operands MN-major
            const uint32_t idesc1 = make_idesc_tf32(128, 16, 1, 1);
            const uint32_t d_main = tmem_base, d_ones = tmem_base + (uint32_t)p.Npad;
            const uint32_t ones_addr = smem_u32(ones);
            if (p.want_colsum) mbar_wait(&ones_bar, 0);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % p.stages;
                const uint32_t ph = (uint32_t)(c / p.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after_sync();
                const uint32_t pa = smem_u32(ring + (size_t)s * stage_bytes), qa = pa + kPPanels * kTnPanelBytes;
#pragma unroll
                for (int g = 0; g < kTnRows / 8; ++g) {
                    // SW128_32B MN-major: LBO = bytes between 32-feature blocks (panel stride), SBO = bytes between 4-row K groups
                    const uint64_t da = make_smem_desc(pa + g * 1024, kTnPanelBytes, 512, kLayoutSw128Base32);
                    const uint64_t db = make_smem_desc(qa + g * 1024, kTnPanelBytes, 512, kLayoutSw128Base32);
                    mma_tf32_ss(d_main, da, db, idesc, (c > 0 || g > 0) ? 1u : 0u);
                    if (p.want_colsum) {
                        const uint64_t d1 = make_smem_desc(ones_addr + g * 1024, kTnPanelBytes, 512, kLayoutSw128Base32);
                        mma_tf32_ss(d_ones, da, d1, idesc1, (c > 0 || g > 0) ? 1u : 0u);
                    }
                }
                mma_commit(&empty_bar[s]);
            }
            mma_commit(&done_bar);
        }
    } else {
        // ================================ epilogue warps ===============================
        const int et = t - 64;
        const int q4 = warp & 3;
        const int prow = q4 * 32 + lane;                                  // D row = M-side feature
        if (p.want_colsum) {                                              // ones tile: B(k, n) = (n == 0)
            for (int idx = et; idx < kTnRows * 8; idx += 128) {
                const int r = idx >> 3, q = idx & 7;
                *reinterpret_cast<float4*>(ones + mn32_chunk_offset(r, q, kTnRows)) = make_float4(q == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ones_bar);
        }
        if (nchunks > 0) {
            mbar_wait(&done_bar, 0);
            tc_fence_after_sync();
        }
        const int total = p.Ka * p.Kb;
        float* part = p.partial + (int64_t)blockIdx.x * (total + (p.want_colsum ? p.Kb : 0));
        const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int c0 = 0; c0 < p.Npad; c0 += 16) {
            float v[16];
            if (nchunks > 0) tmem_ld16(lane_base + (uint32_t)c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (prow >= p.Wp) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int qc = c0 + j;
                if (qc < p.Wq) {
                    if (p.p_is_b) part[(int64_t)qc * p.Kb + prow] = v[j];        // D[kb, ka] -> [ka][kb]
                    else part[(int64_t)prow * p.Kb + qc] = v[j];                 // D[ka, kb]
                }
            }
        }
        if (p.want_colsum) {
            float v[16];
            if (nchunks > 0) tmem_ld16(lane_base + (uint32_t)p.Npad, v);
            else v[0] = 0.f;
            if (prow < p.Wp) part[total + prow] = v[0];                         // colsum of the M-side operand (= B)
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// out / colsum = sum over S partials in ascending order (4 thread rows split S and are combined in a fixed order)
__global__ void __launch_bounds__(256)
tn_reduce_kernel(const float* __restrict__ partial, int S, int rows, int cols, int extra, float* __restrict__ out, int64_t ldo,
                 int transpose_out, float* __restrict__ out2) {
    __shared__ float red[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int total = rows * cols, stride = total + extra;
    const int i = blockIdx.x * 64 + tx;
    float s = 0.f;
    if (i < stride) {
        const int per = (S + 3) / 4;
        const int k0 = ty * per, k1 = min(S, k0 + per);
        for (int k = k0; k < k1; ++k) s += partial[(int64_t)k * stride + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < stride) {
        const float v = ((red[0][tx] + red[1][tx]) + red[2][tx]) + red[3][tx];
        if (i >= total) { out2[i - total] = v; return; }
        const int r = i / cols, c = i - r * cols;
        if (transpose_out) out[(int64_t)c * ldo + r] = v; else out[(int64_t)r * ldo + c] = v;
    }
}

static int tn_grid(int64_t M) {
    int64_t g = (M + kTnRows - 1) / kTnRows;
    if (g > kNumSMs) g = kNumSMs;
    return (int)(g < 1 ? 1 : g);
}

static bool tn_plan(int64_t Ka, int64_t Kb, int want_colsum, TcTnParams* p, size_t* smem) {
    // the operand whose column sums are wanted (B) must sit on the M side; otherwise take any operand <= 128 wide
    int p_is_b;
    if (want_colsum) { if (Kb > 128 || Ka > 256) return false; p_is_b = 1; }
    else if (Ka <= 128 && Kb <= 256) p_is_b = 0;
    else if (Kb <= 128 && Ka <= 256) p_is_b = 1;
    else return false;
    const int Wp = (int)(p_is_b ? Kb : Ka), Wq = (int)(p_is_b ? Ka : Kb);
    const int Npad = (Wq + 31) / 32 * 32;
    const int q_panels = Npad / 32, p_panels = (Wp + 31) / 32;
    const int stage = (kPPanels + q_panels) * kTnPanelBytes;
    int stages = (200 * 1024 - kTnPanelBytes) / stage;
    if (stages > kTnMaxStages) stages = kTnMaxStages;
    if (stages < 2) return false;
    if (p) {
        p->Wp = Wp; p->Wq = Wq; p->p_is_b = p_is_b; p->want_colsum = want_colsum; p->Npad = Npad;
        p->p_panels = p_panels; p->q_panels = q_panels; p->stages = stages;
        p->tmem_cols = (int)tmem_cols_pow2((uint32_t)(Npad + (want_colsum ? 32 : 0)));
        p->Ka = (int)Ka; p->Kb = (int)Kb;
    }
    if (smem) *smem = (size_t)stages * stage + kTnPanelBytes + 1024;
    return true;
}

bool tc_gemm_tn_eligible(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb, int want_colsum) {
    if (g_math_mode_get() == 0 || M < 1 || M >= ((int64_t)1 << 31)) return false;
    if ((lda & 3) || (ldb & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return false;
    return tn_plan(Ka, Kb, want_colsum, nullptr, nullptr);
}

size_t tc_gemm_tn_workspace(int64_t M, int64_t Ka, int64_t Kb, int want_colsum) {
    return sizeof(float) * (size_t)tn_grid(M) * (size_t)(Ka * Kb + (want_colsum ? Kb : 0));
}

int tc_gemm_tn_launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t Ka, int64_t Kb, float* out,
                      int64_t ldo, int transpose_out, float* colsum_b, void* workspace, cudaStream_t stream) {
    TcTnParams p;
    size_t smem = 0;
    GLAM_REQUIRE(tn_plan(Ka, Kb, colsum_b ? 1 : 0, &p, &smem), "tc_gemm_tn: shape not supported");
    p.M = M;
    const int grid = tn_grid(M);
    p.rows_per_cta = ((M + grid - 1) / grid + kTnRows - 1) / kTnRows * kTnRows;
    p.partial = (float*)workspace;
    const float* P = p.p_is_b ? B : A; const float* Q = p.p_is_b ? A : B;
    const int64_t ldp = p.p_is_b ? ldb : lda, ldq = p.p_is_b ? lda : ldb;
    CUtensorMap tp, tq;
    if (int rc = make_tmap_rows(&tp, P, M, p.Wp, ldp, kTnRows, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_tmap_rows(&tq, Q, M, p.Wq, ldq, kTnRows, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    {   // (per DEVICE attribute: set on every launch)
        cudaError_t e = ensure_dyn_smem((const void*)tc_gemm_tn_kernel, (size_t)((int)smem));
        if (e != cudaSuccess) { set_error("tc_gemm_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    tc_gemm_tn_kernel<<<grid, kTnThreads, smem, stream>>>(tp, tq, p);
    GLAM_CHECK_LAUNCH();
    const int extra = colsum_b ? (int)Kb : 0;
    const int total = (int)(Ka * Kb) + extra;
    (void)total;
    return launch_reduce_partials((const float*)workspace, grid, (int)Ka, (int)Kb, extra, out, ldo, transpose_out ? 1 : 0, colsum_b, stream);
}

}  // namespace glam
