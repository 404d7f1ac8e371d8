// Projections on the 5th-generation tensor cores (tcgen05, TF32 operands, fp32 accumulation in TMEM).
//
//   Y[M,N] = epilogue(X[M,K] * W + bias)        M = nodes (10^5..10^8), K,N = 36..288
//
// The contraction is tiny per row, so the kernel is a streaming one: a persistent CTA keeps the whole weight
// matrix resident in shared memory (K-major, SWIZZLE_128B panels, staged once), then for every 128-row tile
// stages X with coalesced 16-byte loads into the swizzled operand image, issues ceil(K/8) tcgen05.mma from one
// thread into a TMEM accumulator (128 lanes x N columns), and drains it with tcgen05.ld straight into the
// epilogue (bias / CELU / CELU' mask / accumulate) and global stores.  Several CTAs share an SM so that one
// CTA's loads overlap another's MMA + epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace glam {

using namespace tc;

constexpr int kTcThreads = 128;
constexpr int kTileM = 128;

enum Epi { EPI_NONE = 0, EPI_CELU = 1, EPI_MUL_CELU_GRAD = 2, EPI_ACCUM = 3 };

struct TcGemmParams {
    const float* X; int64_t ldx;
    const float* W; int64_t w_sk, w_sn;
    const float* bias;
    const float* aux; int64_t ldaux;
    float* Y; int64_t ldy;
    int64_t M; int N, K, epi;
    int KP, Npad, tmem_cols, vec_store;
    int exact_begin, exact_end;   // output columns computed with exact fp32 FMAs (attention-logit columns)
};

__global__ void __launch_bounds__(kTcThreads)
tc_gemm_kernel(const TcGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int panels = (p.KP + kPanelFeatures - 1) / kPanelFeatures;
    uint8_t* Xs = smem;                                         // [panels][128][128 B]
    uint8_t* Ws = smem + (size_t)panels * kTileM * kPanelRowBytes;   // [panels][Npad][128 B]
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) { mbar_init(&mma_bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    // stage W once: Ws(n, k) = W[k*w_sk + n*w_sn], zero padded to [Npad][KP]
    for (int idx = t; idx < p.Npad * p.KP; idx += kTcThreads) {
        int n, k;
        if (p.w_sn == 1) { k = idx / p.Npad; n = idx - k * p.Npad; } else { n = idx / p.KP; k = idx - n * p.KP; }
        float v = (n < p.N && k < p.K) ? p.W[(int64_t)k * p.w_sk + (int64_t)n * p.w_sn] : 0.f;
        *reinterpret_cast<float*>(Ws + panel_offset(n, k, p.Npad)) = v;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(kTileM, p.Npad, 0, 0);
    const uint32_t xs_addr = smem_u32(Xs), ws_addr = smem_u32(Ws);
    const int KQ = p.K >> 2, KPQ = p.KP >> 2;
    const int64_t ntiles = (p.M + kTileM - 1) / kTileM;
    uint32_t phase = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m0 = tile * kTileM;
        // ---- stage X tile (coalesced 16 B loads -> swizzled 16 B stores)
        for (int idx = t; idx < kTileM * KPQ; idx += kTcThreads) {
            const int r = idx / KPQ, q = idx - r * KPQ;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < KQ && m0 + r < p.M) v = *reinterpret_cast<const float4*>(p.X + (m0 + r) * p.ldx + 4 * q);
            *reinterpret_cast<float4*>(Xs + panel_chunk_offset(r, q, kTileM)) = v;
        }
        fence_proxy_async_smem();
        __syncthreads();
        // ---- MMA: one thread, ceil(K/8) instructions, accumulator in TMEM
        if (t == 0) {
            tc_fence_after_sync();
            const int ksteps = p.KP >> 3;
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t pan = ks >> 2, within = (ks & 3) * 32;
                const uint64_t da = make_smem_desc(xs_addr + pan * (kTileM * kPanelRowBytes) + within, 16, 1024);
                const uint64_t db = make_smem_desc(ws_addr + pan * (p.Npad * kPanelRowBytes) + within, 16, 1024);
                mma_tf32_ss(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
            }
            mma_commit(&mma_bar);
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after_sync();
        // ---- epilogue: thread = row = TMEM lane
        const int64_t m = m0 + t;
        const bool row_ok = m < p.M;
        float* yrow = p.Y + m * p.ldy;
        const float* arow = p.aux ? p.aux + m * p.ldaux : nullptr;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < p.Npad; c0 += 16) {
            float v[16];
            tmem_ld16(lane_base + (uint32_t)c0, v);
            if (!row_ok) continue;
            if (c0 < p.exact_end && c0 + 16 > p.exact_begin) {
                // attention-logit columns: exact fp32 dot products from the staged operands (this thread's X row is
                // row t of the swizzled image; the W rows are broadcast reads)
                for (int n = max(c0, p.exact_begin); n < min(c0 + 16, p.exact_end); ++n) {
                    float acc = 0.f;
                    for (int q = 0; q < KQ; ++q) {
                        const float4 xv = *reinterpret_cast<const float4*>(Xs + panel_chunk_offset(t, q, kTileM));
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
                    float a = v[j];
                    if (p.bias) a += p.bias[n];
                    if (p.epi == EPI_CELU) a = celu1(a);
                    else if (p.epi == EPI_MUL_CELU_GRAD) { float y = arow[n]; a *= (y > 0.f ? 1.f : y + 1.f); }
                    else if (p.epi == EPI_ACCUM) a += yrow[n];
                    v[j] = a;
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
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    (void)lane;
}

static int g_math_mode = 1;   // 0 = fp32 on the CUDA cores (exact), 1 = TF32 tensor cores (default)
int g_math_mode_get() { return g_math_mode; }

bool tc_gemm_eligible(const float* X, int64_t ldx, const float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K) {
    if (g_math_mode == 0) return false;
    if (N > 256 || K > 288 || K < 4 || (K & 3) || (ldx & 3) || ((uintptr_t)X & 15)) return false;
    if (M < 1) return false;
    const int64_t panels = ((K + 7) / 8 * 8 + kPanelFeatures - 1) / kPanelFeatures;
    const int64_t smem = panels * (kTileM + (N + 15) / 16 * 16) * kPanelRowBytes + 1024;
    if (smem > 220 * 1024) return false;             // operands must fit next to each other in shared memory
    (void)Y; (void)ldy;
    return true;
}

int tc_gemm_launch(const float* X, int64_t ldx, const float* W, int64_t w_sk, int64_t w_sn, const float* bias,
                   const float* aux, int64_t ldaux, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K, int epi,
                   int exact_begin, int exact_end, cudaStream_t stream) {
    TcGemmParams p;
    p.exact_begin = exact_begin; p.exact_end = exact_end;
    p.X = X; p.ldx = ldx; p.W = W; p.w_sk = w_sk; p.w_sn = w_sn; p.bias = bias; p.aux = aux; p.ldaux = ldaux;
    p.Y = Y; p.ldy = ldy; p.M = M; p.N = (int)N; p.K = (int)K; p.epi = epi;
    p.KP = (int)((K + 7) / 8 * 8);
    p.Npad = (int)((N + 15) / 16 * 16);
    p.tmem_cols = (int)tmem_cols_pow2((uint32_t)p.Npad);
    p.vec_store = ((ldy & 3) == 0 && ((uintptr_t)Y & 15) == 0) ? 1 : 0;
    const int panels = (p.KP + kPanelFeatures - 1) / kPanelFeatures;
    const size_t smem = (size_t)panels * (kTileM + p.Npad) * kPanelRowBytes + 1024;
    GLAM_REQUIRE(smem <= 220 * 1024, "tc_gemm: operands do not fit in shared memory (N=%lld K=%lld)", (long long)N, (long long)K);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("tc_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = smem;
    }
    // CTAs per SM limited by shared memory and by TMEM columns (512 per SM)
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    const int by_tmem = 512 / p.tmem_cols;
    if (per_sm > by_tmem) per_sm = by_tmem;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (M + kTileM - 1) / kTileM;
    int64_t grid = (int64_t)kNumSMs * per_sm;
    if (grid > ntiles) grid = ntiles;
    tc_gemm_kernel<<<(unsigned)grid, kTcThreads, smem, stream>>>(p);
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
