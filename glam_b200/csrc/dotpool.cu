// Cross-graph interaction pool dot_and_global_pool2 (see include/glam_b200.h (6)).
//
// The reference loops over pairs on the host with 4 .item() syncs per pair (src_2gi_ddi/layer.py:275-282);
// here one CTA owns one pair, tiles S = Xa Xb^T through shared memory 32x32 at a time and keeps only the
// running max/argmax.  mean(S) needs no product at all: <sum_a Xa, sum_b Xb> / (na*nb).
// Backward of the full-reduction max routes to the first arg-max in row-major order (ties have measure zero
// for real activations; the oracle test pins this choice).
#include "common.cuh"

namespace glam {

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

    // column sums (fixed row order) and the mean
    float dotp = 0.f;
    for (int k = t; k < C; k += kPairThreads) {
        float sa = 0.f, sb = 0.f;
        for (int a = a0; a < a1; ++a) sa += xa[(int64_t)a * C + k];
        for (int b = b0; b < b1; ++b) sb += xb[(int64_t)b * C + k];
        sum_a[(int64_t)g * C + k] = sa;
        sum_b[(int64_t)g * C + k] = sb;
        dotp = fmaf(sa, sb, dotp);
    }
    dotp = warp_sum(dotp);
    if (lane == 0) red_dot[wid] = dotp;

    float best = -INFINITY;
    long long best_i = 0x7fffffffffffffffLL;
    const int bi = t & 31;
    for (int at = 0; at < na; at += kPairTile) {
        __syncthreads();
        for (int idx = t; idx < kPairTile * C; idx += kPairThreads) {
            int r = idx / C, k = idx - r * C;
            As[r * ld + k] = (at + r < na) ? xa[(int64_t)(a0 + at + r) * C + k] : 0.f;
        }
        for (int bt = 0; bt < nb; bt += kPairTile) {
            __syncthreads();
            for (int idx = t; idx < kPairTile * C; idx += kPairThreads) {
                int r = idx / C, k = idx - r * C;
                Bs[r * ld + k] = (bt + r < nb) ? xb[(int64_t)(b0 + bt + r) * C + k] : 0.f;
            }
            __syncthreads();
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
        cudaFuncSetAttribute(pair_dot_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
        cudaFuncSetAttribute(pair_dot_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pair_dot_pool_fwd_kernel<<<(unsigned)num_pairs, kPairThreads, smem, (cudaStream_t)stream_>>>(xa, xb, ptr_a, ptr_b, idx_b, C, out,
                                                                                               argmax, sum_a, sum_b);
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
