// Library-level entry points: version, error string, launch counter.
#include <map>
#include <mutex>
#include <utility>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace glam {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace glam

namespace glam {
__global__ void __launch_bounds__(256)
reduce_partials_fixed_kernel(const float* __restrict__ partial, int S, int rows, int cols, int extra, float* __restrict__ out,
                             int64_t ldo, int transpose_out, float* __restrict__ out2) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int total = rows * cols, stride = total + extra;
    const int i = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (i < stride) {
        const int per = (S + 7) / 8, k0 = ty * per, k1 = min(S, k0 + per);
        int k = k0;
        for (; k + 4 <= k1; k += 4) {
            const float a = partial[(int64_t)k * stride + i], b = partial[(int64_t)(k + 1) * stride + i];
            const float c = partial[(int64_t)(k + 2) * stride + i], d = partial[(int64_t)(k + 3) * stride + i];
            s = (((s + a) + b) + c) + d;
        }
        for (; k < k1; ++k) s += partial[(int64_t)k * stride + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < stride) {
        float v = red[0][tx];
#pragma unroll
        for (int q = 1; q < 8; ++q) v += red[q][tx];
        if (i >= total) { out2[i - total] = v; return; }
        const int r = i / cols, c = i - r * cols;
        if (transpose_out) out[(int64_t)c * ldo + r] = v; else out[(int64_t)r * ldo + c] = v;
    }
}
cudaError_t ensure_dyn_smem(const void* fn, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> have;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = have[{fn, dev}];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}
int launch_reduce_partials(const float* partial, int S, int rows, int cols, int extra, float* out, int64_t ldo, int transpose_out,
                           float* out2, cudaStream_t stream) {
    const int n = rows * cols + extra;
    reduce_partials_fixed_kernel<<<(n + 31) / 32, 256, 0, stream>>>(partial, S, rows, cols, extra, out, ldo, transpose_out, out2);
    GLAM_CHECK_LAUNCH();
    return 0;
}
}  // namespace glam

extern "C" int glam_abi_version(void) { return GLAM_B200_ABI_VERSION; }
extern "C" const char* glam_last_error(void) { return glam::g_err; }
extern "C" int64_t glam_launch_count(void) { return glam::g_launches.load(std::memory_order_relaxed); }
