// Device helpers shared by the fused message-stack kernels (mp_fused.cu forward, mp_fused_bwd.cu backward): tile constants,
// shared-memory vector accesses, SFU / packed-pair fp32 math, TMEM accesses, the panel swizzle.
#pragma once
#include <cuda.h>
#include <math_constants.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace glam {
namespace mp {

using namespace tc;

constexpr int kMpM = 128;                    // rows per tile = UMMA M
constexpr int kMpMaxEdges = 768;             // in-edges per tile (molecular tiles: ~2.2 per atom)
constexpr int kMpMaxDe = 4;                  // bond types
constexpr int kMpMaxRaw = 16;                // raw atom features of the optional input LinearBlock
constexpr int kMpPanel = kMpM * kPanelRowBytes;      // 16 KB

__device__ __forceinline__ float4 lds128(const void* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts128(void* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// gate non-linearities straight on the SFU: ex2.approx / rcp.approx (relative error ~2^-22 each), 4 instructions per
// sigmoid — the gate epilogue was 31 % of the kernel's instructions with the library forms
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, rcp_approx(1.f + ex2_approx(-2.8853900817779268f * x)), -1.f); }
// CELU(alpha = 1) with exp from ex2.approx: absolute error <= 2e-7 (common.cuh's celu1 costs ~30 instructions per element)
__device__ __forceinline__ float celu_fast(float x) { return x > 0.f ? x : ex2_approx(1.4426950408889634f * x) - 1.f; }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100a): per lane the same IEEE round-to-nearest result as the scalar forms
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 sigmoid2(float2 x) {
    const float2 t = mul2(x, f2(-1.4426950408889634f));
    const float2 d = add2(f2(ex2_approx(t.x), ex2_approx(t.y)), f2(1.f));
    return f2(rcp_approx(d.x), rcp_approx(d.y));
}
__device__ __forceinline__ float2 tanh2(float2 x) {
    const float2 t = mul2(x, f2(-2.8853900817779268f));
    const float2 d = add2(f2(ex2_approx(t.x), ex2_approx(t.y)), f2(1.f));
    return fma2(f2(2.f), f2(rcp_approx(d.x), rcp_approx(d.y)), f2(-1.f));
}
__device__ __forceinline__ float2 celu2(float2 x) {
    const float2 t = mul2(x, f2(1.4426950408889634f));
    const float2 e = add2(f2(ex2_approx(t.x), ex2_approx(t.y)), f2(-1.f));
    return f2(x.x > 0.f ? x.x : e.x, x.y > 0.f ? x.y : e.y);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of the 16-byte chunk q (4 features) of row r inside a 128-byte-row SWIZZLE_128B panel
__device__ __forceinline__ uint32_t pan_off(int r, int q) { return (uint32_t)(r * kPanelRowBytes + (((q & 7) ^ (r & 7)) << 4)); }
// TMEM <- registers: thread i of warp w writes TMEM lane 32*(w%4)+i, 4 consecutive fp32 columns (the mirror of tmem_ld4)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16-byte global -> shared copy without a register round trip (LDGSTS); completion by cp_async_wait_all
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// linear bulk copy global -> shared by the copy engine (UBLKCP), bytes % 16 == 0, completion on an mbarrier (expect_tx first)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace mp
}  // namespace glam
