// Blackwell (sm_100a) tensor-core plumbing: tcgen05 / TMEM / mbarrier wrappers and the UMMA descriptors used
// by the projection kernels.  Everything here is inline PTX; SASS shows UTCHMMA (tcgen05.mma), LDTM
// (tcgen05.ld), UTCBAR (tcgen05.commit).
//
// Shared-memory operand image ("panel" layout) — one image serves both operand majornesses:
//   a tile of R rows x F fp32 features is cut into panels of 32 features (128 bytes per row);
//   panel p, row r, 16-byte chunk c (4 floats) lives at   p*R*128 + r*128 + ((c ^ (r & 7)) * 16)
//   i.e. the canonical SWIZZLE_128B atom (8 rows x 128 B), rows dense, 1024-byte aligned panels.
//   * rows = M or N, features = K  -> K-major operand  (SBO = 1024 B between 8-row groups)
//   * rows = K, features = M or N  -> MN-major operand (LBO = panel stride, SBO = 1024 B between 8-k groups)
// TF32 operands are plain fp32 words: the tensor core ignores the 13 low mantissa bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glam {
namespace tc {

constexpr int kPanelFeatures = 32;           // fp32 features per 128-byte swizzle row
constexpr int kPanelRowBytes = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, feature f) inside a panelised tile with R rows
__device__ __forceinline__ uint32_t panel_offset(int r, int f, int R) {
    const int p = f >> 5, c = (f >> 2) & 7, e = f & 3;
    return (uint32_t)(p * R * kPanelRowBytes + r * kPanelRowBytes + (((c ^ (r & 7)) << 4) | (e << 2)));
}
// byte offset of the 16-byte chunk (row r, chunk index q = f/4) — q counts chunks over all panels
__device__ __forceinline__ uint32_t panel_chunk_offset(int r, int q, int R) {
    const int p = q >> 3, c = q & 7;
    return (uint32_t)(p * R * kPanelRowBytes + r * kPanelRowBytes + ((c ^ (r & 7)) << 4));
}

// MN-major TF32 operands have exactly one legal swizzled layout (CUTLASS sm100_common.inl: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): SWIZZLE_128B_BASE32B = rows (MMA K index) of 128 bytes holding 32
// MN-contiguous fp32, 32-byte chunks XOR-swizzled with (row & 3); 4-row groups are SBO apart (512 B when dense),
// 32-feature blocks LBO apart.  Offset of the 16-byte chunk q (over all panels) of row r:
__device__ __forceinline__ uint32_t mn32_chunk_offset(int r, int q, int R) {
    const int p = q >> 3, c = q & 7;
    return (uint32_t)(p * R * kPanelRowBytes + r * kPanelRowBytes + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
}

// ---- descriptors -----------------------------------------------------------------------------------
// Instruction descriptor, kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout):
//   [4,6) c_format = 1 (F32)  [7,10) a_format = 2 (TF32)  [10,13) b_format = 2  [15] a_major  [16] b_major
//   [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1 (Blackwell), [61,64) layout type (2 = SWIZZLE_128B).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = kLayoutSw128) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA (cp.async.bulk.tensor): box of a 2-D tensor map -> shared memory, completion on an mbarrier -------
// `map` points at a __grid_constant__ CUtensorMap kernel parameter; c0 = inner (feature) coordinate, c1 = row.
__device__ __forceinline__ void tma_load_2d(void* dst, const void* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- proxies / fences -----------------------------------------------------------------------------------
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------
// whole warp; ncols: power of two >= 32; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
    return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

// ---- MMA ------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem], one thread issues for the CTA
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: thread i of warp w reads TMEM lane 32*(w%4)+i, 16 consecutive fp32 columns ---------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 columns with a single wait (two x16 loads in flight)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr + 16u)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// six 4-column loads (addresses a0..a5) in flight behind ONE wait: the gate epilogue of the fused message kernel reads
// r|z|n of both GRU products for a 4-channel chunk.  One asm block, so no consumer can be scheduled before the wait.
__device__ __forceinline__ void tmem_ld4x6(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                           float (&v)[24]) {
    uint32_t r[24];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%26];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%12, %13, %14, %15}, [%27];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%16, %17, %18, %19}, [%28];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%20, %21, %22, %23}, [%29];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5)
        : "memory");
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i] = __uint_as_float(r[i]);
}

// mbarrier wait that traps instead of hanging the device when the expected arrival never comes (a lost tcgen05.commit
// would otherwise spin until the watchdog): ~2^22 failed probes are far beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0;; ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (it > (1u << 22)) __trap();
    }
}

}  // namespace tc
}  // namespace glam
