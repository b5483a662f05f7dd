// Blackwell (sm_100a) building blocks for the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) wrappers and UMMA descriptor builders.  Raw PTX, no CUTLASS.
#pragma once
#include <cuda.h>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------ host: tensor maps
// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint) so that the library has no
// link-time dependency on libcuda.so and still loads on a box without a driver.
typedef CUresult (*sh_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
sh_encode_tiled_fn sh_get_encode_tiled();

// bf16 tensor, `rank` dims (fastest first), 128-byte swizzle, zero fill out of bounds.
int sh_make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

// ------------------------------------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// ---- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {     // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane i of the warp <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- GroupNorm + ReLU applied to an operand tile IN shared memory (between the TMA arrival and the MMA): the tile a convolution
// (or its weight gradient) reads is x, and what it must multiply is a = relu(groupnorm(x)) -- this removes the separate
// gn_relu_fwd pass (read x, write a) and the tensor a itself.  Tile layout: rows of 128 B (one pixel's 64 channels), 128-byte
// swizzle (16-byte chunk j of row r sits at chunk j ^ (r & 7); the tile base is 1024-byte aligned).  A thread owns chunk j
// (channels c0 .. c0+7 of the 64-channel block) and walks rows, so its 8 scale / shift pairs are loop-invariant.
struct GnOperand {
    const float* stats;      // [N, groups, 2] (sum, sum of squares) of x, written by whichever kernel produced x
    const float* gamma;      // [C]
    const float* beta;       // [C]
    int groups;              // 0 = no fused GroupNorm
    int cpg;                 // channels per group (4, 8 or 16)
    float cnt_inv;           // 1 / (H * W * cpg)
    float eps;
};
// y = x * ka + kb with ka = rstd * gamma, kb = beta - mean * ka: the same expressions, in the same order, as gn_relu_fwd_kernel
__device__ __forceinline__ void gn_scale_shift(const GnOperand& q, int n, int c0, float (&ka)[8], float (&kb)[8]) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(q.gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(q.gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(q.beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(q.beta + c0 + 4));
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const int sh = 31 - __clz(q.cpg);                                         // cpg is a power of two
    const int gi0 = c0 >> sh, gi1 = (c0 + 7) >> sh;                           // 8 channels span one group (cpg >= 8) or two (cpg == 4)
    const float2 s0 = __ldg(reinterpret_cast<const float2*>(q.stats + ((size_t)n * q.groups + gi0) * 2));
    const float2 s1 = __ldg(reinterpret_cast<const float2*>(q.stats + ((size_t)n * q.groups + gi1) * 2));
    const float m0 = s0.x * q.cnt_inv, m1 = s1.x * q.cnt_inv;
    const float r0 = rsqrtf(fmaxf(s0.y * q.cnt_inv - m0 * m0, 0.f) + q.eps), r1 = rsqrtf(fmaxf(s1.y * q.cnt_inv - m1 * m1, 0.f) + q.eps);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const bool hi = ((c0 + e) >> sh) != gi0;
        const float mu = hi ? m1 : m0, rs = hi ? r1 : r0;
        ka[e] = rs * ga[e];
        kb[e] = fmaf(-mu, ka[e], be[e]);
    }
}
// Rows r, r+16, r+32, r+48 (r < 16) of chunk j of the tile at shared address `tile`: relu(bf16(x * ka + kb)), in place.  128 threads
// cover a 64-row slab this way (thread t: j = t & 7, r = t >> 3), so the time from "tile landed" to "tile transformed" is one batch
// of four shared-memory round trips: that latency is added to every trip of the TMA ring, whose depth bounds the kernel's bandwidth.
// Explicit shared-space 128-bit accesses, the four loads issued before the arithmetic, the ReLU folded into the bf16 conversion
// (round, then clamp == clamp, then round).
__device__ __forceinline__ void gn_xform_rows4(uint32_t tile, int r, int j, const float (&ka)[8], const float (&kb)[8]) {
    const uint32_t a0 = tile + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));                 // (r + 16 i) & 7 == r & 7
    uint32_t v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]) : "r"(a0 + u * 2048));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 y = __ffma2_rn(make_float2(__uint_as_float(v[u][e] << 16), __uint_as_float(v[u][e] & 0xffff0000u)),
                                        make_float2(ka[2 * e], ka[2 * e + 1]), make_float2(kb[2 * e], kb[2 * e + 1]));
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(v[u][e]) : "f"(y.y), "f"(y.x));
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0 + u * 2048), "r"(v[u][0]), "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3]));
    }
}

// ---- descriptors (bit layouts: PTX ISA "tcgen05 matrix/instruction descriptor"; cross-checked against
//      cute/arch/mma_sm100_desc.hpp in the vendored CUTLASS tree)
// K-major operand tile, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address       bits [0,14)
    d |= (uint64_t)1 << 16;                               // LBO (unused for swizzled K-major) bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                     // SBO = 1024 B        bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                               // layout: SWIZZLE_128B
    return d;
}
// Same, with the 8-row groups `sbo_bytes` apart (a multiple of 128) and a start address on ANY 128-byte row: the swizzle is a
// function of the shared-memory address bits, so a window into a larger TMA-written tile addresses correctly with
// base_offset 0 (measured: tools/umma_probe.cu).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major operand tile, 128-byte swizzle: each K index is one 128-byte row of 64 contiguous MN elements;
// 8 K-rows form a 1024 B group (SBO); successive 64-element MN chunks are `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// bf16 x bf16 -> fp32, M x N tile; a_mn / b_mn = 1 for MN-major operands
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#endif
