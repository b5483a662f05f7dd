// Tile geometry + shared-memory staging shared by the sphere-renderer kernels (R2 and the fused
// MutualProjectionLoss).  A warp owns a 32-lane tile of (tq lanes x trows rows), each lane PX pixels.
#pragma once
#include "common.cuh"

constexpr int kMaxJ = 64;
constexpr float kCullMargin = 1e-2f;   // mm; >> fp32 rounding of |c| + r at |c| <= 1e3

struct TileGeom {
    int px;          // pixels per lane (4 or 1)
    int tq;          // lanes per tile row
    int trows;       // rows per tile (= 32 / tq)
    int tiles_x;     // tiles per image row
    int tiles_y;
};

__host__ __device__ inline TileGeom make_geom(int W, int H, int px) {
    TileGeom g;
    g.px = px;
    int lanes_row = (W + px - 1) / px;
    int tq = 1;
    while (tq < lanes_row && tq < 8) tq <<= 1;
    g.tq = tq;
    g.trows = 32 / tq;
    g.tiles_x = (lanes_row + tq - 1) / tq;
    g.tiles_y = (H + g.trows - 1) / g.trows;
    return g;
}

// Stage the J spheres of image n into shared memory: TMA bulk copy (cp.async.bulk -> UBLKCP) + mbarrier.
__device__ __forceinline__ void stage_spheres(float4* s_sph, uint64_t* s_bar, const float4* __restrict__ g_sph, int J) {
    const uint32_t bytes = (uint32_t)J * 16u;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_sph);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(g_sph), "r"(bytes), "r"(bar) : "memory");
    }
    // all threads wait for phase 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar) : "memory");
}

// Warp-collective accumulation of one (sphere k, dcx, dcy, dcz, dr) contribution per lane into the warp's PRIVATE table:
// for every distinct sphere among the participating lanes the contributions are summed with shuffles and the elected lane
// does a plain read-modify-write -- no shared-memory float atomics (CAS loops, 32-way contended when a tile sits inside one
// sphere).  `has` may differ per lane; must be called by the whole warp.
__device__ __forceinline__ void warp_accumulate(float* __restrict__ acc_w, bool has, int k, float ax, float ay, float az, float ar,
                                                int lane) {
    unsigned todo = __ballot_sync(0xffffffffu, has);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int kk = __shfl_sync(0xffffffffu, k, leader);
        const bool mine = has && k == kk;
        float a = mine ? ax : 0.f, b = mine ? ay : 0.f, c = mine ? az : 0.f, d = mine ? ar : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (lane == leader) {
            float4* dst = reinterpret_cast<float4*>(acc_w) + kk;
            float4 v = *dst;
            v.x += a; v.y += b; v.z += c; v.w += d;
            *dst = v;
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}
