// Hourglass convolutions, forward and data-gradient: persistent tcgen05 implicit-GEMM kernel (sm_100a).
//
// Replaces the cuDNN/ATen kernels behind every stride-1 nn.Conv2d of the reference network
//   Bottleneck.conv1/conv2/conv3, downsample          /root/reference/network/hourglass.py:13-18, 125-128
//   HourglassNet.fc / score / fc_ / score_            /root/reference/network/hourglass.py:110-114, 138-145
// (the data gradient is the same launch on dY with the flipped/transposed weights of sh_pack_weights).
//
// Formulation: D[128 pixels, BN] += A[128 pixels, 64] * B[BN, 64]^T per (tap, 64-channel block); NHWC bf16 in HBM,
// fp32 accumulators in TMEM.  A tile = one TMA box {64 ch, bw, bh, bn} of the 4-D activation tensor at spatial offset
// (kw-1, kh-1): TMA's out-of-bounds zero fill IS the convolution padding (im2col staging without an im2col buffer).
//
// What bounds these layers on a B200 is bytes moved per pixel tile, not MMA issue, so the kernel is organised around
// moving each byte once:
//   * persistent: one CTA per SM walks pixel tiles (static round-robin); barriers / TMEM / descriptors set up once;
//   * BN = ALL output channels (<= 256): the activation tile is fetched once, not once per 128-channel slice;
//   * weights RESIDENT in shared memory whenever taps*Cin*BN*2 B fits (every 1x1 layer, the 64-channel 3x3): loaded
//     once per CTA instead of once per tile; otherwise streamed next to the A tile in the same pipeline stage;
//   * two TMEM accumulator buffers (2*BN columns): the epilogue of tile i overlaps the main loop of tile i+1;
//   * epilogue through shared memory: the residual tile arrives by TMA (prefetched one 64-channel chunk ahead), each
//     thread owns one pixel row (tcgen05.ld 32x32b), adds bias + residual, rounds to bf16, accumulates the GroupNorm
//     statistics of the ROUNDED values (butterfly warp reduction -> one atomic per value), writes the swizzled staging
//     tile in place and one elected thread issues a TMA store; a ring of staging slots lets loads / math / stores overlap;
//   * FOUR epilogue warp-groups: the epilogue (TMEM load, bias, residual, rounding, statistics, staging) is what paces the
//     HBM-bound 1x1 layers (ncu: the MMA and the TMA ring wait on it), and a warp cannot hide its own dependent-issue
//     latencies; group (b, h) serves accumulator buffer b (alternate tiles) and the 64-channel chunks c = h, h+2, ...
//     of it, 32 columns at a time (register budget of a 576-thread CTA), with its own staging slot.
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..17 = epilogue groups 0..3 (4 warps each, one TMEM lane quadrant per warp).  The XF build (sh_conv_fwd_gn, 704 threads)
// adds warps 18..21 = operand transform: the fused GroupNorm + ReLU of a 1x1 layer's input, applied to every A tile in shared
// memory between the TMA arrival and the MMA.
#include "tc_common.cuh"

namespace {

constexpr int kBM = 128;           // pixels per tile = TMEM lanes
constexpr int kBK = 64;            // channels per k-block = one 128-byte swizzle row
constexpr int kThreads = 576;       // producer + MMA + 16 epilogue warps
constexpr int kThreadsXf = 704;     // + 4 operand-transform warps (fused GroupNorm of a 1x1 layer's input): the XF build
constexpr int kMaxStages = 8;
constexpr int kGroups = 4;         // epilogue warp-groups, one 16 KB staging slot each
constexpr int kABytes = kBM * kBK * 2;          // 16 KB
constexpr int kStgBytes = kBM * 64 * 2;         // 16 KB: 128 pixels x 64 channels bf16
constexpr int kSmemLimit = 232448;              // 227 KB opt-in maximum per CTA
constexpr int kIdentBytes = 64 * 128;           // K-major 64x64 identity tile
// ---- halo mode (3x3, images >= 16x8): ONE TMA box {64 ch, 10, 18} per (tile, k-block) instead of nine 16 KB tap boxes; the
// nine taps are UMMA descriptors into that halo tile (start shifted by (dh+1)*10 + (dw+1) rows of 128 B, 8-row groups one
// halo row = 1280 B apart: tools/umma_probe.cu), and every weight tile is used for TWO pixel tiles (four TMEM accumulators),
// so L2 -> SM traffic per 128 pixels drops from 576 KB to 190 KB at Cin = Cout = 128.
constexpr int kHaloW = 10, kHaloH = 18;         // tile 8 x 16 pixels + 1-pixel border
constexpr int kHaloBytes = kHaloW * kHaloH * 128;            // 23,040 B
constexpr int kHaloSlotBytes = (kHaloBytes + 1023) / 1024 * 1024;
constexpr int kHaloSlots = 4;                   // two tiles of the current k-block + two of the next

struct ConvGeom {
    int N, H, W;
    int bw, bh, bn;                // TMA box (pixels): bw*bh*bn == 128
    int tiles_w, tiles_h, num_tiles;
    int taps;                      // 1 or 9
    int kblocks;                   // Cin / 64
    int cout;                      // real output channels
    int cout_pad;                  // rows per tap in the weight matrix (== BN)
    int y_ld;                      // channel stride of the bf16 output (0 = no bf16 output)
    int groups;                    // GroupNorm groups of the statistics side output (0 = none)
    int stages;                    // main-loop pipeline depth
    int b_resident;                // weights loaded once per CTA
    int has_res;                   // residual tile fetched by TMA
    int nchunks;                   // 64-channel output chunks the epilogue processes
    int fuse_gn;                   // A tiles are x; the MMA multiplies relu(groupnorm(x)) (1x1 layers, <= 2 images per tile)
};

struct ConvPtrs {
    const float* bias;             // [cout] or null
    float* y_nchw;                 // [N,cout,H,W] fp32 or null
    float* stats;                  // [N,groups,2] fp32 (sum, sum of squares), accumulated atomically
    GnOperand gn;                  // fused GroupNorm of the INPUT (groups == 0: none)
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

struct TileCoord { int n0, h0, w0; };
__device__ __forceinline__ TileCoord tile_coord(const ConvGeom& g, int tile) {
    TileCoord t;
    const int tw = tile % g.tiles_w; tile /= g.tiles_w;
    const int th = tile % g.tiles_h; tile /= g.tiles_h;
    t.n0 = tile * g.bn; t.h0 = th * g.bh; t.w0 = tw * g.bw;
    return t;
}

// Butterfly (recursive-halving) warp reduction of V per-lane values over the R lanes that share an image: each split
// step halves the values a lane is responsible for, so V values cost V-1 (+ log2(R/V)) shuffles instead of V*log2(R).
template <int N, int O>
struct Bfly {
    static __device__ __forceinline__ void run(float* v, int lane, int& base) {
        if constexpr (O >= 1) {
            if constexpr (N > 1) {
                constexpr int half = N / 2;
                const bool upper = (lane & O) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const float send = upper ? v[i] : v[i + half];
                    const float keep = upper ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
                }
                if (upper) base += half;
                Bfly<half, O / 2>::run(v, lane, base);
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], O);
                Bfly<1, O / 2>::run(v, lane, base);
            }
        }
    }
};
__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x / 2); }

// Per-(sample, group) sum / sum of squares of one thread row's 2*NW rounded channels (NW packed bf16 pairs), reduced over
// the R lanes of the warp that belong to the same image; the lanes left owning a value issue one atomic each.
template <int GS, int R, int NW>
__device__ __forceinline__ void row_stats(const uint32_t (&packed)[NW], bool row_ok, int lane, int cg, int cout,
                                          float* __restrict__ stats_n) {
    constexpr int NG = 2 * NW / GS;                        // groups in this slice
    constexpr int V = 2 * NG;
    float vals[V];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);      // packed f32x2: one issue slot per channel pair
#pragma unroll
        for (int j = 0; j < GS / 2; ++j) {
            const uint32_t pk = packed[gi * (GS / 2) + j];
            const float2 xp = make_float2(__uint_as_float(pk << 16), __uint_as_float(pk & 0xffff0000u));
            s1 = __fadd2_rn(s1, xp);
            s2 = __ffma2_rn(xp, xp, s2);
        }
        vals[2 * gi] = row_ok ? s1.x + s1.y : 0.f;
        vals[2 * gi + 1] = row_ok ? s2.x + s2.y : 0.f;
    }
    int base = 0;
    Bfly<V, R / 2>::run(vals, lane, base);
    constexpr int split = ilog2(V) < ilog2(R) ? ilog2(V) : ilog2(R);
    constexpr int left = V >> split;                       // values a lane still owns
    constexpr int plain = ilog2(R) - split;                // trailing all-reduce steps: only lanes with those bits 0 write
    if ((lane & ((1 << plain) - 1)) == 0 && row_ok) {
#pragma unroll
        for (int j = 0; j < left; ++j) {
            const int idx = base + j;                       // = 2 * group-in-slice + {0: sum, 1: sum of squares}
            if (cg + (idx >> 1) * GS < cout) atomicAdd(stats_n + (cg / GS) * 2 + idx, vals[j]);
        }
    }
}

template <int BN, bool HALO, bool XF>
__global__ void __launch_bounds__(XF ? kThreadsXf : kThreads, 1) conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmRes,
                                                               const __grid_constant__ CUtensorMap tmOut,
                                                               const ConvGeom g, const ConvPtrs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kBBytes = BN * kBK * 2;                       // one (tap, k-block) weight tile
    constexpr int NACC = HALO ? 4 : 2;                          // TMEM accumulator buffers
    const int num_k = g.taps * g.kblocks;
    // halo mode: the ring holds weight tiles only (the activation halos have their own slots in front of it)
    const int stage_bytes = HALO ? kBBytes : kABytes + (g.b_resident ? 0 : kBBytes);
    uint8_t* s_halo = smem;
    uint8_t* s_pipe = smem + (HALO ? kHaloSlots * kHaloSlotBytes : 0);
    uint8_t* s_bres = s_pipe + g.stages * stage_bytes;          // resident weights (size 0 when streamed)
    uint8_t* s_ident = s_bres + (g.b_resident ? num_k * kBBytes : 0);   // 64x64 bf16 identity (residual add on the tensor core)
    uint8_t* s_stg = s_ident + (g.has_res ? kIdentBytes : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stg + kGroups * kStgBytes);
    uint64_t* full_bar = bars;                                  // [kMaxStages]
    uint64_t* empty_bar = bars + kMaxStages;                    // [kMaxStages]
    uint64_t* b_full = bars + 2 * kMaxStages;                   // [1]
    uint64_t* acc_full = b_full + 1;                            // [4]
    uint64_t* acc_empty = acc_full + 4;                         // [4]
    uint64_t* halo_full = acc_empty + 4;                        // [kHaloSlots]
    uint64_t* halo_empty = halo_full + kHaloSlots;              // [kHaloSlots]
    uint64_t* xf_bar = halo_empty + kHaloSlots;                 // [kMaxStages]: the A tile of the stage has been transformed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf_bar + kMaxStages);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);    // [256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (g.has_res) tma_prefetch_desc(&tmRes);
        if (g.y_ld) tma_prefetch_desc(&tmOut);
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(b_full, 1);
        // a buffer is drained by one group (single-chunk layers) or by the two groups that split its chunks
        for (int b = 0; b < 4; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], g.nchunks >= 2 ? 256 : 128); }
        for (int s = 0; s < kHaloSlots; ++s) { mbar_init(&halo_full[s], 1); mbar_init(&halo_empty[s], 1); }
        for (int s = 0; s < kMaxStages; ++s) mbar_init(&xf_bar[s], 128);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = (p.bias && i < g.cout) ? p.bias[i] : 0.f;
    if (g.has_res) {
        // I[n][k] = (n == k), K-major rows of 128 B with the 128-byte swizzle: 16-byte chunk j of row n sits at chunk j ^ (n & 7)
        for (int i = threadIdx.x; i < kIdentBytes / 16; i += blockDim.x) {
            const int n = i >> 3, j = i & 7;                    // row, logical chunk (channels 8j .. 8j+7)
            uint32_t w4[4] = {0u, 0u, 0u, 0u};
            if ((n >> 3) == j) w4[(n & 7) >> 1] = (n & 1) ? 0x3f800000u : 0x00003f80u;      // bf16 1.0 at element n % 8
            *reinterpret_cast<uint4*>(s_ident + n * 128 + ((j ^ (n & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        fence_async_smem();                                     // generic-proxy writes -> visible to tcgen05.mma
    }
    if (warp == 1) tmem_alloc(tmem_slot, NACC * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one elected lane) =====================
        if constexpr (HALO) {
            if (lane == 0) {
                // item sequence of this CTA: super-tiles of two pixel tiles; per k-block two halos and nine weight tiles.
                const int my_tiles = (g.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
                const int n_items = ((my_tiles + 1) / 2) * g.kblocks;
                int hit = 0;                                        // halo slots issued so far
                auto issue_halos = [&](int item) {
                    const int st = item / g.kblocks, kb = item - st * g.kblocks;
                    for (int sub = 0; sub < 2; ++sub) {
                        const int tl = 2 * st + sub;
                        if (tl >= my_tiles) break;
                        const TileCoord t = tile_coord(g, blockIdx.x + tl * gridDim.x);
                        const int hs = hit % kHaloSlots, hph = (hit / kHaloSlots) & 1;
                        ++hit;
                        mbar_wait(&halo_empty[hs], hph ^ 1);
                        mbar_expect_tx(&halo_full[hs], (uint32_t)kHaloBytes);
                        tma_load_4d(s_halo + hs * kHaloSlotBytes, &tmA, &halo_full[hs], kb * kBK, t.w0 - 1, t.h0 - 1, t.n0);
                    }
                };
                if (n_items > 0) issue_halos(0);
                int it = 0;
                for (int item = 0; item < n_items; ++item) {
                    const int kb = item % g.kblocks;
                    for (int tap = 0; tap < 9; ++tap, ++it) {
                        // the next item's halos go out after four weight tiles of this one: by then (ring of >= 4 slots) the
                        // MMA has started this item, i.e. the slots of item-1 are free and the wait below cannot hold up the
                        // weight stream; they then have five taps (~2500 MMA cycles) to land
                        if (tap == 4 && item + 1 < n_items) issue_halos(item + 1);
                        const int s = it % g.stages, ph = (it / g.stages) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_expect_tx(&full_bar[s], (uint32_t)kBBytes);
                        tma_load_2d(s_pipe + s * stage_bytes, &tmB, &full_bar[s], kb * kBK, tap * g.cout_pad);
                    }
                }
            }
        } else
        if (lane == 0) {
            if (g.b_resident) {
                mbar_expect_tx(b_full, (uint32_t)(num_k * kBBytes));
                for (int k = 0; k < num_k; ++k) {
                    const int tap = k / g.kblocks, kb = k - tap * g.kblocks;
                    tma_load_2d(s_bres + k * kBBytes, &tmB, b_full, kb * kBK, tap * g.cout_pad);
                }
            }
            int it = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord(g, tile);
                for (int k = 0; k < num_k; ++k, ++it) {
                    const int s = it % g.stages, ph = (it / g.stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int tap = k / g.kblocks, kb = k - tap * g.kblocks;
                    const int dh = g.taps == 9 ? tap / 3 - 1 : 0, dw = g.taps == 9 ? tap % 3 - 1 : 0;
                    uint8_t* a_dst = s_pipe + s * stage_bytes;
                    mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    tma_load_4d(a_dst, &tmA, &full_bar[s], kb * kBK, t.w0 + dw, t.h0 + dh, t.n0);
                    if (!g.b_resident) tma_load_2d(a_dst + kABytes, &tmB, &full_bar[s], kb * kBK, tap * g.cout_pad);
                }
                if (g.has_res) {
                    // the residual tile rides the same ring: one more "k-block" per 64 output channels, multiplied by I
                    for (int c = 0; c < g.nchunks; ++c, ++it) {
                        const int s = it % g.stages, ph = (it / g.stages) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_expect_tx(&full_bar[s], (uint32_t)kABytes);
                        tma_load_4d(s_pipe + s * stage_bytes, &tmRes, &full_bar[s], c * 64, t.w0, t.h0, t.n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        if constexpr (HALO) {
            if (lane == 0) {
                constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
                const int my_tiles = (g.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
                int it = 0, hit = 0;
                for (int tl0 = 0; tl0 < my_tiles; tl0 += 2) {
                    const int nsub = tl0 + 1 < my_tiles ? 2 : 1;
                    for (int sub = 0; sub < nsub; ++sub)            // the epilogue has drained these accumulators
                        mbar_wait(&acc_empty[(tl0 + sub) & 3], (uint32_t)((((tl0 + sub) >> 2) & 1) ^ 1));
                    tc_fence_after();
                    for (int kb = 0; kb < g.kblocks; ++kb) {
                        uint32_t halo_addr[2];
                        int hslot[2];
                        for (int sub = 0; sub < nsub; ++sub, ++hit) {
                            hslot[sub] = hit % kHaloSlots;
                            mbar_wait(&halo_full[hslot[sub]], (uint32_t)((hit / kHaloSlots) & 1));
                            halo_addr[sub] = smem_u32(s_halo + hslot[sub] * kHaloSlotBytes);
                        }
                        tc_fence_after();
                        for (int tap = 0; tap < 9; ++tap, ++it) {
                            const int s = it % g.stages, ph = (it / g.stages) & 1;
                            mbar_wait(&full_bar[s], ph);
                            tc_fence_after();
                            const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(s_pipe + s * stage_bytes));
                            const uint32_t shift = (uint32_t)((tap / 3) * kHaloW + tap % 3) * 128u;     // (dh+1)*10 + (dw+1) rows
                            for (int sub = 0; sub < nsub; ++sub) {
                                const uint64_t ad = umma_desc_kmajor_sw128_sbo(halo_addr[sub] + shift, kHaloW * 128);
                                const uint32_t d_tmem = tmem_base + (uint32_t)(((tl0 + sub) & 3) * BN);
#pragma unroll
                                for (int kk = 0; kk < kBK / 16; ++kk)
                                    umma_bf16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kb | tap | kk) != 0);
                            }
                            umma_commit(&empty_bar[s]);             // frees the weight slot when these MMAs retire
                        }
                        for (int sub = 0; sub < nsub; ++sub) umma_commit(&halo_empty[hslot[sub]]);
                    }
                    for (int sub = 0; sub < nsub; ++sub) umma_commit(&acc_full[(tl0 + sub) & 3]);
                }
            }
        } else
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
            if (g.b_resident) mbar_wait(b_full, 0);
            int it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++tcount) {
                const int buf = tcount & 1, aph = (tcount >> 1) & 1;
                mbar_wait(&acc_empty[buf], aph ^ 1);            // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int k = 0; k < num_k; ++k, ++it) {
                    const int s = it % g.stages, ph = (it / g.stages) & 1;
                    mbar_wait(XF ? &xf_bar[s] : &full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(s_pipe + s * stage_bytes);
                    const uint32_t b_addr = g.b_resident ? smem_u32(s_bres + k * kBBytes) : a_addr + kABytes;
                    const uint64_t ad = umma_desc_kmajor_sw128(a_addr);
                    const uint64_t bd = umma_desc_kmajor_sw128(b_addr);
#pragma unroll
                    for (int kk = 0; kk < kBK / 16; ++kk)       // +32 B inside the 128-byte swizzle row per K=16 step
                        umma_bf16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
                    umma_commit(&empty_bar[s]);                 // frees the smem slot when these MMAs retire
                }
                if (g.has_res) {
                    // D[:, 64c .. 64c+63] += R_c * I : bf16 residual values enter the fp32 accumulator exactly
                    constexpr uint32_t idesc64 = umma_idesc_bf16(kBM, 64, 0, 0);
                    const uint64_t bd = umma_desc_kmajor_sw128(smem_u32(s_ident));
                    for (int c = 0; c < g.nchunks; ++c, ++it) {
                        const int s = it % g.stages, ph = (it / g.stages) & 1;
                        mbar_wait(XF ? &xf_bar[s] : &full_bar[s], ph);
                        tc_fence_after();
                        const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(s_pipe + s * stage_bytes));
#pragma unroll
                        for (int kk = 0; kk < kBK / 16; ++kk)
                            umma_bf16(d_tmem + (uint32_t)(c * 64), ad + 2 * kk, bd + 2 * kk, idesc64, 1u);
                        umma_commit(&empty_bar[s]);
                    }
                }
                umma_commit(&acc_full[buf]);                    // accumulator complete
            }
        }
    } else if (XF && warp >= 18) {
        // ===================== operand transform: A := relu(groupnorm(x)) in shared memory =====================
        if constexpr (XF) {
            // All four warps work on the same stage: what matters is the LATENCY from "tile landed" to "tile transformed" (it is
            // added to every trip of the TMA ring, whose depth bounds the bandwidth of these layers), not the warps' throughput.
            const int tt = (int)threadIdx.x - 18 * 32;          // 0..127
            const int j = tt & 7, r = tt >> 3;                  // 16-byte chunk of the 128-byte rows; rows r + 16 i
            int it = 0;
            for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
                const TileCoord t = tile_coord(g, tile);
                const int nA = min(t.n0, g.N - 1), nB = min(t.n0 + g.bn - 1, g.N - 1);   // rows 0..63 / 64..127 (bn <= 2)
                for (int k = 0; k < num_k; ++k, ++it) {         // taps == 1: k is the 64-channel block
                    const int s = it % g.stages, ph = (it / g.stages) & 1;
                    float ka[8], kb[8];
                    gn_scale_shift(p.gn, nA, k * kBK + j * 8, ka, kb);                   // (global loads: issued before the wait)
                    mbar_wait(&full_bar[s], ph);
                    const uint32_t tileA = smem_u32(s_pipe + s * stage_bytes);
                    gn_xform_rows4(tileA, r, j, ka, kb);                                 // rows 0..63   (image nA)
                    if (nB != nA) gn_scale_shift(p.gn, nB, k * kBK + j * 8, ka, kb);
                    gn_xform_rows4(tileA + 64 * 128, r, j, ka, kb);                      // rows 64..127 (image nB)
                    fence_async_smem();                         // generic-proxy writes -> visible to tcgen05.mma
                    mbar_arrive(&xf_bar[s]);
                }
                if (g.has_res) {                                // the residual tiles ride the same ring untransformed
                    for (int c = 0; c < g.nchunks; ++c, ++it) {
                        const int s = it % g.stages, ph = (it / g.stages) & 1;
                        // every transform thread meets every phase of every full barrier, in order: a thread that skipped these
                        // could reach its next wait on a barrier two phases early and pass on the stale parity
                        mbar_wait(&full_bar[s], ph);
                        mbar_arrive(&xf_bar[s]);
                    }
                }
            }
        }
    } else {
        // ===================== epilogue warp-groups =====================
        const int egrp = (warp - 2) >> 2;                       // 0..3
        const int buf = egrp & 1;                               // accumulator buffer / tile parity this group serves
        const int half = egrp >> 1;                             // which chunks of the tile: c = half, half + 2, ...
        const int quad = warp & 3;                              // TMEM lane quadrant this warp may read
        const int m = quad * 32 + lane;                         // tile row = pixel (box order: w fastest, then h, n)
        const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);   // issues this group's residual loads / output stores
        const int px_per_img = g.bw * g.bh;
        const int ln = m / px_per_img, lh = (m / g.bw) % g.bh, lw = m % g.bw;
        const int gs = g.groups > 0 ? g.cout / g.groups : 64;   // channels per group
        const uint32_t row_off = (uint32_t)m * 128u;
        const uint32_t sw = (uint32_t)(m & 7);
        uint8_t* stg = s_stg + egrp * kStgBytes;
        const int first_tile = blockIdx.x + buf * gridDim.x, tile_step = 2 * gridDim.x;
        const bool active = half == 0 || g.nchunks >= 2;
        const bool use_slot = g.y_ld != 0;
        int lt = 0;                                             // tiles this group has processed
        for (int tile = first_tile; active && tile < g.num_tiles; tile += tile_step, ++lt) {
            const TileCoord t = tile_coord(g, tile);
            const int n = t.n0 + ln, h = t.h0 + lh, w = t.w0 + lw;
            const bool row_ok = n < g.N;
            const int tl = buf + 2 * lt;                        // local tile index; its accumulator and barrier phase
            const int abuf = tl & (NACC - 1);
            mbar_wait(&acc_full[abuf], (uint32_t)((tl / NACC) & 1));
            tc_fence_after();
            for (int c = half; c < g.nchunks; c += 2) {
                const int cg = c * 64;                          // first channel of this chunk
                if (use_slot) {
                    if (issuer) bulk_wait_read<0>();            // this group's previous store has finished reading the slot
                    epi_bar(egrp);                              // ... and everyone knows
                }
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {                // 32 accumulator columns at a time
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(abuf * BN + cg + hh * 32), v);
                    tmem_ld_wait();
                    const int ch0 = cg + hh * 32;
                    uint32_t packed[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {               // 8 channels = one 16-byte smem chunk
                        const uint32_t chunk16 = (uint32_t)(hh * 4 + q);
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 bb = *reinterpret_cast<const float2*>(s_bias + ch0 + q * 8 + 2 * e);
                            const float2 t2 = __fadd2_rn(make_float2(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1])), bb);
                            f[2 * e] = t2.x;
                            f[2 * e + 1] = t2.y;
                        }
                        if (p.y_nchw && row_ok) {
                            const size_t hw = (size_t)g.H * g.W;
                            float* o = p.y_nchw + ((size_t)n * g.cout + ch0 + q * 8) * hw + (size_t)h * g.W + w;
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (ch0 + q * 8 + e < g.cout) o[(size_t)e * hw] = f[e];
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                            packed[q * 4 + e] = *reinterpret_cast<const uint32_t*>(&b2);
                        }
                        if (g.y_ld)
                            *reinterpret_cast<uint4*>(stg + row_off + ((chunk16 ^ sw) << 4)) =
                                make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
                    }
                    if (p.stats && ch0 < g.cout) {
                        // statistics of the ROUNDED values; a group never straddles a 32-channel slice (gs | 32)
                        float* sn = p.stats + (size_t)n * g.groups * 2;
                        if (px_per_img >= 32) {
                            if (gs == 4) row_stats<4, 32, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                            else if (gs == 8) row_stats<8, 32, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                            else if (gs == 16) row_stats<16, 32, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                            else row_stats<32, 32, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                        } else {                                // 4x4 images: a warp covers two images
                            if (gs == 4) row_stats<4, 16, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                            else if (gs == 8) row_stats<8, 16, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                            else if (gs == 16) row_stats<16, 16, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                            else row_stats<32, 16, 16>(packed, row_ok, lane, ch0, g.cout, sn);
                        }
                    }
                }
                if (g.y_ld) {
                    fence_async_smem();                         // generic-proxy writes -> visible to the TMA store
                    epi_bar(egrp);
                    if (issuer) {
                        tma_store_4d(&tmOut, stg, cg, t.w0, t.h0, t.n0);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[abuf]);                      // 128 arrivals per serving group free the accumulator buffer
        }
        if (issuer) bulk_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, NACC * BN);
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

int make_act_tmap(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    return sh_make_tmap_bf16(m, base, 4, dims, strides, box);
}

template <int BN, bool HALO, bool XF = false>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmRes, const CUtensorMap& tmOut,
                const ConvGeom& g, const ConvPtrs& p, int grid, size_t smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        SH_CUDA(cudaFuncSetAttribute(conv_fwd_kernel<BN, HALO, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr = true;
    }
    conv_fwd_kernel<BN, HALO, XF><<<grid, XF ? kThreadsXf : kThreads, smem, st>>>(tmA, tmB, tmRes, tmOut, g, p);
    SH_CHECK_LAUNCH("conv_fwd_kernel");
    return SH_OK;
}

}  // namespace

// Implicit-GEMM convolution, stride 1, 'same' zero padding, NHWC bf16 in, fp32 accumulate.
//   x        bf16 [N,H,W,Cin]           Cin % 64 == 0
//   w        bf16 [taps, cout_pad, Cin] (tap-major, K contiguous); rows >= Cout must be zero; cout_pad in {64,128,256}
//   residual bf16 [N,H,W,Cout] or null (Cout % 8 == 0)
//   y        bf16 [N,H,W,y_ld] or null; y_nchw fp32 [N,Cout,H,W] or null; stats fp32 [N,groups,2] (accumulated) or null
// The data-gradient is the same call on dY with flipped/transposed weights.
static int conv_fwd_impl(const void* x, const void* w, const void* bias, const void* residual, int N, int H, int W,
                         int Cin, int Cout, int cout_pad, int taps, void* y, int y_ld, void* y_nchw, void* stats,
                         int groups, const GnOperand& gn, void* stream) {
    SH_REQUIRE(x && w && (y || y_nchw), "sh_conv_fwd: null pointer");
    SH_REQUIRE(taps == 1 || taps == 9, "sh_conv_fwd: taps must be 1 or 9");
    SH_REQUIRE(N >= 1 && is_pow2(H) && is_pow2(W) && H >= 4 && W >= 4, "sh_conv_fwd: H, W must be powers of two >= 4");
    SH_REQUIRE(Cin % 64 == 0 && Cout >= 1 && cout_pad >= Cout && (cout_pad == 64 || cout_pad == 128 || cout_pad == 256),
               "sh_conv_fwd: Cin %% 64 == 0 and cout_pad in {64,128,256} required (Cin=%d Cout=%d cout_pad=%d)", Cin, Cout, cout_pad);
    SH_REQUIRE(!y || (y_ld % 8 == 0 && y_ld >= Cout && y_ld <= cout_pad), "sh_conv_fwd: y_ld must be a multiple of 8 in [Cout, cout_pad]");
    SH_REQUIRE(!residual || Cout % 8 == 0, "sh_conv_fwd: residual needs Cout %% 8 == 0");
    SH_REQUIRE(!stats || (groups > 0 && Cout % groups == 0 && (Cout / groups == 4 || Cout / groups == 8 || Cout / groups == 16 || Cout / groups == 32)),
               "sh_conv_fwd: unsupported GroupNorm grouping");
    cudaStream_t st = (cudaStream_t)stream;
    ConvGeom g;
    g.N = N; g.H = H; g.W = W;
    // halo mode: 3x3 without a residual on images that hold an 8 x 16 pixel tile; SH_CONV_HALO=0 keeps the nine-box path
    const char* halo_env = getenv("SH_CONV_HALO");          // read per call: the parity tests flip it inside one process
    const bool halo_on = !halo_env || atoi(halo_env) != 0;
    const bool halo = halo_on && taps == 9 && !residual && H >= 16 && W >= 8 && cout_pad <= 128;
    g.bw = halo ? 8 : (W < 16 ? W : 16);
    g.bh = H < 128 / g.bw ? H : 128 / g.bw;
    g.bn = 128 / (g.bw * g.bh);
    SH_REQUIRE(g.bn <= 8, "sh_conv_fwd: image too small");
    g.tiles_w = W / g.bw; g.tiles_h = H / g.bh;
    g.num_tiles = g.tiles_w * g.tiles_h * ((N + g.bn - 1) / g.bn);
    g.taps = taps; g.kblocks = Cin / 64; g.cout = Cout; g.cout_pad = cout_pad;
    g.y_ld = y ? y_ld : 0;
    g.groups = stats ? groups : 0;
    g.has_res = residual ? 1 : 0;
    g.fuse_gn = gn.groups > 0 ? 1 : 0;
    if (g.fuse_gn) {
        SH_REQUIRE(taps == 1 && g.bn <= 2, "sh_conv_fwd_gn: the fused input GroupNorm needs a 1x1 layer on images of >= 64 pixels");
        SH_REQUIRE(gn.stats && gn.gamma && gn.beta && Cin % gn.groups == 0 && (gn.cpg == 4 || gn.cpg == 8 || gn.cpg == 16) &&
                   (((uintptr_t)gn.gamma | (uintptr_t)gn.beta) & 15) == 0 && ((uintptr_t)gn.stats & 7) == 0,
                   "sh_conv_fwd_gn: bad GroupNorm operand (4, 8 or 16 channels per group; 16-byte aligned gamma / beta)");
    }
    const int top = (y && y_ld > Cout) ? y_ld : Cout;
    g.nchunks = (top + 63) / 64;
    // shared-memory plan: one 16 KB staging slot per epilogue group, resident weights if they fit next to >= 2 A stages
    const int num_k = taps * g.kblocks;
    const int b_tile = cout_pad * 128;
    const long resident = (long)num_k * b_tile;
    size_t smem = 0;
    {
        const int fixed = 1024 /*alignment*/ + kGroups * kStgBytes + 2048 /*barriers + bias*/ + (residual ? kIdentBytes : 0) +
                          (halo ? kHaloSlots * kHaloSlotBytes : 0);
        const int avail = kSmemLimit - fixed;
        if (halo) {
            g.b_resident = 0;
            g.stages = avail / b_tile;
            if (g.stages > kMaxStages) g.stages = kMaxStages;
            smem = (size_t)fixed + (size_t)g.stages * b_tile;
        } else if (resident + 2 * kABytes <= avail) {
            g.b_resident = 1;
            g.stages = (int)((avail - resident) / kABytes);
        } else {
            g.b_resident = 0;
            g.stages = avail / (kABytes + b_tile);
        }
        if (!halo) {
            if (g.stages > kMaxStages) g.stages = kMaxStages;
            smem = (size_t)fixed + (size_t)g.stages * (kABytes + (g.b_resident ? 0 : b_tile)) + (g.b_resident ? resident : 0);
        }
    }
    SH_REQUIRE(g.stages >= 2, "sh_conv_fwd: shared-memory plan failed");
    ConvPtrs p{(const float*)bias, (float*)y_nchw, (float*)stats, gn};
    CUtensorMap tmA, tmB, tmRes, tmOut;
    int rc = halo ? make_act_tmap(&tmA, x, N, H, W, Cin, kHaloW, kHaloH, 1) : make_act_tmap(&tmA, x, N, H, W, Cin, g.bw, g.bh, g.bn);
    if (rc) return rc;
    const uint64_t wd[2] = {(uint64_t)Cin, (uint64_t)taps * cout_pad};
    const uint64_t ws[1] = {(uint64_t)Cin * 2};
    const uint32_t wb[2] = {64, (uint32_t)cout_pad};
    rc = sh_make_tmap_bf16(&tmB, w, 2, wd, ws, wb);
    if (rc) return rc;
    rc = make_act_tmap(&tmRes, x, N, H, W, Cin, g.bw, g.bh, g.bn);      // placeholders unless replaced below
    if (rc) return rc;
    tmOut = tmRes;
    if (residual) { rc = make_act_tmap(&tmRes, residual, N, H, W, Cout, g.bw, g.bh, g.bn); if (rc) return rc; }
    if (y) { rc = make_act_tmap(&tmOut, y, N, H, W, y_ld, g.bw, g.bh, g.bn); if (rc) return rc; }
    const int grid = g.num_tiles < SH_NUM_SMS ? g.num_tiles : SH_NUM_SMS;
    if (halo) {
        if (cout_pad == 128) return launch_conv<128, true>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
        return launch_conv<64, true>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
    }
    if (g.fuse_gn) {            // the build with the four operand-transform warps (704 threads, 80 registers)
        if (cout_pad == 256) return launch_conv<256, false, true>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
        if (cout_pad == 128) return launch_conv<128, false, true>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
        return launch_conv<64, false, true>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
    }
    if (cout_pad == 256) return launch_conv<256, false>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
    if (cout_pad == 128) return launch_conv<128, false>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
    return launch_conv<64, false>(tmA, tmB, tmRes, tmOut, g, p, grid, smem, st);
}

SH_EXPORT int sh_conv_fwd(const void* x, const void* w, const void* bias, const void* residual, int N, int H, int W,
                           int Cin, int Cout, int cout_pad, int taps, void* y, int y_ld, void* y_nchw, void* stats,
                           int groups, void* stream) {
    GnOperand none{};
    return conv_fwd_impl(x, w, bias, residual, N, H, W, Cin, Cout, cout_pad, taps, y, y_ld, y_nchw, stats, groups, none, stream);
}

// The same 1x1 convolution on a = relu(groupnorm(x)): x is the RAW input of the GroupNorm (gn_stats [N,gn_groups,2] = its per-
// (sample, group) sum / sum of squares, gn_gamma / gn_beta [Cin]); the normalisation + ReLU is applied to every operand tile in
// shared memory between the TMA arrival and the MMA, so neither the gn_relu_fwd pass nor the tensor a exists.  Bit-identical to
// sh_gn_relu_fwd followed by sh_conv_fwd.  Needs taps == 1 and images of >= 64 pixels.
SH_EXPORT int sh_conv_fwd_gn(const void* x, const void* gn_stats, const void* gn_gamma, const void* gn_beta, int gn_groups, float gn_eps,
                              const void* w, const void* bias, const void* residual, int N, int H, int W, int Cin, int Cout,
                              int cout_pad, void* y, int y_ld, void* y_nchw, void* stats, int groups, void* stream) {
    SH_REQUIRE(gn_groups >= 1 && Cin % gn_groups == 0, "sh_conv_fwd_gn: bad gn_groups");
    GnOperand gn{(const float*)gn_stats, (const float*)gn_gamma, (const float*)gn_beta, gn_groups, Cin / gn_groups,
                 1.f / ((float)(H * W) * (float)(Cin / gn_groups)), gn_eps};
    return conv_fwd_impl(x, w, bias, residual, N, H, W, Cin, Cout, cout_pad, 1, y, y_ld, y_nchw, stats, groups, gn, stream);
}
