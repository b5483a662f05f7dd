// Hourglass convolutions, weight gradient, on the 5th-generation tensor cores: tcgen05.mma + TMEM + TMA (sm_100a).
// (forward / data-gradient: conv_fwd.cu)
//
// Replaces the cuDNN/ATen kernels behind every nn.Conv2d of the reference network
//   Bottleneck.conv1/conv2/conv3, downsample          /root/reference/network/hourglass.py:13-18, 125-128
//   HourglassNet.fc / score / fc_ / score_            /root/reference/network/hourglass.py:110-114, 138-145
// forward, data-gradient (same kernel on flipped/transposed weights) and weight-gradient.
//
// Formulation: implicit GEMM over NHWC bf16 activations, fp32 accumulation in TMEM.
//   forward / dgrad :  D[128 pixels, BN channels] += A[128 pixels, 64] * B[BN, 64]^T   per (tap, 64-channel block)
//       A tile = one TMA box {64 ch, bw, bh, bn} of the 4-D activation tensor at spatial offset (kw-1, kh-1): TMA's
//       out-of-bounds zero fill IS the convolution padding (im2col staging with no im2col buffer); 128-byte swizzle,
//       K-major UMMA descriptors; B tile = 2-D TMA box of the [tap][Cout][Cin] weight matrix.
//   wgrad          :  dW[tap][128 co, 128 ci] += dY[pixels, co]^T * X_shifted[pixels, ci]   (K = pixels)
//       both operands are "MN-major" in shared memory (channels contiguous), K split across CTAs, fp32 atomics.
// Warp roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / residual / GroupNorm statistics -> bf16 NHWC and/or fp32 NCHW).
// smem <-> MMA pipeline: kStages-deep mbarrier ring (full/empty); MMA -> epilogue: one tcgen05.commit barrier.
#include "tc_common.cuh"

// ------------------------------------------------------------------------------------------------ host: tensor maps
sh_encode_tiled_fn sh_get_encode_tiled() {
    static sh_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (sh_encode_tiled_fn)p;
    }
    return fn;
}

int sh_make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
    sh_encode_tiled_fn enc = sh_get_encode_tiled();
    if (!enc) {
        snprintf(g_sh_last_error, sizeof(g_sh_last_error), "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return SH_ERR_CUDA;
    }
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_sh_last_error, sizeof(g_sh_last_error), "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u",
                 (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return SH_ERR_CUDA;
    }
    return SH_OK;
}

namespace {

constexpr int kThreads = 192;
constexpr int kSmemMax = 232448;    // 227 KB opt-in maximum per CTA

// ------------------------------------------------------------------------------------------------ wgrad
// dW[tap][co][ci] += sum over pixels of dY[pix, co] * X[pix + tap, ci];  M = co (128), N = ci (BNW), K = pixels.
constexpr int kWStages = 3;
constexpr int kWPix = 64;          // pixels (K) per stage
template <int BNW>
struct WgradSmem {
    static constexpr int kABytes = 2 * kWPix * 128;            // two 64-channel chunks of dY   (16 KB)
    static constexpr int kBBytes = (BNW / 64) * kWPix * 128;   // BNW/64 chunks of X
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kWStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + 256 + 1024;
};

struct WgradGeom {
    int N, H, W;
    int bw, bh, bn;                // pixel box of kWPix = bw*bh*bn pixels
    int tiles_w, tiles_h, tiles_n; // pixel tiles
    int taps;
    int cin, cout;                 // real channel counts = extents of the [Cout,Cin,k,k] gradient tensor
    int cin_pad, cout_pad;         // GEMM extents (multiples of BNW / 128); channels beyond the tensors read as zero
    int ksplit;                    // CTAs along the pixel dimension
    int tap_major;                 // 1: dw is the [tap][Cout][Cin] scratch of sh_conv_wgrad3x3 (contiguous in Cin: vector reductions)
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d);

template <int BNW>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                            const __grid_constant__ CUtensorMap tmX,
                                                            const WgradGeom g, float* __restrict__ dw) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    using S = WgradSmem<BNW>;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty_bar = full_bar + kWStages;
    uint64_t* accum_bar = empty_bar + kWStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = ((tap * co_tiles + co_t) * ci_tiles + ci_t), blockIdx.y = k-split index
    const int ci_tiles = g.cin_pad / BNW, co_tiles = g.cout_pad / 128;
    int t = blockIdx.x;
    const int ci_t = t % ci_tiles; t /= ci_tiles;
    const int co_t = t % co_tiles; t /= co_tiles;
    const int tap = t;
    const int dh = g.taps == 9 ? tap / 3 - 1 : 0, dw_ = g.taps == 9 ? tap % 3 - 1 : 0;
    const int total_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
    const int per = (total_tiles + g.ksplit - 1) / g.ksplit;
    const int tile_begin = blockIdx.y * per;
    const int tile_end = min(tile_begin + per, total_tiles);
    const int num_k = max(tile_end - tile_begin, 0);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < kWStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BNW);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int k = 0; k < num_k; ++k) {
                const int s = k % kWStages, it = k / kWStages;
                mbar_wait(&empty_bar[s], (it & 1) ^ 1);
                int tt = tile_begin + k;
                const int tw = tt % g.tiles_w; tt /= g.tiles_w;
                const int th = tt % g.tiles_h; tt /= g.tiles_h;
                const int n0 = tt * g.bn, h0 = th * g.bh, w0 = tw * g.bw;
                uint8_t* a_dst = smem + s * S::kStageBytes;
                uint8_t* b_dst = a_dst + S::kABytes;
                mbar_expect_tx(&full_bar[s], S::kStageBytes);
                for (int c = 0; c < 2; ++c)
                    tma_load_4d(a_dst + c * kWPix * 128, &tmDY, &full_bar[s], co_t * 128 + c * 64, w0, h0, n0);
                for (int c = 0; c < BNW / 64; ++c)
                    tma_load_4d(b_dst + c * kWPix * 128, &tmX, &full_bar[s], ci_t * BNW + c * 64, w0 + dw_, h0 + dh, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, BNW, 1, 1);
            for (int k = 0; k < num_k; ++k) {
                const int s = k % kWStages, it = k / kWStages;
                mbar_wait(&full_bar[s], it & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
                const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
                for (int kk = 0; kk < kWPix / 16; ++kk) {
                    // 16 pixels (K) = 16 rows of 128 B = 2048 B further into each 64-channel chunk
                    const uint64_t ad = umma_desc_mnmajor_sw128(a_addr + kk * 2048, kWPix * 128);
                    const uint64_t bd = umma_desc_mnmajor_sw128(b_addr + kk * 2048, kWPix * 128);
                    umma_bf16(tmem_base, ad, bd, idesc, (k | kk) != 0);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        const int quad = warp & 3;
        const int co = co_t * 128 + quad * 32 + lane;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (num_k > 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < BNW; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
                if (co < g.cout) {
                    const int ci0 = ci_t * BNW + c0;
                    if (g.tap_major) {
                        // the tap-major scratch of the 3x3 layers: 32 contiguous floats per thread, eight 16-byte reductions
                        float* dst = dw + ((size_t)tap * g.cout + co) * g.cin + ci0;
                        if ((g.cin & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (ci0 + j < g.cin)
                                    red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                               __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (ci0 + j < g.cin) atomicAdd(dst + j, __uint_as_float(v[j]));
                        }
                    } else {
                        // straight into the reference layout [Cout][Cin][kh][kw]
                        float* dst = dw + ((size_t)co * g.cin + ci0) * g.taps + tap;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (ci0 + j < g.cin) atomicAdd(dst + (size_t)j * g.taps, __uint_as_float(v[j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BNW);
}

// ------------------------------------------------------------------------------------------------ wgrad, 1x1 layers
// These layers are HBM-bound (2*HW*(Cin+Cout) bytes per image for 2*HW*Cin*Cout flops): every byte must be read ONCE.
// One CTA owns the WHOLE [Cout <= 256] x [Cin <= 256] gradient: co_tiles (1 or 2) accumulators of 128 x cin_pad fp32 in
// TMEM (<= 512 columns), so a pixel tile of dY and of X is fetched once chip-wide; K (= pixels) is split over one CTA
// per SM with a deep TMA ring; the epilogue adds the partials with vector reductions (red.global.add.v4.f32).
constexpr int kW1MaxStages = 8;
struct Wgrad1Geom {
    int N, H, W;
    int bw, bh, bn;
    int tiles_w, tiles_h, total_tiles;
    int cin, cout;                 // real channel counts
    int cin_pad;                   // N of the MMA (64, 128 or 256)
    int co_tiles;                  // 1 or 2 accumulators of 128 rows
    int stages, stage_bytes;
    int fuse_gn;                   // the X tiles are x; the MMA multiplies relu(groupnorm(x)) (one image per pixel tile)
    GnOperand gn;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) wgrad1x1_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const Wgrad1Geom g, float* __restrict__ dw) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
    uint64_t* empty_bar = full_bar + kW1MaxStages;
    uint64_t* accum_bar = empty_bar + kW1MaxStages;
    uint64_t* xf_bar = accum_bar + 1;                        // [kW1MaxStages]: the X chunks of the stage have been transformed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf_bar + kW1MaxStages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (g.total_tiles + gridDim.x - 1) / gridDim.x;
    const int tile_begin = blockIdx.x * per;
    const int tile_end = min(tile_begin + per, g.total_tiles);
    const int num_k = max(tile_end - tile_begin, 0);
    const int a_bytes = g.co_tiles * 2 * kWPix * 128;
    const uint32_t tmem_cols = (uint32_t)(g.co_tiles * g.cin_pad);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < kW1MaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&xf_bar[s], 128); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int k = 0; k < num_k; ++k) {
                const int s = k % g.stages, it = k / g.stages;
                mbar_wait(&empty_bar[s], (it & 1) ^ 1);
                int tt = tile_begin + k;
                const int tw = tt % g.tiles_w; tt /= g.tiles_w;
                const int th = tt % g.tiles_h; tt /= g.tiles_h;
                const int n0 = tt * g.bn, h0 = th * g.bh, w0 = tw * g.bw;
                uint8_t* a_dst = smem + s * g.stage_bytes;
                uint8_t* b_dst = a_dst + a_bytes;
                mbar_expect_tx(&full_bar[s], (uint32_t)g.stage_bytes);
                for (int c = 0; c < g.co_tiles * 2; ++c)
                    tma_load_4d(a_dst + c * kWPix * 128, &tmDY, &full_bar[s], c * 64, w0, h0, n0);
                for (int c = 0; c < g.cin_pad / 64; ++c)
                    tma_load_4d(b_dst + c * kWPix * 128, &tmX, &full_bar[s], c * 64, w0, h0, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, g.cin_pad, 1, 1);
            for (int k = 0; k < num_k; ++k) {
                const int s = k % g.stages, it = k / g.stages;
                mbar_wait(g.fuse_gn ? &xf_bar[s] : &full_bar[s], it & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * g.stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
                for (int ct = 0; ct < g.co_tiles; ++ct) {
#pragma unroll
                    for (int kk = 0; kk < kWPix / 16; ++kk) {
                        const uint64_t ad = umma_desc_mnmajor_sw128(a_addr + ct * 2 * kWPix * 128 + kk * 2048, kWPix * 128);
                        const uint64_t bd = umma_desc_mnmajor_sw128(b_addr + kk * 2048, kWPix * 128);
                        umma_bf16(tmem_base + (uint32_t)(ct * g.cin_pad), ad, bd, idesc, (k | kk) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        if (g.fuse_gn) {
            // the four epilogue warps are idle until the accumulator is complete: they apply GroupNorm + ReLU to the X chunks of
            // every stage in shared memory (X in HBM is the raw GroupNorm input; its normalised form is never materialised)
            // All four warps work on the same stage: the latency "tile landed -> tile transformed" is added to every trip of the ring.
            const int tt = (int)threadIdx.x - 64;               // 0..127
            const int j = tt & 7, r = tt >> 3;                  // 16-byte chunk of the 128-byte rows; rows r, r+16, r+32, r+48
            const int tiles_per_img = g.tiles_w * g.tiles_h;    // bn == 1
            static_assert(kWPix == 64, "the transform covers a 64-pixel slab per 64-channel chunk");
            for (int k = 0; k < num_k; ++k) {
                const int s = k % g.stages, it = k / g.stages;
                const int n = min((tile_begin + k) / tiles_per_img, g.N - 1);
                const uint32_t b_dst = smem_u32(smem + s * g.stage_bytes + a_bytes);
                for (int c = 0; c < g.cin_pad / 64; ++c) {
                    float ka[8], kb[8];
                    gn_scale_shift(g.gn, n, c * 64 + j * 8, ka, kb);          // (global loads: issued before the wait)
                    if (c == 0) mbar_wait(&full_bar[s], it & 1);
                    gn_xform_rows4(b_dst + c * kWPix * 128, r, j, ka, kb);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&xf_bar[s]);
            }
        }
        const int quad = warp & 3;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (num_k > 0) {
            const int vec = (g.cin % 4 == 0) ? 4 : ((g.cin % 2 == 0) ? 2 : 1);
            for (int ct = 0; ct < g.co_tiles; ++ct) {
                const int co = ct * 128 + quad * 32 + lane;
#pragma unroll 1
                for (int c0 = 0; c0 < g.cin_pad; c0 += 32) {
                    if (c0 >= g.cin) break;                       // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ct * g.cin_pad + c0), v);
                    tmem_ld_wait();
                    if (co < g.cout) {
                        float* dst = dw + (size_t)co * g.cin + c0;
                        if (vec == 4) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (c0 + j < g.cin)
                                    red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                               __uint_as_float(v[j + 3]));
                        } else if (vec == 2) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2)
                                if (c0 + j < g.cin) red_add_v2(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < g.cin) atomicAdd(dst + j, __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ wgrad, 3x3 layers
// dW[kh][kw][co][ci] += sum_p dY[p][co] * X[p + (kh-1, kw-1)][ci].  With one TMA box per tap the smem fill traffic (32 KB
// per 256 MMA cycles) is 3x what the L2 can feed.  Here a CTA owns ONE kernel row kh and keeps THREE accumulators
// (kw = 0, 1, 2) in TMEM: per 64-pixel step it loads the dY tile once and ONE X box that is two pixels wider
// ({64 ch, 18, 4}); the three taps are three views of that box -- the MN-major descriptor start shifted by kw rows
// (tcgen05 swizzles on absolute smem address bits, tools/umma_probe.cu) -- so 34 KB feed 768 MMA cycles.
// Output: fp32 partial sums added with red.global.add.v4 into a [tap][Cout][Cin] scratch (contiguous in ci), unpacked
// into the reference layout [Cout][Cin][kh][kw] for all layers at once by sh_unpack_wgrad_batch.
constexpr int kW3MaxStages = 6;
constexpr int kW3XChunk = 18 * 4 * 128;        // 9216 B: one 64-channel chunk of the widened X box
struct Wgrad3Geom {
    int N, H, W;
    int tiles_w, tiles_h, total_tiles;   // 64-pixel tiles (16 wide x 4 tall)
    int cin, cout;
    int ci_tiles, co_tiles;              // nmma-channel / 128-channel tiles
    int nmma;                            // N of the MMA = input channels per CTA (64 or 128)
    int ksplit, stages, stage_bytes;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const Wgrad3Geom g, float* __restrict__ scratch) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
    uint64_t* empty_bar = full_bar + kW3MaxStages;
    uint64_t* accum_bar = empty_bar + kW3MaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;                                      // ((kh * co_tiles + co_t) * ci_tiles + ci_t)
    const int ci_t = t % g.ci_tiles; t /= g.ci_tiles;
    const int co_t = t % g.co_tiles; t /= g.co_tiles;
    const int kh = t;
    const int per = (g.total_tiles + g.ksplit - 1) / g.ksplit;
    const int tile_begin = blockIdx.y * per;
    const int tile_end = min(tile_begin + per, g.total_tiles);
    const int num_k = max(tile_end - tile_begin, 0);
    constexpr int kABytes3 = 2 * kWPix * 128;               // dY: two 64-channel chunks of 64 pixels

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < kW3MaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int k = 0; k < num_k; ++k) {
                const int s = k % g.stages, it = k / g.stages;
                mbar_wait(&empty_bar[s], (it & 1) ^ 1);
                int tt = tile_begin + k;
                const int tw = tt % g.tiles_w; tt /= g.tiles_w;
                const int th = tt % g.tiles_h; tt /= g.tiles_h;
                const int n0 = tt, h0 = th * 4, w0 = tw * 16;
                uint8_t* a_dst = smem + s * g.stage_bytes;
                uint8_t* b_dst = a_dst + kABytes3;
                mbar_expect_tx(&full_bar[s], (uint32_t)g.stage_bytes);
                for (int c = 0; c < 2; ++c)
                    tma_load_4d(a_dst + c * kWPix * 128, &tmDY, &full_bar[s], co_t * 128 + c * 64, w0, h0, n0);
                for (int c = 0; c < g.nmma / 64; ++c)        // X rows h0 + kh - 1 .., columns w0 - 1 .. w0 + 16 (zero fill = padding)
                    tma_load_4d(b_dst + c * kW3XChunk, &tmX, &full_bar[s], ci_t * g.nmma + c * 64, w0 - 1, h0 + kh - 1, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, g.nmma, 1, 1);
            for (int k = 0; k < num_k; ++k) {
                const int s = k % g.stages, it = k / g.stages;
                mbar_wait(&full_bar[s], it & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * g.stage_bytes);
                const uint32_t b_addr = a_addr + kABytes3;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {         // K = 16 pixels = one image row of the tile
                        const uint64_t ad = umma_desc_mnmajor_sw128(a_addr + hh * 2048, kWPix * 128);
                        const uint64_t bd = umma_desc_mnmajor_sw128(b_addr + (uint32_t)(hh * 18 + kw) * 128u, kW3XChunk);
                        umma_bf16(tmem_base + (uint32_t)(kw * g.nmma), ad, bd, idesc, (k | hh) != 0);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        const int quad = warp & 3;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (num_k > 0) {
            const int co = co_t * 128 + quad * 32 + lane;
            const bool vec4 = (g.cin % 4) == 0;
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                float* base = scratch + ((size_t)(kh * 3 + kw) * g.cout + co) * g.cin;
#pragma unroll 1
                for (int c0 = 0; c0 < g.nmma; c0 += 32) {
                    const int ci0 = ci_t * g.nmma + c0;
                    if (ci0 >= g.cin) break;                 // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kw * g.nmma + c0), v);
                    tmem_ld_wait();
                    if (co < g.cout) {
                        if (vec4) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (ci0 + j < g.cin)
                                    red_add_v4(base + ci0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                               __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (ci0 + j < g.cin) atomicAdd(base + ci0 + j, __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// grad[co][ci][tap] += scratch[tap][co][ci] for every table row (scratch offset, grad offset, Cout, Cin), one launch
__global__ void unpack_wgrad_batch_kernel(const int* __restrict__ table, const float* __restrict__ scratch, float* __restrict__ grad) {
    const int* t = table + blockIdx.y * 4;
    const float* src = scratch + t[0];
    float* dst = grad + t[1];
    const int cout = t[2], cin = t[3];
    const long n = (long)cout * cin * 9;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int tap = (int)(i % 9);
        const long cc = i / 9;                               // co * cin + ci
        dst[i] += src[(size_t)tap * cout * cin + cc];
    }
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

int make_act_tmap(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    return sh_make_tmap_bf16(m, base, 4, dims, strides, box);
}

}  // namespace

static int launch_wgrad_generic(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, int taps,
                                void* dw, int tap_major, cudaStream_t st) {
    WgradGeom g;
    g.N = N; g.H = H; g.W = W;
    g.bw = W < 16 ? W : 16;
    g.bh = H < kWPix / g.bw ? H : kWPix / g.bw;
    g.bn = kWPix / (g.bw * g.bh);
    g.tiles_w = W / g.bw; g.tiles_h = H / g.bh; g.tiles_n = (N + g.bn - 1) / g.bn;
    g.taps = taps; g.cin = Cin; g.cout = Cout;
    const int bnw = (x_C % 128 == 0) ? 128 : 64;
    g.cin_pad = x_C;
    g.cout_pad = (dy_C + 127) / 128 * 128;
    const int out_tiles = taps * (g.cout_pad / 128) * (g.cin_pad / bnw);
    const int total_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
    // tap-major (small 3x3 layers): one wave of CTAs -- every extra k-split is another full set of output reductions, and at these
    // sizes the reductions, not the MMAs, are the cost
    int ksplit = ((tap_major ? 1 : 2) * SH_NUM_SMS + out_tiles - 1) / out_tiles;
    if (ksplit > total_tiles) ksplit = total_tiles;
    if (ksplit < 1) ksplit = 1;
    g.ksplit = ksplit;
    g.tap_major = tap_major;
    CUtensorMap tmDY, tmX;
    int rc = make_act_tmap(&tmDY, dy, N, H, W, dy_C, g.bw, g.bh, g.bn);
    if (rc) return rc;
    rc = make_act_tmap(&tmX, x, N, H, W, x_C, g.bw, g.bh, g.bn);
    if (rc) return rc;
    dim3 grid(out_tiles, ksplit);
    if (bnw == 128) {
        static bool attr = false;
        if (!attr) { SH_CUDA(cudaFuncSetAttribute(wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<128>::kTotal)); attr = true; }
        wgrad_kernel<128><<<grid, kThreads, WgradSmem<128>::kTotal, st>>>(tmDY, tmX, g, (float*)dw);
    } else {
        static bool attr = false;
        if (!attr) { SH_CUDA(cudaFuncSetAttribute(wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<64>::kTotal)); attr = true; }
        wgrad_kernel<64><<<grid, kThreads, WgradSmem<64>::kTotal, st>>>(tmDY, tmX, g, (float*)dw);
    }
    SH_CHECK_LAUNCH("wgrad_kernel");
    return SH_OK;
}


// Weight gradient: dw fp32 [Cout, Cin, k, k] (the reference layout) += dY^T X_shifted, accumulated atomically (zero it
// first).  dy bf16 [N,H,W,dy_C], x bf16 [N,H,W,x_C]; Cout <= dy_C, Cin <= x_C are the real channel counts; channels
// beyond dy_C / x_C needed to fill the 128-wide MMA tile are supplied as zeros by TMA's out-of-bounds fill.
static int conv_wgrad_impl(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, int taps,
                           void* dw, const GnOperand& gn, void* stream) {
    SH_REQUIRE(dy && x && dw, "sh_conv_wgrad: null pointer");
    SH_REQUIRE(taps == 1 || taps == 9, "sh_conv_wgrad: taps must be 1 or 9");
    SH_REQUIRE(N >= 1 && is_pow2(H) && is_pow2(W) && H >= 4 && W >= 4, "sh_conv_wgrad: H, W must be powers of two >= 4");
    SH_REQUIRE(x_C % 64 == 0 && dy_C % 64 == 0 && Cin >= 1 && Cin <= x_C && Cout >= 1 && Cout <= dy_C,
               "sh_conv_wgrad: channel counts must be multiples of 64 in memory");
    cudaStream_t st = (cudaStream_t)stream;
    if (taps == 1 && x_C <= 256 && dy_C <= 256) {
        Wgrad1Geom g1;
        g1.N = N; g1.H = H; g1.W = W;
        g1.bw = W < 16 ? W : 16;
        g1.bh = H < kWPix / g1.bw ? H : kWPix / g1.bw;
        g1.bn = kWPix / (g1.bw * g1.bh);
        g1.tiles_w = W / g1.bw; g1.tiles_h = H / g1.bh;
        g1.total_tiles = g1.tiles_w * g1.tiles_h * ((N + g1.bn - 1) / g1.bn);
        g1.cin = Cin; g1.cout = Cout;
        g1.cin_pad = x_C;
        g1.co_tiles = (dy_C + 127) / 128;
        g1.fuse_gn = gn.groups > 0 ? 1 : 0;
        g1.gn = gn;
        if (g1.fuse_gn) {
            SH_REQUIRE(g1.bn == 1, "sh_conv_wgrad_gn: the fused input GroupNorm needs images of >= 64 pixels");
            SH_REQUIRE(gn.stats && gn.gamma && gn.beta && x_C % gn.groups == 0 && (gn.cpg == 4 || gn.cpg == 8 || gn.cpg == 16) &&
                       (((uintptr_t)gn.gamma | (uintptr_t)gn.beta) & 15) == 0 && ((uintptr_t)gn.stats & 7) == 0,
                       "sh_conv_wgrad_gn: bad GroupNorm operand (4, 8 or 16 channels per group; 16-byte aligned gamma / beta)");
        }
        g1.stage_bytes = (g1.co_tiles * 2 + g1.cin_pad / 64) * kWPix * 128;
        g1.stages = (kSmemMax - 2048) / g1.stage_bytes;
        if (g1.stages > kW1MaxStages) g1.stages = kW1MaxStages;
        CUtensorMap tmDY, tmX;
        int rc = make_act_tmap(&tmDY, dy, N, H, W, dy_C, g1.bw, g1.bh, g1.bn);
        if (rc) return rc;
        rc = make_act_tmap(&tmX, x, N, H, W, x_C, g1.bw, g1.bh, g1.bn);
        if (rc) return rc;
        static bool attr1 = false;
        if (!attr1) { SH_CUDA(cudaFuncSetAttribute(wgrad1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)); attr1 = true; }
        // every CTA ends with a full [Cout x Cin] set of reductions: on the small levels a CTA gets at least 4 pixel tiles (measured
        // on the 8x8 level: 22 -> 17 us; larger levels already have more tiles than that per SM)
        const int mink = 4;
        int grid1 = (g1.total_tiles + mink - 1) / mink;
        if (grid1 > SH_NUM_SMS) grid1 = SH_NUM_SMS;
        if (grid1 < 1) grid1 = 1;
        const size_t smem1 = (size_t)g1.stages * g1.stage_bytes + 1024 + 512;
        wgrad1x1_kernel<<<grid1, kThreads, smem1, st>>>(tmDY, tmX, g1, (float*)dw);
        SH_CHECK_LAUNCH("wgrad1x1_kernel");
        return SH_OK;
    }
    SH_REQUIRE(gn.groups == 0, "sh_conv_wgrad_gn: only 1x1 layers with <= 256 channels take a fused input GroupNorm");
    return launch_wgrad_generic(dy, x, N, H, W, x_C, Cin, dy_C, Cout, taps, dw, 0, st);
}

SH_EXPORT int sh_conv_wgrad(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, int taps,
                            void* dw, void* stream) {
    GnOperand none{};
    return conv_wgrad_impl(dy, x, N, H, W, x_C, Cin, dy_C, Cout, taps, dw, none, stream);
}

// The 1x1 weight gradient against a = relu(groupnorm(x)) where only the raw x exists in HBM (see sh_conv_fwd_gn): the X tiles are
// normalised in shared memory by the warps that otherwise wait for the accumulator.  x_C == Cin (the GroupNorm's channel count).
SH_EXPORT int sh_conv_wgrad_gn(const void* dy, const void* x, const void* gn_stats, const void* gn_gamma, const void* gn_beta, int gn_groups,
                               float gn_eps, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, void* dw, void* stream) {
    SH_REQUIRE(gn_groups >= 1 && x_C % gn_groups == 0 && x_C == Cin, "sh_conv_wgrad_gn: bad gn_groups / channel count");
    GnOperand gn{(const float*)gn_stats, (const float*)gn_gamma, (const float*)gn_beta, gn_groups, x_C / gn_groups,
                 1.f / ((float)(H * W) * (float)(x_C / gn_groups)), gn_eps};
    return conv_wgrad_impl(dy, x, N, H, W, x_C, Cin, dy_C, Cout, 1, dw, gn, stream);
}

// 3x3 weight gradient into a [9][Cout][Cin] fp32 scratch (accumulated: zero it once per step).  W >= 16: the kernel-row kernel;
// narrower images (the 8x8 / 4x4 levels): the per-tap kernel writing the same scratch with vector reductions.
SH_EXPORT int sh_conv_wgrad3x3(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout,
                               void* scratch, void* stream) {
    SH_REQUIRE(dy && x && scratch, "sh_conv_wgrad3x3: null pointer");
    SH_REQUIRE(N >= 1 && is_pow2(H) && is_pow2(W) && H >= 4 && W >= 4, "sh_conv_wgrad3x3: H, W must be powers of two >= 4");
    SH_REQUIRE(x_C % 64 == 0 && dy_C % 64 == 0 && Cin >= 1 && Cin <= x_C && Cout >= 1 && Cout <= dy_C,
               "sh_conv_wgrad3x3: channel counts must be multiples of 64 in memory");
    cudaStream_t st = (cudaStream_t)stream;
    if (W < 16) return launch_wgrad_generic(dy, x, N, H, W, x_C, Cin, dy_C, Cout, 9, scratch, 1, st);
    Wgrad3Geom g;
    g.N = N; g.H = H; g.W = W;
    g.tiles_w = W / 16; g.tiles_h = H / 4;
    g.total_tiles = g.tiles_w * g.tiles_h * N;
    g.cin = Cin; g.cout = Cout;
    g.nmma = x_C % 128 == 0 ? 128 : 64;
    g.ci_tiles = x_C / g.nmma; g.co_tiles = (dy_C + 127) / 128;
    const int out_tiles = 3 * g.ci_tiles * g.co_tiles;
    int ksplit = SH_NUM_SMS / out_tiles;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > g.total_tiles) ksplit = g.total_tiles;
    g.ksplit = ksplit;
    g.stage_bytes = 2 * kWPix * 128 + (g.nmma / 64) * kW3XChunk;      // 16 KB + 9 KB per 64 input channels: 1024-aligned
    g.stages = kW3MaxStages;
    CUtensorMap tmDY, tmX;
    int rc = make_act_tmap(&tmDY, dy, N, H, W, dy_C, 16, 4, 1);
    if (rc) return rc;
    rc = make_act_tmap(&tmX, x, N, H, W, x_C, 18, 4, 1);
    if (rc) return rc;
    const size_t smem = (size_t)g.stages * g.stage_bytes + 1024 + 512;
    static bool attr = false;
    if (!attr) { SH_CUDA(cudaFuncSetAttribute(wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax)); attr = true; }
    wgrad3x3_kernel<<<dim3(out_tiles, ksplit), kThreads, smem, st>>>(tmDY, tmX, g, (float*)scratch);
    SH_CHECK_LAUNCH("wgrad3x3_kernel");
    return SH_OK;
}

// table int32 [n,4] (device): (scratch offset, grad offset, Cout, Cin) in floats; grad[co][ci][kh][kw] += scratch[tap][co][ci]
SH_EXPORT int sh_unpack_wgrad_batch(const void* table, int n, const void* scratch, void* grad, void* stream) {
    SH_REQUIRE(table && scratch && grad && n >= 0, "sh_unpack_wgrad_batch: bad arguments");
    if (n == 0) return SH_OK;
    unpack_wgrad_batch_kernel<<<dim3(32, n), 256, 0, (cudaStream_t)stream>>>((const int*)table, (const float*)scratch, (float*)grad);
    SH_CHECK_LAUNCH("unpack_wgrad_batch_kernel");
    return SH_OK;
}
