// Hourglass convolutions on the 5th-generation tensor cores: tcgen05.mma + TMEM accumulators + TMA staging (sm_100a).
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

constexpr int kStages = 3;
constexpr int kBM = 128;           // pixels per tile = TMEM lanes
constexpr int kBK = 64;            // channels per k-block = one 128-byte swizzle row
constexpr int kThreads = 192;

struct ConvGeom {
    int N, H, W;                   // activation tensor
    int bw, bh, bn;                // TMA box (pixels): bw*bh*bn == 128
    int tiles_w, tiles_h;
    int taps;                      // 1 or 9
    int kblocks;                   // Cin / 64
    int cout;                      // real output channels
    int cout_pad;                  // rows per tap in the weight matrix (multiple of BN)
    int y_ld;                      // channel stride of the bf16 output (>= cout)
    int groups;                    // GroupNorm groups for the statistics side output (0 = none)
};

struct ConvPtrs {
    const float* bias;             // [cout] or null
    const __nv_bfloat16* residual; // [N,H,W,cout] or null
    __nv_bfloat16* y;              // [N,H,W,y_ld] or null
    float* y_nchw;                 // [N,cout,H,W] fp32 or null
    float* stats;                  // [N,groups,2] fp32 (sum, sum of squares), accumulated atomically
};

template <int BN>
struct ConvSmem {
    static constexpr int kABytes = kBM * kBK * 2;         // 16 KB
    static constexpr int kBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + 256 + 1024; // barriers + slack for 1024-B alignment
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const ConvGeom g, const ConvPtrs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    using S = ConvSmem<BN>;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* accum_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    __shared__ float s_stats[8][32][2];     // [local image][group][sum, sumsq]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile coordinates
    int t = blockIdx.x;
    const int tw = t % g.tiles_w; t /= g.tiles_w;
    const int th = t % g.tiles_h; t /= g.tiles_h;
    const int n0 = t * g.bn, h0 = th * g.bh, w0 = tw * g.bw;
    const int co0 = blockIdx.y * BN;
    const int num_k = g.taps * g.kblocks;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < 8 * 32 * 2; i += kThreads) (&s_stats[0][0][0])[i] = 0.f;
    if (warp == 1) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            for (int k = 0; k < num_k; ++k) {
                const int s = k % kStages, it = k / kStages;
                mbar_wait(&empty_bar[s], (it & 1) ^ 1);
                const int tap = k / g.kblocks, kb = k - tap * g.kblocks;
                const int dh = g.taps == 9 ? tap / 3 - 1 : 0, dw = g.taps == 9 ? tap % 3 - 1 : 0;
                uint8_t* a_dst = smem + s * S::kStageBytes;
                uint8_t* b_dst = a_dst + S::kABytes;
                mbar_expect_tx(&full_bar[s], S::kStageBytes);
                tma_load_4d(a_dst, &tmA, &full_bar[s], kb * kBK, w0 + dw, h0 + dh, n0);
                tma_load_2d(b_dst, &tmB, &full_bar[s], kb * kBK, tap * g.cout_pad + co0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
            for (int k = 0; k < num_k; ++k) {
                const int s = k % kStages, it = k / kStages;
                mbar_wait(&full_bar[s], it & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
                const uint32_t b_addr = a_addr + S::kABytes;
                const uint64_t ad = umma_desc_kmajor_sw128(a_addr);
                const uint64_t bd = umma_desc_kmajor_sw128(b_addr);
#pragma unroll
                for (int kk = 0; kk < kBK / 16; ++kk) {
                    // advance 16 bf16 = 32 B inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                    umma_bf16(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0);
                }
                umma_commit(&empty_bar[s]);       // frees the smem slot when these MMAs retire
            }
            umma_commit(accum_bar);               // accumulator complete
        }
    } else {
        // ===================== epilogue warps (TMEM -> registers -> global) =====================
        const int quad = warp & 3;                // TMEM lane quadrant this warp may read
        const int m = quad * 32 + lane;           // tile row = pixel
        const int ln = m / (g.bw * g.bh), lh = (m / g.bw) % g.bh, lw = m % g.bw;
        const int n = n0 + ln, h = h0 + lh, w = w0 + lw;
        const bool row_ok = n < g.N;
        const size_t pix = ((size_t)n * g.H + h) * g.W + w;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const bool warp_one_image = ((g.bw * g.bh) % 32) == 0;
        const int gs = g.groups > 0 ? g.cout / g.groups : 1;      // channels per group
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            const int cg = co0 + c0;              // first global channel of this chunk
            if (cg >= g.cout) break;              // padded output channels (warp-uniform)
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] += (cg + j < g.cout) ? __ldg(p.bias + cg + j) : 0.f;
            }
            if (p.residual && row_ok) {
                const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + pix * g.cout + cg);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (cg + q * 8 < g.cout) {
                        const uint4 rv = __ldg(r4 + q);
                        const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            f[q * 8 + 2 * e] += __uint_as_float(rr[e] << 16);
                            f[q * 8 + 2 * e + 1] += __uint_as_float(rr[e] & 0xffff0000u);
                        }
                    }
                }
            }
            // round to the stored precision first so that the statistics describe the stored tensor
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                packed[j] = *reinterpret_cast<const uint32_t*>(&b2);
            }
            if (p.y && row_ok) {
                uint4* o4 = reinterpret_cast<uint4*>(p.y + pix * g.y_ld + cg);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (cg + q * 8 < g.y_ld) o4[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            if (p.y_nchw && row_ok) {
                const size_t hw = (size_t)g.H * g.W;
                float* o = p.y_nchw + ((size_t)n * g.cout + cg) * hw + (size_t)h * g.W + w;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (cg + j < g.cout) o[(size_t)j * hw] = f[j];
            }
            if (p.stats) {
                // per-thread partial sums over the channels of each group in this chunk, on the rounded values
                const int ngroups = 32 / gs;                       // gs in {4, 8, 16, 32}
                for (int gi = 0; gi < ngroups; ++gi) {
                    float s1 = 0.f, s2 = 0.f;
                    if (row_ok) {
                        for (int j = gi * gs; j < (gi + 1) * gs; ++j) {
                            const uint32_t pk = packed[j >> 1];
                            const float x = __uint_as_float((j & 1) ? (pk & 0xffff0000u) : (pk << 16));
                            s1 += x;
                            s2 += x * x;
                        }
                    }
                    const int grp = (cg + gi * gs) / gs;
                    if (warp_one_image) {
                        s1 = warp_sum(s1);
                        s2 = warp_sum(s2);
                        if (lane == 0 && cg + gi * gs < g.cout) {
                            atomicAdd(&s_stats[ln][grp][0], s1);
                            atomicAdd(&s_stats[ln][grp][1], s2);
                        }
                    } else if (row_ok && cg + gi * gs < g.cout) {
                        atomicAdd(&s_stats[ln][grp][0], s1);
                        atomicAdd(&s_stats[ln][grp][1], s2);
                    }
                }
            }
        }
        // publish the tile's statistics: epilogue warps only (named barrier 1, 128 threads)
        if (p.stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int et = threadIdx.x - 64;      // 0..127
            for (int i = et; i < g.bn * g.groups; i += 128) {
                const int li = i / g.groups, gi = i % g.groups;
                if (n0 + li < g.N) {
                    float* dst = p.stats + ((size_t)(n0 + li) * g.groups + gi) * 2;
                    atomicAdd(dst, s_stats[li][gi][0]);
                    atomicAdd(dst + 1, s_stats[li][gi][1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN);
}

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
};

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
                // straight into the reference layout [Cout][Cin][kh][kw]
                if (co < g.cout) {
                    const int ci0 = ci_t * BNW + c0;
                    float* dst = dw + ((size_t)co * g.cin + ci0) * g.taps + tap;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (ci0 + j < g.cin) atomicAdd(dst + (size_t)j * g.taps, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BNW);
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

int make_act_tmap(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    return sh_make_tmap_bf16(m, base, 4, dims, strides, box);
}

}  // namespace

// Implicit-GEMM convolution, stride 1, 'same' zero padding, NHWC bf16 in, fp32 accumulate.
//   x        bf16 [N,H,W,Cin]           Cin % 64 == 0
//   w        bf16 [taps, cout_pad, Cin] (tap-major, K contiguous); rows >= Cout must be zero; cout_pad % BN == 0
//   y        bf16 [N,H,W,y_ld] or null; y_nchw fp32 [N,Cout,H,W] or null; stats fp32 [N,groups,2] (accumulated) or null
// The data-gradient is the same call on dY with flipped/transposed weights.
SH_EXPORT int sh_conv_fwd(const void* x, const void* w, const void* bias, const void* residual, int N, int H, int W,
                           int Cin, int Cout, int cout_pad, int taps, void* y, int y_ld, void* y_nchw, void* stats,
                           int groups, void* stream) {
    SH_REQUIRE(x && w && (y || y_nchw), "sh_conv_fwd: null pointer");
    SH_REQUIRE(taps == 1 || taps == 9, "sh_conv_fwd: taps must be 1 or 9");
    SH_REQUIRE(N >= 1 && is_pow2(H) && is_pow2(W) && H >= 4 && W >= 4, "sh_conv_fwd: H, W must be powers of two >= 4");
    SH_REQUIRE(Cin % 64 == 0 && Cout >= 1 && cout_pad >= Cout && cout_pad % 64 == 0, "sh_conv_fwd: bad channel counts");
    SH_REQUIRE(!y || (y_ld % 8 == 0 && y_ld >= Cout), "sh_conv_fwd: y_ld must be a multiple of 8 and >= Cout");
    SH_REQUIRE(!residual || Cout % 8 == 0, "sh_conv_fwd: residual needs Cout %% 8 == 0");
    SH_REQUIRE(!stats || (groups > 0 && Cout % groups == 0 && 32 % (Cout / groups) == 0 && groups <= 32),
               "sh_conv_fwd: unsupported GroupNorm grouping");
    cudaStream_t st = (cudaStream_t)stream;
    ConvGeom g;
    g.N = N; g.H = H; g.W = W;
    g.bw = W < 16 ? W : 16;
    g.bh = H < 128 / g.bw ? H : 128 / g.bw;
    g.bn = 128 / (g.bw * g.bh);
    SH_REQUIRE(g.bn <= 8, "sh_conv_fwd: image too small");
    g.tiles_w = W / g.bw; g.tiles_h = H / g.bh;
    g.taps = taps; g.kblocks = Cin / 64; g.cout = Cout; g.cout_pad = cout_pad; g.y_ld = y_ld; g.groups = stats ? groups : 0;
    ConvPtrs p{(const float*)bias, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, (float*)y_nchw, (float*)stats};
    const int bn_tile = (cout_pad % 128 == 0) ? 128 : 64;
    CUtensorMap tmA, tmB;
    int rc = make_act_tmap(&tmA, x, N, H, W, Cin, g.bw, g.bh, g.bn);
    if (rc) return rc;
    const uint64_t wd[2] = {(uint64_t)Cin, (uint64_t)taps * cout_pad};
    const uint64_t ws[1] = {(uint64_t)Cin * 2};
    const uint32_t wb[2] = {64, (uint32_t)bn_tile};
    rc = sh_make_tmap_bf16(&tmB, w, 2, wd, ws, wb);
    if (rc) return rc;
    const int m_tiles = g.tiles_w * g.tiles_h * ((N + g.bn - 1) / g.bn);
    dim3 grid(m_tiles, cout_pad / bn_tile);
    if (bn_tile == 128) {
        static bool attr = false;
        if (!attr) { SH_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<128>::kTotal)); attr = true; }
        conv_gemm_kernel<128><<<grid, kThreads, ConvSmem<128>::kTotal, st>>>(tmA, tmB, g, p);
    } else {
        static bool attr = false;
        if (!attr) { SH_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<64>::kTotal)); attr = true; }
        conv_gemm_kernel<64><<<grid, kThreads, ConvSmem<64>::kTotal, st>>>(tmA, tmB, g, p);
    }
    SH_CHECK_LAUNCH("conv_gemm_kernel");
    return SH_OK;
}

// Weight gradient: dw fp32 [Cout, Cin, k, k] (the reference layout) += dY^T X_shifted, accumulated atomically (zero it
// first).  dy bf16 [N,H,W,dy_C], x bf16 [N,H,W,x_C]; Cout <= dy_C, Cin <= x_C are the real channel counts; channels
// beyond dy_C / x_C needed to fill the 128-wide MMA tile are supplied as zeros by TMA's out-of-bounds fill.
SH_EXPORT int sh_conv_wgrad(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, int taps,
                            void* dw, void* stream) {
    SH_REQUIRE(dy && x && dw, "sh_conv_wgrad: null pointer");
    SH_REQUIRE(taps == 1 || taps == 9, "sh_conv_wgrad: taps must be 1 or 9");
    SH_REQUIRE(N >= 1 && is_pow2(H) && is_pow2(W) && H >= 4 && W >= 4, "sh_conv_wgrad: H, W must be powers of two >= 4");
    SH_REQUIRE(x_C % 64 == 0 && dy_C % 64 == 0 && Cin >= 1 && Cin <= x_C && Cout >= 1 && Cout <= dy_C,
               "sh_conv_wgrad: channel counts must be multiples of 64 in memory");
    cudaStream_t st = (cudaStream_t)stream;
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
    int ksplit = (2 * SH_NUM_SMS + out_tiles - 1) / out_tiles;
    if (ksplit > total_tiles) ksplit = total_tiles;
    if (ksplit < 1) ksplit = 1;
    g.ksplit = ksplit;
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
