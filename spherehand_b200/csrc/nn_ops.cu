// Memory-bound layers of the hourglass (sm_100a): GroupNorm(+ReLU) forward/backward, 2x2 max-pool, bilinear x2
// up-sampling fused with the skip add, the 5x5 stride-2 stem convolution (Cin = 1), layout conversions, weight
// re-packing and the flat Adam step.
//
// Replaces the ATen kernels behind
//   nn.GroupNorm + ReLU(inplace)        /root/reference/network/hourglass.py:26-36, 97, 139-145, 153-155
//   F.max_pool2d(2, 2)                  /root/reference/network/hourglass.py:70, 158
//   F.interpolate(x2, bilinear) + add   /root/reference/network/hourglass.py:79-81
//   conv1 (5x5, stride 2, 1 -> 64)      /root/reference/network/hourglass.py:95-96, 153
//   torch.optim.Adam(lr, wd=1e-5)       /root/reference/network/engine.py:95-97
// All activations are NHWC bf16; every kernel that produces a tensor consumed by a GroupNorm also accumulates that
// tensor's per-(sample, group) sum / sum-of-squares (fp32 atomics into stats[N,G,2]) so normalisation never needs
// a separate reduction pass; kernels that produce an output-gradient tensor optionally accumulate its per-channel
// column sum (the bias gradient of the producing convolution).  These kernels are HBM-bound: 16-byte vector
// accesses (8 bf16 channels per thread), grid = (chunks, N) with >= 2 waves of 148 SMs.
#include <stdlib.h>
#include "tc_common.cuh"

namespace {

constexpr int kT = 256;

struct bf8 { float v[8]; };

__device__ __forceinline__ bf8 load8(const __nv_bfloat16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t r[4] = {u.x, u.y, u.z, u.w};
    bf8 o;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        o.v[2 * e] = __uint_as_float(r[e] << 16);
        o.v[2 * e + 1] = __uint_as_float(r[e] & 0xffff0000u);
    }
    return o;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, bf8& x) {
    uint32_t r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(x.v[2 * e], x.v[2 * e + 1]);
        r[e] = *reinterpret_cast<const uint32_t*>(&b);
        x.v[2 * e] = __uint_as_float(r[e] << 16);              // hand back the rounded values
        x.v[2 * e + 1] = __uint_as_float(r[e] & 0xffff0000u);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
}

// Per-thread running sum / sumsq of its 8 channels (the channel vector of a thread never changes, so its group(s)
// are fixed): st = {s1, s2} for gs >= 8, {s1_lo, s2_lo, s1_hi, s2_hi} for gs == 4.  Flushed once into s_st[G][2].
__device__ __forceinline__ void stats_accum(float* st, const bf8& x, int gs) {
    // packed f32x2 partial sums over the channel pairs (short dependency chains, one issue slot per pair), folded per call
    const float2 p0 = make_float2(x.v[0], x.v[1]), p1 = make_float2(x.v[2], x.v[3]), p2 = make_float2(x.v[4], x.v[5]),
                 p3 = make_float2(x.v[6], x.v[7]);
    const float2 slo = __fadd2_rn(p0, p1), shi = __fadd2_rn(p2, p3);
    const float2 qlo = __ffma2_rn(p1, p1, __fmul2_rn(p0, p0)), qhi = __ffma2_rn(p3, p3, __fmul2_rn(p2, p2));
    if (gs >= 8) {
        const float2 s = __fadd2_rn(slo, shi), q = __fadd2_rn(qlo, qhi);
        st[0] += s.x + s.y;
        st[1] += q.x + q.y;
    } else {
        st[0] += slo.x + slo.y; st[1] += qlo.x + qlo.y;
        st[2] += shi.x + shi.y; st[3] += qhi.x + qhi.y;
    }
}
// Sum `v[8]` (this thread's 8 channels c0..c0+7) over all threads of the block that own the same channels, into
// dst[C] (+=), without shared-memory float atomics (those compile to CAS loops): shuffle over the row slots inside the warp,
// then the warps take turns.  Must be called by every thread of the NT-thread block; C/8 must divide 32.
template <int NT>
__device__ __forceinline__ void block_channel_sum(float* v, float* dst, int vecs, int c0) {
    for (int o = vecs; o < 32; o <<= 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool owner = vecs >= 32 || lane < vecs;
    for (int wi = 0; wi < NT / 32; ++wi) {
        if (warp == wi && owner) {
            float4* d = reinterpret_cast<float4*>(dst + c0);
            float4 a = d[0], b = d[1];
            a.x += v[0]; a.y += v[1]; a.z += v[2]; a.w += v[3];
            b.x += v[4]; b.y += v[5]; b.z += v[6]; b.w += v[7];
            d[0] = a; d[1] = b;
        }
        __syncthreads();
    }
}
// Block-wide reduction of the per-thread statistics of stats_accum and ONE global atomic per (group, moment):
// s_tmp[256] must be zero on entry (and is left dirty); gs = channels per group of the consumer GroupNorm.
template <int NT>
__device__ __forceinline__ void stats_flush_block(const float* st, float* s_tmp, int vecs, int c0, int gs, int G_out,
                                                  float* __restrict__ stats_out_n) {
    float v[8] = {st[0], st[1], st[2], st[3], 0.f, 0.f, 0.f, 0.f};
    block_channel_sum<NT>(v, s_tmp, vecs, c0);
    if (threadIdx.x < G_out) {
        const int g = threadIdx.x;
        float s1 = 0.f, s2 = 0.f;
        if (gs >= 8) {
            const int lpg = gs / 8;                               // channel vectors per group
            for (int cv = g * lpg; cv < (g + 1) * lpg; ++cv) { s1 += s_tmp[cv * 8]; s2 += s_tmp[cv * 8 + 1]; }
        } else {
            s1 = s_tmp[(g >> 1) * 8 + (g & 1) * 2];
            s2 = s_tmp[(g >> 1) * 8 + (g & 1) * 2 + 1];
        }
        atomicAdd(stats_out_n + g * 2, s1);
        atomicAdd(stats_out_n + g * 2 + 1, s2);
    }
}

// Thread mapping shared by the elementwise NHWC kernels: blockIdx.y = sample n, a block walks pixel rows
// [blockIdx.x * ppb, ...) of that sample; lane-in-pixel = channel vector.
struct Walk {
    int vecs;        // C / 8
    int rows;        // pixel rows processed concurrently = kT / vecs
    int cv;          // this thread's channel vector
    int r;           // this thread's row slot
};
__device__ __forceinline__ Walk make_walk(int C) {
    Walk w;
    w.vecs = C / 8;
    w.rows = kT / w.vecs;
    w.cv = threadIdx.x % w.vecs;
    w.r = threadIdx.x / w.vecs;
    return w;
}

// ------------------------------------------------------------------------------------------------ GroupNorm + ReLU
// y = relu((x - mean) * rstd * gamma + beta), statistics from stats_in[N,G,2] over `cnt` = HW*C/G elements.
__global__ void __launch_bounds__(kT) gn_relu_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats_in,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         int HW, int C, int G, int ppb, float eps, __nv_bfloat16* __restrict__ y,
                                                         float* __restrict__ stats_out, int G_out) {
    __shared__ float s_mr[32 * 2];
    __shared__ __align__(16) float s_st[256];
    const int n = blockIdx.y;
    const int gs = C / G;
    const float cnt_inv = 1.f / ((float)HW * gs);
    if (threadIdx.x < G) {
        const float s1 = stats_in[((size_t)n * G + threadIdx.x) * 2], s2 = stats_in[((size_t)n * G + threadIdx.x) * 2 + 1];
        const float mean = s1 * cnt_inv;
        const float var = fmaxf(s2 * cnt_inv - mean * mean, 0.f);
        s_mr[threadIdx.x * 2] = mean;
        s_mr[threadIdx.x * 2 + 1] = rsqrtf(var + eps);
    }
    s_st[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float ka[8], kb[8];                          // y = relu(x * ka + kb)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float mu = s_mr[((c0 + e) / gs) * 2], rs = s_mr[((c0 + e) / gs) * 2 + 1];
        ka[e] = rs * gamma[c0 + e];
        kb[e] = fmaf(-mu, ka[e], beta[c0 + e]);
    }
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    const int p_end = min((int)(blockIdx.x + 1) * ppb, HW);
    constexpr int kU = 4;                        // independent 16-byte loads in flight per thread
    const __nv_bfloat16* xb = x + (size_t)n * HW * C + c0;
    __nv_bfloat16* yb = y + (size_t)n * HW * C + c0;
    for (int p0 = blockIdx.x * ppb + w.r; p0 < p_end; p0 += w.rows * kU) {
        uint4 raw[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int pp = p0 + u * w.rows;
            raw[u] = pp < p_end ? *reinterpret_cast<const uint4*>(xb + (size_t)pp * C) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int pp = p0 + u * w.rows;
            if (pp >= p_end) break;
            const uint32_t r[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            bf8 v;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 y = __ffma2_rn(make_float2(__uint_as_float(r[e] << 16), __uint_as_float(r[e] & 0xffff0000u)),
                                            make_float2(ka[2 * e], ka[2 * e + 1]), make_float2(kb[2 * e], kb[2 * e + 1]));
                v.v[2 * e] = fmaxf(y.x, 0.f);
                v.v[2 * e + 1] = fmaxf(y.y, 0.f);
            }
            store8(yb + (size_t)pp * C, v);
            if (stats_out) stats_accum(st, v, C / G_out);
        }
    }
    if (stats_out) stats_flush_block<kT>(st, s_st, w.vecs, c0, C / G_out, G_out, stats_out + (size_t)n * G_out * 2);
}

// Backward, ONE pass over HBM: a block stages its contiguous [ppb pixels x C] slab of x, da (and the addend) in shared
// memory with 1-D bulk async copies (TMA), so every byte is read once and up to 96 KB per block are in flight.
//   phase 1 (from smem): per-(n,g) sums A = sum dxh, Bq = sum dxh * xhat (dxh = dy*gamma, dy = da*[y>0]) and per-channel
//            dgamma = sum dy*xhat, dbeta = sum dy;  block partials -> fp32 atomics into red[N,G,2], dgamma/dbeta[C];
//   per-sample barrier: the blocks of one sample are adjacent in dispatch order (blockIdx.x fastest) and co-resident
//            (<= 128 blocks per sample vs 592 resident), so a device-scope arrive counter + acquire spin is safe;
//   phase 2 (from smem): dx = rstd * (dxh - A/m - xhat * Bq/m) (+ addend), written in place over the da slab and
//            stored with one bulk async copy; optional column sum of the ROUNDED dx into colsum[C].
constexpr int kBwdMaxElems = 8192;      // slab elements (pixels x channels): 16 KB bf16 per tensor, 48 KB per block
constexpr int kBT = 128;                // threads per block: 4 blocks (16 warps) per SM in different phases
constexpr int kBwdSmallElems = 16384;   // a whole sample of up to this many elements is handled by one block

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// MB = resident blocks per SM the register allocation aims at: 4 with an addend (three 16 KB slabs per block: shared memory caps it
// there anyway), 6 without (two slabs: 80 registers instead of 109 let six blocks = 24 warps share an SM; the kernel is latency-
// bound -- neither the issue slots nor HBM are saturated -- so resident warps are what it is short of).
template <int MB>
__global__ void __launch_bounds__(kBT, MB) gn_relu_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ x,
                                                             const float* __restrict__ stats_in, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const __nv_bfloat16* __restrict__ addend,
                                                             int HW, int C, int G, int ppb, float eps, float* __restrict__ red,
                                                             unsigned* __restrict__ counter, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx,
                                                             float* __restrict__ colsum) {
    extern __shared__ __align__(128) uint8_t bwd_smem[];
    __shared__ float s_mr[32 * 2];
    __shared__ float s_red[32 * 2];
    __shared__ __align__(16) float s_gb[2][256];
    __shared__ __align__(16) float s_cs[256];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * ppb;
    const int npix = min(ppb, HW - p0);
    const uint32_t bytes = (uint32_t)npix * C * 2;
    const uint32_t slab = (uint32_t)ppb * C * 2;
    uint8_t* s_x = bwd_smem;
    uint8_t* s_da = bwd_smem + slab;
    uint8_t* s_add = bwd_smem + 2 * slab;
    const size_t goff = ((size_t)n * HW + p0) * C;
    const int gs = C / G;
    const float cnt_inv = 1.f / ((float)HW * gs);
    if (threadIdx.x == 0) {
        // the slab loads go out before anything else: the statistics fetch and the table clearing below overlap them
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar[0], 2 * bytes);
        for (uint32_t o = 0; o < bytes; o += 4096) {          // 4 KB pieces: several bulk requests in flight per tensor
            const uint32_t b = min(4096u, bytes - o);
            bulk_g2s(s_x + o, reinterpret_cast<const uint8_t*>(x + goff) + o, b, &s_bar[0]);
            bulk_g2s(s_da + o, reinterpret_cast<const uint8_t*>(da + goff) + o, b, &s_bar[0]);
        }
        if (addend) {
            mbar_expect_tx(&s_bar[1], bytes);
            for (uint32_t o = 0; o < bytes; o += 4096)
                bulk_g2s(s_add + o, reinterpret_cast<const uint8_t*>(addend + goff) + o, min(4096u, bytes - o), &s_bar[1]);
        }
    }
    if (threadIdx.x < G) {
        const float s1 = stats_in[((size_t)n * G + threadIdx.x) * 2], s2 = stats_in[((size_t)n * G + threadIdx.x) * 2 + 1];
        const float mean = s1 * cnt_inv;
        const float var = fmaxf(s2 * cnt_inv - mean * mean, 0.f);
        s_mr[threadIdx.x * 2] = mean;
        s_mr[threadIdx.x * 2 + 1] = rsqrtf(var + eps);
    }
    if (threadIdx.x < 64) s_red[threadIdx.x] = 0.f;
    for (int i = threadIdx.x; i < 512; i += kBT) (&s_gb[0][0])[i] = 0.f;
    for (int i = threadIdx.x; i < 256; i += kBT) s_cs[i] = 0.f;
    __syncthreads();
    const int vecs = C / 8, rows = kBT / vecs;
    const int cv = threadIdx.x % vecs, r0 = threadIdx.x / vecs;
    const int c0 = cv * 8;
    // y = x*ka + kb (the ReLU argument), xhat = x*rs + nm
    float ka[8], kb[8], rs[8], nm[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float mu = s_mr[((c0 + e) / gs) * 2];
        rs[e] = s_mr[((c0 + e) / gs) * 2 + 1];
        nm[e] = -mu * rs[e];
        ka[e] = rs[e] * gamma[c0 + e];
        kb[e] = fmaf(nm[e], gamma[c0 + e], beta[c0 + e]);
    }
    mbar_wait(&s_bar[0], 0);
    {
        float dg[8], db[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { dg[e] = 0.f; db[e] = 0.f; }
#pragma unroll 2
        for (int pp = r0; pp < npix; pp += rows) {
            const uint32_t o = ((uint32_t)pp * C + c0) * 2;
            const bf8 xv = load8(reinterpret_cast<const __nv_bfloat16*>(s_x + o));
            const bf8 gv = load8(reinterpret_cast<const __nv_bfloat16*>(s_da + o));
#pragma unroll
            for (int j = 0; j < 4; ++j) {                     // packed f32x2: one issue slot per channel pair, same IEEE results
                const float2 x2 = make_float2(xv.v[2 * j], xv.v[2 * j + 1]);
                const float2 y = __ffma2_rn(x2, make_float2(ka[2 * j], ka[2 * j + 1]), make_float2(kb[2 * j], kb[2 * j + 1]));
                const float2 xh = __ffma2_rn(x2, make_float2(rs[2 * j], rs[2 * j + 1]), make_float2(nm[2 * j], nm[2 * j + 1]));
                const float2 dy = make_float2(y.x > 0.f ? gv.v[2 * j] : 0.f, y.y > 0.f ? gv.v[2 * j + 1] : 0.f);
                const float2 g2 = __ffma2_rn(dy, xh, make_float2(dg[2 * j], dg[2 * j + 1]));
                const float2 b2 = __fadd2_rn(make_float2(db[2 * j], db[2 * j + 1]), dy);
                dg[2 * j] = g2.x; dg[2 * j + 1] = g2.y;
                db[2 * j] = b2.x; db[2 * j + 1] = b2.y;
            }
        }
        block_channel_sum<kBT>(dg, s_gb[0], vecs, c0);
        block_channel_sum<kBT>(db, s_gb[1], vecs, c0);
    }
    // per-group sums from the per-channel ones: A = sum gamma*dbeta_c, Bq = sum gamma*dgamma_c
    if (threadIdx.x < G) {
        float A = 0.f, Bq = 0.f;
        for (int c = threadIdx.x * gs; c < (threadIdx.x + 1) * gs; ++c) {
            const float ga = gamma[c];
            A = fmaf(ga, s_gb[1][c], A);
            Bq = fmaf(ga, s_gb[0][c], Bq);
        }
        s_red[threadIdx.x * 2] = A;
        s_red[threadIdx.x * 2 + 1] = Bq;
    }
    __syncthreads();
    if (gridDim.x > 1) {
        if (threadIdx.x < G * 2) {
            atomicAdd(&red[(size_t)n * G * 2 + threadIdx.x], s_red[threadIdx.x]);
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&counter[n], 1u);
    }
    for (int c = threadIdx.x; c < C; c += kBT) { atomicAdd(&dgamma[c], s_gb[0][c]); atomicAdd(&dbeta[c], s_gb[1][c]); }
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            while (ld_relaxed_u32(&counter[n]) < gridDim.x) __nanosleep(20);
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x < G * 2) s_red[threadIdx.x] = __ldcg(&red[(size_t)n * G * 2 + threadIdx.x]);
        __syncthreads();
    }
    // dx = dy*ka + x*k2 + k3 (+ addend):  rstd*(dy*gamma - A/m - xhat*Bq/m) with xhat = x*rs + nm folded into k2, k3
    float k2[8], k3[8], cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int g = (c0 + e) / gs;
        const float mA = s_red[g * 2] * cnt_inv, mB = s_red[g * 2 + 1] * cnt_inv;
        k2[e] = -rs[e] * rs[e] * mB;
        k3[e] = -rs[e] * fmaf(nm[e], mB, mA);
        cs[e] = 0.f;
    }
    if (addend) mbar_wait(&s_bar[1], 0);
#pragma unroll 2
    for (int pp = r0; pp < npix; pp += rows) {
        const uint32_t o = ((uint32_t)pp * C + c0) * 2;
        const bf8 xv = load8(reinterpret_cast<const __nv_bfloat16*>(s_x + o));
        const bf8 gv = load8(reinterpret_cast<const __nv_bfloat16*>(s_da + o));
        bf8 r;
        if (addend) r = load8(reinterpret_cast<const __nv_bfloat16*>(s_add + o));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 x2 = make_float2(xv.v[2 * j], xv.v[2 * j + 1]);
            const float2 ka2 = make_float2(ka[2 * j], ka[2 * j + 1]);
            const float2 y = __ffma2_rn(x2, ka2, make_float2(kb[2 * j], kb[2 * j + 1]));
            const float2 dy = make_float2(y.x > 0.f ? gv.v[2 * j] : 0.f, y.y > 0.f ? gv.v[2 * j + 1] : 0.f);
            float2 d = __ffma2_rn(dy, ka2, __ffma2_rn(x2, make_float2(k2[2 * j], k2[2 * j + 1]), make_float2(k3[2 * j], k3[2 * j + 1])));
            if (addend) d = __fadd2_rn(d, make_float2(r.v[2 * j], r.v[2 * j + 1]));
            r.v[2 * j] = d.x;
            r.v[2 * j + 1] = d.y;
        }
        store8(reinterpret_cast<__nv_bfloat16*>(s_da + o), r);
        if (colsum) {
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[e] += r.v[e];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t o = 0; o < bytes; o += 4096)
            bulk_s2g(reinterpret_cast<uint8_t*>(dx + goff) + o, s_da + o, min(4096u, bytes - o));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (colsum) {
        block_channel_sum<kBT>(cs, s_cs, vecs, c0);
        for (int c = threadIdx.x; c < C; c += kBT) atomicAdd(&colsum[c], s_cs[c]);
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ pooling / up-sampling
// y[n,h,w,:] = max over the 2x2 window of x; statistics of y; x is [N,2H,2W,C].
__global__ void __launch_bounds__(kT) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int ppb,
                                                         __nv_bfloat16* __restrict__ y, float* __restrict__ stats_out, int G_out) {
    __shared__ __align__(16) float s_st[256];
    const int n = blockIdx.y, HW = H * W;
    s_st[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    const int p_end = min((int)(blockIdx.x + 1) * ppb, HW);
    for (int pp = blockIdx.x * ppb + w.r; pp < p_end; pp += w.rows) {
        const int h = pp / W, ww = pp % W;
        const __nv_bfloat16* src = x + (((size_t)n * 2 * H + 2 * h) * 2 * W + 2 * ww) * C + c0;
        const bf8 a = load8(src), b = load8(src + C), c = load8(src + (size_t)2 * W * C), d = load8(src + (size_t)2 * W * C + C);
        bf8 r;
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = fmaxf(fmaxf(a.v[e], b.v[e]), fmaxf(c.v[e], d.v[e]));
        store8(y + ((size_t)n * HW + pp) * C + c0, r);
        if (stats_out) stats_accum(st, r, C / G_out);
    }
    if (stats_out) stats_flush_block<kT>(st, s_st, w.vecs, c0, C / G_out, G_out, stats_out + (size_t)n * G_out * 2);
}

// dx[n,2h+i,2w+j,:] (+)= dy[n,h,w,:] for the FIRST maximal element of the window in (0,0),(0,1),(1,0),(1,1) order
// (ATen's max_pool2d backward routes to the first maximum), zero elsewhere; optional addend (another gradient
// flowing into x) and column sum.
__global__ void __launch_bounds__(kT) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ addend, int H, int W, int C, int ppb,
                                                         __nv_bfloat16* __restrict__ dx, float* __restrict__ colsum) {
    __shared__ __align__(16) float s_cs[256];
    const int n = blockIdx.y, HW = H * W;
    if (threadIdx.x < 256) s_cs[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
    const int p_end = min((int)(blockIdx.x + 1) * ppb, HW);
    for (int pp = blockIdx.x * ppb + w.r; pp < p_end; pp += w.rows) {
        const int h = pp / W, ww = pp % W;
        const size_t base = (((size_t)n * 2 * H + 2 * h) * 2 * W + 2 * ww) * C + c0;
        const size_t offs[4] = {0, (size_t)C, (size_t)2 * W * C, (size_t)2 * W * C + C};
        bf8 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = load8(x + base + offs[q]);
        const bf8 g = load8(dy + ((size_t)n * HW + pp) * C + c0);
        bf8 o[4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int best = 0;
            float bv = v[0].v[e];
#pragma unroll
            for (int q = 1; q < 4; ++q) if (v[q].v[e] > bv) { bv = v[q].v[e]; best = q; }
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q].v[e] = (q == best) ? g.v[e] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (addend) {
                const bf8 r = load8(addend + base + offs[q]);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[q].v[e] += r.v[e];
            }
            store8(dx + base + offs[q], o[q]);
            if (colsum) {
#pragma unroll
                for (int e = 0; e < 8; ++e) cs[e] += o[q].v[e];
            }
        }
    }
    if (colsum) {
        block_channel_sum<kT>(cs, s_cs, w.vecs, c0);
        for (int c = threadIdx.x; c < C; c += kT) atomicAdd(&colsum[c], s_cs[c]);
    }
}

// y = up1 + bilinear_x2(low), align_corners=False: src = (dst + 0.5)/2 - 0.5 clamped -> weights (0.75, 0.25);
// y is [N,2h,2w,C], low is [N,h,w,C].
__device__ __forceinline__ void up_taps(int d, int n_src, int& i0, int& i1, float& w0, float& w1) {
    // d even: src = d/2 - 0.25 -> i0 = d/2 - 1 (w 0.25), i1 = d/2 (w 0.75); d odd: i0 = d/2 (0.75), i1 = d/2 + 1 (0.25)
    const int hlf = d >> 1;
    if (d & 1) { i0 = hlf; i1 = min(hlf + 1, n_src - 1); w0 = 0.75f; w1 = 0.25f; }
    else { i0 = max(hlf - 1, 0); i1 = hlf; w0 = 0.25f; w1 = 0.75f; }
}

__global__ void __launch_bounds__(kT) upsample_add_fwd_kernel(const __nv_bfloat16* __restrict__ up1, const __nv_bfloat16* __restrict__ low,
                                                              int h, int w_, int C, int ppb, __nv_bfloat16* __restrict__ y,
                                                              float* __restrict__ stats_out, int G_out) {
    __shared__ __align__(16) float s_st[256];
    const int n = blockIdx.y, H = 2 * h, W = 2 * w_, HW = H * W;
    s_st[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    const int p_end = min((int)(blockIdx.x + 1) * ppb, HW);
    const __nv_bfloat16* lb = low + (size_t)n * h * w_ * C + c0;
    constexpr int kU = 2;                        // output pixels in flight per thread (10 independent 16-byte loads)
    for (int p0 = blockIdx.x * ppb + w.r; p0 < p_end; p0 += w.rows * kU) {
        uint4 ra[kU], rb[kU], rc[kU], rd[kU], ru[kU];
        float wy0[kU], wy1[kU], wx0[kU], wx1[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int pp = min(p0 + u * w.rows, p_end - 1);            // clamped: the duplicate is not stored
            const int yy = pp / W, xx = pp % W;
            int y0, y1, x0, x1;
            up_taps(yy, h, y0, y1, wy0[u], wy1[u]);
            up_taps(xx, w_, x0, x1, wx0[u], wx1[u]);
            ra[u] = *reinterpret_cast<const uint4*>(lb + ((size_t)y0 * w_ + x0) * C);
            rb[u] = *reinterpret_cast<const uint4*>(lb + ((size_t)y0 * w_ + x1) * C);
            rc[u] = *reinterpret_cast<const uint4*>(lb + ((size_t)y1 * w_ + x0) * C);
            rd[u] = *reinterpret_cast<const uint4*>(lb + ((size_t)y1 * w_ + x1) * C);
            ru[u] = *reinterpret_cast<const uint4*>(up1 + ((size_t)n * HW + pp) * C + c0);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int pp = p0 + u * w.rows;
            if (pp >= p_end) break;
            const uint32_t wa[4] = {ra[u].x, ra[u].y, ra[u].z, ra[u].w}, wb[4] = {rb[u].x, rb[u].y, rb[u].z, rb[u].w};
            const uint32_t wc[4] = {rc[u].x, rc[u].y, rc[u].z, rc[u].w}, wd[4] = {rd[u].x, rd[u].y, rd[u].z, rd[u].w};
            const uint32_t wu[4] = {ru[u].x, ru[u].y, ru[u].z, ru[u].w};
            bf8 r;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a0 = __uint_as_float(wa[e] << 16), a1 = __uint_as_float(wa[e] & 0xffff0000u);
                const float b0 = __uint_as_float(wb[e] << 16), b1 = __uint_as_float(wb[e] & 0xffff0000u);
                const float c0f = __uint_as_float(wc[e] << 16), c1f = __uint_as_float(wc[e] & 0xffff0000u);
                const float d0 = __uint_as_float(wd[e] << 16), d1 = __uint_as_float(wd[e] & 0xffff0000u);
                // up1 + wy0*(wx0*a + wx1*b) + wy1*(wx0*c + wx1*d), packed f32x2 (one issue slot per channel pair; the kernel is
                // instruction-issue bound: ~150 instructions per 16 bytes of output)
                const float2 X0 = make_float2(wx0[u], wx0[u]), X1 = make_float2(wx1[u], wx1[u]);
                const float2 top = __ffma2_rn(X1, make_float2(b0, b1), __fmul2_rn(X0, make_float2(a0, a1)));
                const float2 bot = __ffma2_rn(X1, make_float2(d0, d1), __fmul2_rn(X0, make_float2(c0f, c1f)));
                const float2 mix = __ffma2_rn(make_float2(wy0[u], wy0[u]), top, __fmul2_rn(make_float2(wy1[u], wy1[u]), bot));
                const float2 o2 = __fadd2_rn(make_float2(__uint_as_float(wu[e] << 16), __uint_as_float(wu[e] & 0xffff0000u)), mix);
                r.v[2 * e] = o2.x;
                r.v[2 * e + 1] = o2.y;
            }
            store8(y + ((size_t)n * HW + pp) * C + c0, r);
            if (stats_out) stats_accum(st, r, C / G_out);
        }
    }
    if (stats_out) stats_flush_block<kT>(st, s_st, w.vecs, c0, C / G_out, G_out, stats_out + (size_t)n * G_out * 2);
}

// dlow[n,i,j,:] = sum over the (up to 4x4) fine pixels that read (i,j) of weight * dy  (transpose of the above), + column sum
__global__ void __launch_bounds__(kT) upsample_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int h, int w_, int C, int ppb,
                                                          __nv_bfloat16* __restrict__ dlow, float* __restrict__ colsum) {
    __shared__ __align__(16) float s_cs[256];
    const int n = blockIdx.y, hw = h * w_, H = 2 * h, W = 2 * w_;
    if (threadIdx.x < 256) s_cs[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
    const int p_end = min((int)(blockIdx.x + 1) * ppb, hw);
    for (int pp = blockIdx.x * ppb + w.r; pp < p_end; pp += w.rows) {
        const int i = pp / w_, j = pp % w_;
        // the fine rows 2i-1 .. 2i+2 read low row i with weights (0.25, 0.75, 0.75, 0.25); at the image border the clamped tap
        // folds the missing neighbour's weight into the edge row (1.0 instead of 0.75).  Same for columns: 16 independent loads.
        float wy[4] = {0.25f, 0.75f, 0.75f, 0.25f}, wx[4] = {0.25f, 0.75f, 0.75f, 0.25f};
        if (i == 0) { wy[0] = 0.f; wy[1] = 1.f; }
        if (i == h - 1) { wy[3] = 0.f; wy[2] = 1.f; }
        if (j == 0) { wx[0] = 0.f; wx[1] = 1.f; }
        if (j == w_ - 1) { wx[3] = 0.f; wx[2] = 1.f; }
        const __nv_bfloat16* base = dy + (size_t)n * H * W * C + c0;
        uint4 raw[16];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = min(max(2 * i - 1 + a, 0), H - 1);          // clamped rows / columns carry weight 0
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int xx = min(max(2 * j - 1 + b, 0), W - 1);
                raw[a * 4 + b] = *reinterpret_cast<const uint4*>(base + ((size_t)yy * W + xx) * C);
            }
        }
        bf8 acc;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc.v[e] = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float wgt = wy[a] * wx[b];
                const uint32_t r[4] = {raw[a * 4 + b].x, raw[a * 4 + b].y, raw[a * 4 + b].z, raw[a * 4 + b].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc.v[2 * e] = fmaf(wgt, __uint_as_float(r[e] << 16), acc.v[2 * e]);
                    acc.v[2 * e + 1] = fmaf(wgt, __uint_as_float(r[e] & 0xffff0000u), acc.v[2 * e + 1]);
                }
            }
        }
        store8(dlow + ((size_t)n * hw + pp) * C + c0, acc);
        if (colsum) {
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[e] += acc.v[e];
        }
    }
    if (colsum) {
        block_channel_sum<kT>(cs, s_cs, w.vecs, c0);
        for (int c = threadIdx.x; c < C; c += kT) atomicAdd(&colsum[c], s_cs[c]);
    }
}

// y = a + b (+ c), optional statistics of y and column sum of y (used both for activations and for gradient fan-in)
__global__ void __launch_bounds__(kT) add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                 const __nv_bfloat16* __restrict__ c, int HW, int C, int ppb, __nv_bfloat16* __restrict__ y,
                                                 float* __restrict__ stats_out, int G_out, float* __restrict__ colsum) {
    __shared__ __align__(16) float s_st[256];
    __shared__ __align__(16) float s_cs[256];
    const int n = blockIdx.y;
    s_st[threadIdx.x] = 0.f;
    s_cs[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    const int p_end = min((int)(blockIdx.x + 1) * ppb, HW);
    for (int pp = blockIdx.x * ppb + w.r; pp < p_end; pp += w.rows) {
        const size_t o = ((size_t)n * HW + pp) * C + c0;
        bf8 r = load8(a + o);
        const bf8 s = load8(b + o);
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] += s.v[e];
        if (c) {
            const bf8 t = load8(c + o);
#pragma unroll
            for (int e = 0; e < 8; ++e) r.v[e] += t.v[e];
        }
        store8(y + o, r);
        if (stats_out) stats_accum(st, r, C / G_out);
        if (colsum) {
#pragma unroll
            for (int e = 0; e < 8; ++e) cs[e] += r.v[e];
        }
    }
    if (stats_out) stats_flush_block<kT>(st, s_st, w.vecs, c0, C / G_out, G_out, stats_out + (size_t)n * G_out * 2);
    if (colsum) {
        block_channel_sum<kT>(cs, s_cs, w.vecs, c0);
        for (int ch = threadIdx.x; ch < C; ch += kT) atomicAdd(&colsum[ch], s_cs[ch]);
    }
}

// column sum of a bf16 NHWC tensor (bias gradient when no producer kernel could fuse it)
__global__ void __launch_bounds__(kT) colsum_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int ppb, float* __restrict__ colsum) {
    __shared__ __align__(16) float s_cs[256];
    const int n = blockIdx.y;
    if (threadIdx.x < 256) s_cs[threadIdx.x] = 0.f;
    __syncthreads();
    const Walk w = make_walk(C);
    const int c0 = w.cv * 8;
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.f;
    const int p_end = min((int)(blockIdx.x + 1) * ppb, HW);
    constexpr int kU = 4;                        // independent 16-byte loads in flight per thread (a pure read: nothing else hides them)
    for (int p0 = blockIdx.x * ppb + w.r; p0 < p_end; p0 += w.rows * kU) {
        uint4 raw[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u)
            raw[u] = *reinterpret_cast<const uint4*>(x + ((size_t)n * HW + min(p0 + u * w.rows, p_end - 1)) * C + c0);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (p0 + u * w.rows >= p_end) break;                        // (the clamped duplicate is not added)
            const uint32_t wv[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                cs[2 * e] += __uint_as_float(wv[e] << 16);
                cs[2 * e + 1] += __uint_as_float(wv[e] & 0xffff0000u);
            }
        }
    }
    block_channel_sum<kT>(cs, s_cs, w.vecs, c0);
    for (int ch = threadIdx.x; ch < C; ch += kT) atomicAdd(&colsum[ch], s_cs[ch]);
}

// ------------------------------------------------------------------------------------------------ stem: 5x5 stride-2 conv, Cin = 1
// y[n,oh,ow,co] = b[co] + sum_{kh,kw} img[n, 2oh+kh-2, 2ow+kw-2] * w[co,kh,kw];  img fp32 [N,S,S]; y bf16 [N,S/2,S/2,64] + stats.
// K = 25 is too thin for the tensor cores; register tiling on the CUDA cores instead: a block owns 2 output rows x 64
// output columns of one image, the 7 x 132 input patch and the [tap][co] weights sit in shared memory, a thread computes
// 4 adjacent output pixels x 8 channels (32 accumulators; per kernel row 11 input loads feed 160 FMAs).
constexpr int kStemCols = 64;                 // output columns per block
constexpr int kStemRows = 2;                  // output rows per block
constexpr int kStemPatchW = 2 * kStemCols + 4;
constexpr int kStemUnits = 4;                 // row pairs per block: the [tap][co] weights are gathered once for all of them
__global__ void __launch_bounds__(kT, 3) stem_conv_fwd_kernel(const float* __restrict__ img, const float* __restrict__ wgt,
                                                           const float* __restrict__ bias, int S, __nv_bfloat16* __restrict__ y,
                                                           float* __restrict__ stats_out, int G_out) {
    // weights as [tap][half][cv][4]: a thread's two float4 (channels c0..c0+3, c0+4..c0+7) are each one conflict-free 128-byte
    // wavefront across the 8 channel vectors (the [tap][co] form put cv and cv+4 on the same banks: 13 M conflicts per launch)
    __shared__ __align__(16) float s_w[25 * 64];
    __shared__ float s_in[(2 * kStemRows + 3) * kStemPatchW];
    __shared__ float s_st[8][8];
    const int n = blockIdx.z, O = S / 2;
    const int ow0 = blockIdx.x * kStemCols;
    for (int i = threadIdx.x; i < 25 * 64; i += kT) {
        const int tap = i >> 6, co = i & 63;
        s_w[tap * 64 + ((co >> 2) & 1) * 32 + (co >> 3) * 4 + (co & 3)] = wgt[co * 25 + tap];
    }
    const int cv = threadIdx.x & 7, q = threadIdx.x >> 3;            // 8 channel vectors x 32 pixel quads
    const int c0 = cv * 8;
    const int lr = q / (kStemCols / 4), lc = (q % (kStemCols / 4)) * 4;   // local output row, first local output column
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    for (int u = 0; u < kStemUnits; ++u) {
        const int oh0 = (blockIdx.y * kStemUnits + u) * kStemRows;
        if (oh0 >= O) break;                                         // block-uniform
        __syncthreads();                                             // the previous unit's patch has been consumed (and s_w is written)
        for (int i = threadIdx.x; i < (2 * kStemRows + 3) * kStemPatchW; i += kT) {
            const int r = i / kStemPatchW, c = i - r * kStemPatchW;
            const int ih = 2 * oh0 - 2 + r, iw = 2 * ow0 - 2 + c;
            s_in[i] = (ih >= 0 && ih < S && iw >= 0 && iw < S) ? __ldg(img + ((size_t)n * S + ih) * S + iw) : 0.f;
        }
        __syncthreads();
        float2 acc2[4][4];                                           // [pixel][channel pair]: packed FFMA2, one issue slot per pair
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2[i][e] = make_float2(bias[c0 + 2 * e], bias[c0 + 2 * e + 1]);
#pragma unroll
        for (int kh = 0; kh < 5; ++kh) {
            float in[11];
            const float* row = s_in + (2 * lr + kh) * kStemPatchW + 2 * lc;
#pragma unroll
            for (int j = 0; j < 11; ++j) in[j] = row[j];
#pragma unroll
            for (int kw = 0; kw < 5; ++kw) {
                const float4 w0 = *reinterpret_cast<const float4*>(s_w + (kh * 5 + kw) * 64 + cv * 4);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w + (kh * 5 + kw) * 64 + 32 + cv * 4);
                const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 iv = make_float2(in[2 * i + kw], in[2 * i + kw]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc2[i][e] = __ffma2_rn(iv, wv[e], acc2[i][e]);
                }
            }
        }
        const int oh = oh0 + lr;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ow = ow0 + lc + i;
            if (oh < O && ow < O) {
                bf8 v;
#pragma unroll
                for (int e = 0; e < 4; ++e) { v.v[2 * e] = acc2[i][e].x; v.v[2 * e + 1] = acc2[i][e].y; }
                store8(y + (((size_t)n * O + oh) * O + ow) * 64 + c0, v);
                if (stats_out) stats_accum(st, v, 64 / G_out);
            }
        }
    }
    if (stats_out) {
        // thread-private sums -> block sums without shared float atomics.  gs = 64/G_out >= 8: a thread's 8 channels
        // lie in ONE group (st[0], st[1]); lanes that share the group: same cv / (gs/8).
        const int gs = 64 / G_out;
        const int lanes_per_group = gs / 8;                           // consecutive cv values in one group (1, 2, 4 or 8)
        float s1 = st[0], s2 = st[1];
        for (int o = 1; o < lanes_per_group; o <<= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        for (int o = 8; o < 32; o <<= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane < 8 && (lane % lanes_per_group) == 0) {
            s_st[warp][(lane / lanes_per_group) * 2 % 8] = s1;        // G_out <= 4 here (8 slots = 4 groups x 2)
            s_st[warp][((lane / lanes_per_group) * 2 + 1) % 8] = s2;
        }
        __syncthreads();
        if (threadIdx.x < G_out * 2) {
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) t += s_st[w8][threadIdx.x];
            atomicAdd(&stats_out[(size_t)n * G_out * 2 + threadIdx.x], t);
        }
    }
}

// dW[co,kh,kw] = sum_{n,oh,ow} dy[n,oh,ow,co] * img[n,2oh+kh-2,2ow+kw-2];  db[co] = sum dy.   dw: fp32 [64,25], db: [64]
// Persistent blocks walk (image, 4 x 64 output-pixel band) units: the 11 x 132 input patch of a unit is staged in shared memory
// once (zero-filled borders: no per-tap bounds checks), a warp walks 32 consecutive output pixels of one row, lane l owns
// channels 2l, 2l+1 (the 128-byte dy row of a pixel is one coalesced load) and keeps its 25 x 2 partial sums in registers
// across ALL units; per pixel 15 broadcast shared-memory loads feed 25 packed FFMA2.  The earlier version (25 predicated
// global loads per pixel, one block per band with 1664 global reductions each) ran 311 M instructions for 1.7 G MACs.
constexpr int kStemWT = 256;
constexpr int kSWRows = 4;                    // output rows per unit
constexpr int kSWPatchW = 2 * kStemCols + 4;  // 132
constexpr int kSWPatchH = 2 * kSWRows + 3;    // 11
__global__ void __launch_bounds__(kStemWT, 2) stem_conv_wgrad_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ dy,
                                                                     int N, int S, float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float s_dw[26 * 64];         // [tap | bias][co]
    __shared__ __align__(16) float s_in[kSWPatchH * kSWPatchW];
    const int O = S / 2;
    const int cblocks = (O + kStemCols - 1) / kStemCols, bands = (O + kSWRows - 1) / kSWRows;
    const int units = N * bands * cblocks;
    for (int i = threadIdx.x; i < 26 * 64; i += kStemWT) s_dw[i] = 0.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lr = warp >> 1, lc0 = (warp & 1) * 32;                  // the warp's output row / first column inside the unit
    float2 acc[25], bsum = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 25; ++t) acc[t] = make_float2(0.f, 0.f);
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int cb = unit % cblocks, band = (unit / cblocks) % bands, n = unit / (cblocks * bands);
        const int oh0 = band * kSWRows, ow0 = cb * kStemCols;
        __syncthreads();                                              // the previous unit's patch is no longer read
        const float* im = img + (size_t)n * S * S;
        for (int i = threadIdx.x; i < kSWPatchH * kSWPatchW; i += kStemWT) {
            const int r = i / kSWPatchW, c = i - r * kSWPatchW;
            const int ih = 2 * oh0 - 2 + r, iw = 2 * ow0 - 2 + c;
            s_in[i] = (ih >= 0 && ih < S && iw >= 0 && iw < S) ? __ldg(im + (size_t)ih * S + iw) : 0.f;
        }
        __syncthreads();
        const int oh = oh0 + lr;
        const uint32_t* dyw = reinterpret_cast<const uint32_t*>(dy + (((size_t)n * O + oh) * O + ow0 + lc0) * 64) + lane;
        constexpr int kU = 4;                                         // pixels in flight per warp
#pragma unroll 1
        for (int j = 0; j < 32; j += kU) {
            uint32_t raw[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) raw[u] = (oh < O && ow0 + lc0 + j + u < O) ? __ldg(dyw + (size_t)(j + u) * 32) : 0u;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const float2 g = make_float2(__uint_as_float(raw[u] << 16), __uint_as_float(raw[u] & 0xffff0000u));
                const float* pr = s_in + (2 * lr) * kSWPatchW + 2 * (lc0 + j + u);       // even index: 8-byte aligned rows
#pragma unroll
                for (int kh = 0; kh < 5; ++kh) {
                    const float2 a = *reinterpret_cast<const float2*>(pr + kh * kSWPatchW);
                    const float2 b2 = *reinterpret_cast<const float2*>(pr + kh * kSWPatchW + 2);
                    const float c4 = pr[kh * kSWPatchW + 4];
                    acc[kh * 5 + 0] = __ffma2_rn(make_float2(a.x, a.x), g, acc[kh * 5 + 0]);
                    acc[kh * 5 + 1] = __ffma2_rn(make_float2(a.y, a.y), g, acc[kh * 5 + 1]);
                    acc[kh * 5 + 2] = __ffma2_rn(make_float2(b2.x, b2.x), g, acc[kh * 5 + 2]);
                    acc[kh * 5 + 3] = __ffma2_rn(make_float2(b2.y, b2.y), g, acc[kh * 5 + 3]);
                    acc[kh * 5 + 4] = __ffma2_rn(make_float2(c4, c4), g, acc[kh * 5 + 4]);
                }
                bsum = __fadd2_rn(bsum, g);
            }
        }
    }
    __syncthreads();
    // block reduction over the 8 warps: the warps take turns on the shared table (no shared float atomics)
    for (int wi = 0; wi < kStemWT / 32; ++wi) {
        if (warp == wi) {
#pragma unroll
            for (int t = 0; t < 25; ++t) {
                s_dw[t * 64 + 2 * lane] += acc[t].x;
                s_dw[t * 64 + 2 * lane + 1] += acc[t].y;
            }
            s_dw[25 * 64 + 2 * lane] += bsum.x;
            s_dw[25 * 64 + 2 * lane + 1] += bsum.y;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 25 * 64; i += kStemWT) atomicAdd(&dw[(i % 64) * 25 + i / 64], s_dw[i]);
    if (threadIdx.x < 64) atomicAdd(&db[threadIdx.x], s_dw[25 * 64 + threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ layout / weights / optimiser
// fp32 NCHW [N,C,H,W] -> bf16 NHWC [N,H,W,Cp] (channels >= C zero-filled)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int HW, int Cp, __nv_bfloat16* __restrict__ y) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? x[((size_t)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < Cp) y[((size_t)n * HW + p) * Cp + c] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}

// bf16 NHWC [N,H,W,C] -> fp32 NCHW [N,C,H,W]
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int N, int C, int HW, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(x[((size_t)n * HW + p) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) y[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
    }
}

// fp32 master weight [Cout,Cin,kh,kw] -> bf16 forward layout wf[tap][cout_pad][cin_pad] and (optionally) the
// data-gradient layout wb[tap'][cin_pad_rows][cout_pad_cols] = w[co][ci][K-1-kh][K-1-kw]  (flipped + transposed)
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int cout_pad, int cin_pad,
                                    int b_rows, int b_cols, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wb) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nf = (long)taps * cout_pad * cin_pad;
    if (i < nf) {
        const int ci = (int)(i % cin_pad), co = (int)((i / cin_pad) % cout_pad), tap = (int)(i / ((long)cin_pad * cout_pad));
        const float v = (co < Cout && ci < Cin) ? w[((size_t)co * Cin + ci) * taps + tap] : 0.f;
        wf[i] = __float2bfloat16(v);
    }
    const long nb = (long)taps * b_rows * b_cols;
    if (wb && i < nb) {
        const int co = (int)(i % b_cols), ci = (int)((i / b_cols) % b_rows), tap = (int)(i / ((long)b_cols * b_rows));
        const float v = (co < Cout && ci < Cin) ? w[((size_t)co * Cin + ci) * taps + (taps - 1 - tap)] : 0.f;
        wb[i] = __float2bfloat16(v);
    }
}

// All convolutions of a network in ONE launch: blockIdx.y = table row (src offset into the flat fp32 parameter buffer,
// Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols, wf offset, wb offset (-1 = none) into the bf16 arena).
__global__ void pack_weights_batch_kernel(const float* __restrict__ flat, const int* __restrict__ table,
                                          __nv_bfloat16* __restrict__ arena) {
    const int* t = table + blockIdx.y * 10;
    const float* w = flat + t[0];
    const int Cout = t[1], Cin = t[2], taps = t[3], cout_pad = t[4], cin_pad = t[5], b_rows = t[6], b_cols = t[7];
    __nv_bfloat16* wf = arena + t[8];
    const long nf = (long)taps * cout_pad * cin_pad;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_pad), co = (int)((i / cin_pad) % cout_pad), tap = (int)(i / ((long)cin_pad * cout_pad));
        wf[i] = __float2bfloat16((co < Cout && ci < Cin) ? w[((size_t)co * Cin + ci) * taps + tap] : 0.f);
    }
    if (t[9] >= 0) {
        __nv_bfloat16* wb = arena + t[9];
        const long nb = (long)taps * b_rows * b_cols;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (long)gridDim.x * blockDim.x) {
            const int co = (int)(i % b_cols), ci = (int)((i / b_cols) % b_rows), tap = (int)(i / ((long)b_cols * b_rows));
            wb[i] = __float2bfloat16((co < Cout && ci < Cin) ? w[((size_t)co * Cin + ci) * taps + (taps - 1 - tap)] : 0.f);
        }
    }
}

// dW fp32 [tap][Cout][Cin] (tensor-core accumulation layout) -> grad fp32 [Cout][Cin][kh][kw] (the reference layout)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, int Cout, int Cin, int taps, int cout_ld, int cin_ld,
                                    float* __restrict__ grad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)Cout * Cin * taps) return;
    const int tap = (int)(i % taps), ci = (int)((i / taps) % Cin), co = (int)(i / ((long)taps * Cin));
    grad[i] = dw[((size_t)tap * cout_ld + co) * cin_ld + ci];
}

// Adam with L2 weight decay folded into the gradient (torch.optim.Adam semantics, engine.py:95-97), flat buffers.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                            float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float gscale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pi = p[i];
    const float gi = g[i] * gscale + wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

inline int pick_ppb(int HW, int N, int rows) {
    // pixels per block: aim for >= 2 waves over 148 SMs x 8 resident CTAs, at least 4 iterations per thread row
    int ppb = HW;
    while (ppb > rows * 4 && (long)N * ((HW + ppb - 1) / ppb) < 2L * SH_NUM_SMS * 8) ppb = (ppb + 1) / 2;
    return ppb;
}

}  // namespace

#define NHWC_CHECK(name)                                                                                    \
    SH_REQUIRE(C % 8 == 0 && C <= 256 && kT % (C / 8) == 0, name ": C must be a multiple of 8, <= 256 and divide 2048")

SH_EXPORT int sh_gn_relu_fwd(const void* x, const void* stats_in, const void* gamma, const void* beta, int N, int HW, int C,
                              int G, float eps, void* y, void* stats_out, int G_out, void* stream) {
    SH_REQUIRE(x && stats_in && gamma && beta && y, "sh_gn_relu_fwd: null pointer");
    NHWC_CHECK("sh_gn_relu_fwd");
    SH_REQUIRE(G >= 1 && G <= 32 && C % G == 0 && (!stats_out || (G_out >= 1 && G_out <= 32 && C % G_out == 0 && (C / G_out) % 4 == 0)),
               "sh_gn_relu_fwd: bad grouping");
    if (N == 0) return SH_OK;
    const int ppb = pick_ppb(HW, N, kT / (C / 8));
    dim3 grid(sh_div_up(HW, ppb), N);
    gn_relu_fwd_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const float*)stats_in, (const float*)gamma,
                                                              (const float*)beta, HW, C, G, ppb, eps, (__nv_bfloat16*)y,
                                                              (float*)stats_out, G_out);
    SH_CHECK_LAUNCH("gn_relu_fwd_kernel");
    return SH_OK;
}

// Scratch words (4 bytes each) sh_gn_relu_bwd needs in `red`: the per-(n,g) sums plus one arrive counter per sample.
SH_EXPORT size_t sh_gn_relu_bwd_scratch_words(int N, int G) { return (size_t)N * G * 2 + (size_t)N; }

// red: scratch of sh_gn_relu_bwd_scratch_words(N,G) words (zeroed here); dgamma/dbeta accumulated (caller zeroes once per step)
static int gn_relu_bwd_impl(const void* da, const void* x, const void* stats_in, const void* gamma, const void* beta,
                            const void* addend, int N, int HW, int C, int G, float eps, void* red, void* dgamma, void* dbeta,
                            void* dx, void* colsum, void* stream, bool zero_red) {
    SH_REQUIRE(da && x && stats_in && gamma && beta && red && dgamma && dbeta && dx, "sh_gn_relu_bwd: null pointer");
    NHWC_CHECK("sh_gn_relu_bwd");
    SH_REQUIRE(G >= 1 && G <= 32 && C % G == 0 && (C / G) % 4 == 0, "sh_gn_relu_bwd: bad grouping");
    if (N == 0) return SH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SH_REQUIRE(kBT % (C / 8) == 0, "sh_gn_relu_bwd: C / 8 must divide 128");
    const int rows = kBT / (C / 8);
    // pixels per block: the largest slab that fits (16 KB per tensor), halved while that leaves SMs idle.  A sample of up to
    // kBwdSmallElems elements (the 4x4 and 8x8 levels of the hourglass) stays in ONE block: 96 KB of slabs, two blocks per
    // SM, and no cross-block publish / wait chain at all (that chain, not the data, is what a 14 us launch on 8 MB was made of).
    int ppb = kBwdMaxElems / C;
    if ((long)HW * C <= kBwdSmallElems) {
        ppb = HW;
    } else {
        if (ppb > HW) ppb = HW;
        while (ppb > 2 * rows && (long)N * ((HW + ppb - 1) / ppb) < 2L * SH_NUM_SMS) ppb = (ppb + 1) / 2;
    }
    const int nblk = sh_div_up(HW, ppb);
    SH_REQUIRE(nblk <= 128, "sh_gn_relu_bwd: sample too large for the co-resident per-sample barrier (HW*C <= 2M elements)");
    if (zero_red) SH_CUDA(cudaMemsetAsync(red, 0, sh_gn_relu_bwd_scratch_words(N, G) * 4, st));
    const size_t smem = (size_t)ppb * C * 2 * (addend ? 3 : 2);
    static bool attr = false;
    if (!attr) {
        SH_CUDA(cudaFuncSetAttribute(gn_relu_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kBwdSmallElems * 2));
        SH_CUDA(cudaFuncSetAttribute(gn_relu_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kBwdSmallElems * 2));
        attr = true;
    }
    dim3 grid(nblk, N);
    const char* mb_env = getenv("SH_GN_BWD_MB");            // tuning probe: 4 forces the four-block build everywhere
    const bool six = !addend && smem <= 36 * 1024 && !(mb_env && atoi(mb_env) == 4);
    if (six)
        gn_relu_bwd_kernel<6><<<grid, kBT, smem, st>>>((const __nv_bfloat16*)da, (const __nv_bfloat16*)x, (const float*)stats_in,
                                                      (const float*)gamma, (const float*)beta, (const __nv_bfloat16*)addend, HW, C, G, ppb,
                                                      eps, (float*)red, (unsigned*)red + (size_t)N * G * 2, (float*)dgamma, (float*)dbeta,
                                                      (__nv_bfloat16*)dx, (float*)colsum);
    else
        gn_relu_bwd_kernel<4><<<grid, kBT, smem, st>>>((const __nv_bfloat16*)da, (const __nv_bfloat16*)x, (const float*)stats_in,
                                                      (const float*)gamma, (const float*)beta, (const __nv_bfloat16*)addend, HW, C, G, ppb,
                                                      eps, (float*)red, (unsigned*)red + (size_t)N * G * 2, (float*)dgamma, (float*)dbeta,
                                                      (__nv_bfloat16*)dx, (float*)colsum);
    SH_CHECK_LAUNCH("gn_relu_bwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_gn_relu_bwd(const void* da, const void* x, const void* stats_in, const void* gamma, const void* beta,
                              const void* addend, int N, int HW, int C, int G, float eps, void* red, void* dgamma, void* dbeta,
                              void* dx, void* colsum, void* stream) {
    return gn_relu_bwd_impl(da, x, stats_in, gamma, beta, addend, N, HW, C, G, eps, red, dgamma, dbeta, dx, colsum, stream, true);
}
// Same, for a `red` the caller has already cleared (e.g. one arena for all layers of a backward pass, zeroed once)
SH_EXPORT int sh_gn_relu_bwd_prezeroed(const void* da, const void* x, const void* stats_in, const void* gamma, const void* beta,
                                        const void* addend, int N, int HW, int C, int G, float eps, void* red, void* dgamma,
                                        void* dbeta, void* dx, void* colsum, void* stream) {
    return gn_relu_bwd_impl(da, x, stats_in, gamma, beta, addend, N, HW, C, G, eps, red, dgamma, dbeta, dx, colsum, stream, false);
}

SH_EXPORT int sh_maxpool_fwd(const void* x, int N, int H, int W, int C, void* y, void* stats_out, int G_out, void* stream) {
    SH_REQUIRE(x && y, "sh_maxpool_fwd: null pointer");
    NHWC_CHECK("sh_maxpool_fwd");
    if (N == 0) return SH_OK;
    const int ppb = pick_ppb(H * W, N, kT / (C / 8));
    dim3 grid(sh_div_up(H * W, ppb), N);
    maxpool_fwd_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, H, W, C, ppb, (__nv_bfloat16*)y,
                                                              (float*)stats_out, G_out);
    SH_CHECK_LAUNCH("maxpool_fwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_maxpool_bwd(const void* dy, const void* x, const void* addend, int N, int H, int W, int C, void* dx,
                              void* colsum, void* stream) {
    SH_REQUIRE(dy && x && dx, "sh_maxpool_bwd: null pointer");
    NHWC_CHECK("sh_maxpool_bwd");
    if (N == 0) return SH_OK;
    const int ppb = pick_ppb(H * W, N, kT / (C / 8));
    dim3 grid(sh_div_up(H * W, ppb), N);
    maxpool_bwd_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                              (const __nv_bfloat16*)addend, H, W, C, ppb, (__nv_bfloat16*)dx,
                                                              (float*)colsum);
    SH_CHECK_LAUNCH("maxpool_bwd_kernel");
    return SH_OK;
}

// y[N,2h,2w,C] = up1 + bilinear_x2(low[N,h,w,C])
SH_EXPORT int sh_upsample_add_fwd(const void* up1, const void* low, int N, int h, int w, int C, void* y, void* stats_out,
                                   int G_out, void* stream) {
    SH_REQUIRE(up1 && low && y, "sh_upsample_add_fwd: null pointer");
    NHWC_CHECK("sh_upsample_add_fwd");
    if (N == 0) return SH_OK;
    const int ppb = pick_ppb(4 * h * w, N, kT / (C / 8));
    dim3 grid(sh_div_up(4 * h * w, ppb), N);
    upsample_add_fwd_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)up1, (const __nv_bfloat16*)low, h, w, C, ppb,
                                                                   (__nv_bfloat16*)y, (float*)stats_out, G_out);
    SH_CHECK_LAUNCH("upsample_add_fwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_upsample_bwd(const void* dy, int N, int h, int w, int C, void* dlow, void* colsum, void* stream) {
    SH_REQUIRE(dy && dlow, "sh_upsample_bwd: null pointer");
    NHWC_CHECK("sh_upsample_bwd");
    if (N == 0) return SH_OK;
    const int ppb = pick_ppb(h * w, N, kT / (C / 8));
    dim3 grid(sh_div_up(h * w, ppb), N);
    upsample_bwd_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, h, w, C, ppb, (__nv_bfloat16*)dlow,
                                                               (float*)colsum);
    SH_CHECK_LAUNCH("upsample_bwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_add(const void* a, const void* b, const void* c, int N, int HW, int C, void* y, void* stats_out, int G_out,
                      void* colsum, void* stream) {
    SH_REQUIRE(a && b && y, "sh_add: null pointer");
    NHWC_CHECK("sh_add");
    if (N == 0) return SH_OK;
    const int ppb = pick_ppb(HW, N, kT / (C / 8));
    dim3 grid(sh_div_up(HW, ppb), N);
    add_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)c, HW, C,
                                                      ppb, (__nv_bfloat16*)y, (float*)stats_out, G_out, (float*)colsum);
    SH_CHECK_LAUNCH("add_kernel");
    return SH_OK;
}

SH_EXPORT int sh_colsum(const void* x, int N, int HW, int C, void* colsum, void* stream) {
    SH_REQUIRE(x && colsum, "sh_colsum: null pointer");
    NHWC_CHECK("sh_colsum");
    if (N == 0) return SH_OK;
    // a pure reduction: every block ends in C same-address reductions, so as few blocks as still fill the machine (one wave of 6 per SM)
    const int rows = kT / (C / 8);
    int ppb = HW;
    while (ppb > rows * 16 && (long)N * ((HW + ppb - 1) / ppb) < 6L * SH_NUM_SMS) ppb = (ppb + 1) / 2;
    dim3 grid(sh_div_up(HW, ppb), N);
    colsum_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, HW, C, ppb, (float*)colsum);
    SH_CHECK_LAUNCH("colsum_kernel");
    return SH_OK;
}

SH_EXPORT int sh_stem_conv_fwd(const void* img, const void* w, const void* b, int N, int S, void* y, void* stats_out, int G_out,
                                void* stream) {
    SH_REQUIRE(img && w && b && y, "sh_stem_conv_fwd: null pointer");
    SH_REQUIRE(S % 2 == 0 && S >= 2, "sh_stem_conv_fwd: S must be even");
    SH_REQUIRE(!stats_out || G_out == 1 || G_out == 2 || G_out == 4, "sh_stem_conv_fwd: G_out must be 1, 2 or 4");
    SH_REQUIRE(N <= 65535, "sh_stem_conv_fwd: N > 65535");
    if (N == 0) return SH_OK;
    const int O = S / 2;
    dim3 grid(sh_div_up(O, kStemCols), sh_div_up(O, kStemRows * kStemUnits), N);
    stem_conv_fwd_kernel<<<grid, kT, 0, (cudaStream_t)stream>>>((const float*)img, (const float*)w, (const float*)b, S,
                                                                (__nv_bfloat16*)y, (float*)stats_out, G_out);
    SH_CHECK_LAUNCH("stem_conv_fwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_stem_conv_wgrad(const void* img, const void* dy, int N, int S, void* dw, void* db, void* stream) {
    SH_REQUIRE(img && dy && dw && db, "sh_stem_conv_wgrad: null pointer");
    SH_REQUIRE(S >= 2 && (S & (S - 1)) == 0, "sh_stem_conv_wgrad: S must be a power of two");
    if (N == 0) return SH_OK;
    const int O = S / 2;
    const long units = (long)N * ((O + kSWRows - 1) / kSWRows) * ((O + kStemCols - 1) / kStemCols);
    const int grid = (int)(units < 2L * SH_NUM_SMS ? units : 2L * SH_NUM_SMS);
    stem_conv_wgrad_kernel<<<grid, kStemWT, 0, (cudaStream_t)stream>>>((const float*)img, (const __nv_bfloat16*)dy, N, S, (float*)dw,
                                                                       (float*)db);
    SH_CHECK_LAUNCH("stem_conv_wgrad_kernel");
    return SH_OK;
}

SH_EXPORT int sh_nchw_to_nhwc(const void* x, int N, int C, int HW, int Cp, void* y, void* stream) {
    SH_REQUIRE(x && y && Cp >= C, "sh_nchw_to_nhwc: bad arguments");
    if (N == 0) return SH_OK;
    dim3 grid(sh_div_up(HW, 32), sh_div_up(Cp, 32), N), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float*)x, N, C, HW, Cp, (__nv_bfloat16*)y);
    SH_CHECK_LAUNCH("nchw_to_nhwc_kernel");
    return SH_OK;
}

SH_EXPORT int sh_nhwc_to_nchw(const void* x, int N, int C, int HW, void* y, void* stream) {
    SH_REQUIRE(x && y, "sh_nhwc_to_nchw: bad arguments");
    if (N == 0) return SH_OK;
    dim3 grid(sh_div_up(HW, 32), sh_div_up(C, 32), N), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, N, C, HW, (float*)y);
    SH_CHECK_LAUNCH("nhwc_to_nchw_kernel");
    return SH_OK;
}

SH_EXPORT int sh_pack_weights(const void* w, int Cout, int Cin, int taps, int cout_pad, int cin_pad, int b_rows, int b_cols,
                               void* wf, void* wb, void* stream) {
    SH_REQUIRE(w && wf, "sh_pack_weights: null pointer");
    long n = (long)taps * cout_pad * cin_pad;
    if (wb && (long)taps * b_rows * b_cols > n) n = (long)taps * b_rows * b_cols;
    pack_weights_kernel<<<sh_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)w, Cout, Cin, taps, cout_pad, cin_pad,
                                                                            b_rows, b_cols, (__nv_bfloat16*)wf, (__nv_bfloat16*)wb);
    SH_CHECK_LAUNCH("pack_weights_kernel");
    return SH_OK;
}

SH_EXPORT int sh_pack_weights_batch(const void* flat, const void* table, int n, void* arena, void* stream) {
    SH_REQUIRE(flat && table && arena && n >= 0, "sh_pack_weights_batch: bad arguments");
    if (n == 0) return SH_OK;
    pack_weights_batch_kernel<<<dim3(24, n), 256, 0, (cudaStream_t)stream>>>((const float*)flat, (const int*)table,
                                                                            (__nv_bfloat16*)arena);
    SH_CHECK_LAUNCH("pack_weights_batch_kernel");
    return SH_OK;
}

SH_EXPORT int sh_unpack_wgrad(const void* dw, int Cout, int Cin, int taps, int cout_ld, int cin_ld, void* grad, void* stream) {
    SH_REQUIRE(dw && grad, "sh_unpack_wgrad: null pointer");
    const long n = (long)Cout * Cin * taps;
    unpack_wgrad_kernel<<<sh_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)dw, Cout, Cin, taps, cout_ld, cin_ld,
                                                                            (float*)grad);
    SH_CHECK_LAUNCH("unpack_wgrad_kernel");
    return SH_OK;
}

SH_EXPORT int sh_adam_step(void* p, const void* g, void* m, void* v, long n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, int step, float grad_scale, void* stream) {
    SH_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "sh_adam_step: bad arguments");
    if (n == 0) return SH_OK;
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adam_kernel<<<sh_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>((float*)p, (const float*)g, (float*)m, (float*)v, n, lr, beta1,
                                                                    beta2, eps, weight_decay, bc1, bc2, grad_scale);
    SH_CHECK_LAUNCH("adam_kernel");
    return SH_OK;
}
