// Heat-map heads: soft-argmax to 3-D joints fused with the heat-map MSE terms (sm_100a).
//
// Replaces
//   SpatialSoftmax / SpatialNormalization / RecoverXYZCoordinateFromHeatmap.forward
//       /root/reference/network/util_modules.py:126-141, 144-161, 182-201
//   the channel split of HeatmapEstimationNetwork.forward (uv = [:J], d = [J:])
//       /root/reference/network/create_network_and_criterion.py:111-123
//   the heat-map MSE terms of MultiTaskLoss.forward (synt_uv vs target, real uv vs 0)
//       /root/reference/network/create_network_and_criterion.py:188-193, 231-235
// and their autograd backward.  One warp per (sample, joint); each heat-map is read once per pass.
//
// HBM layout: score fp32 [N, C, h, w] (NCHW; uv channels [0,J), depth channels [J,2J)); xyz fp32 [N,J,3] mm;
// target_uv fp32 [Ns,J,h,w] (only for samples n < Ns, may be null); sse double[2] = (sum (uv-target)^2 over
// synthetic samples, sum uv^2 over real samples).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128) softargmax_fwd_kernel(const float* __restrict__ score, int N, int Ns, int J,
                                                             int C, int h, int w, float depth_scale_inv,
                                                             const float* __restrict__ target_uv,
                                                             float* __restrict__ xyz, double* __restrict__ sse, float* __restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= N * J) return;
    const int n = item / J, k = item - n * J;
    const int hw = h * w;
    const float* uv = score + ((size_t)n * C + k) * hw;
    const float* dh = score + ((size_t)n * C + J + k) * hw;
    const float* tg = (target_uv && n < Ns) ? target_uv + ((size_t)n * J + k) * hw : nullptr;
    float mx = -3.4e38f;
    for (int i = lane; i < hw; i += 32) mx = fmaxf(mx, uv[i] * 20.f);
    mx = warp_max(mx);
    float S = 0.f, su = 0.f, sv = 0.f, R = 0.f, dn = 0.f, se = 0.f;
    for (int i = lane; i < hw; i += 32) {
        const float x = uv[i];
        const float e = __expf(x * 20.f - mx);
        const int vv = i / w, uu = i - vv * w;
        S += e;
        su += e * (float)uu;
        sv += e * (float)vv;
        const float rl = fmaxf(x, 0.f);
        R += rl;
        dn += dh[i] * rl;
        const float df = tg ? x - tg[i] : x;
        se += df * df;
    }
    S = warp_sum(S); su = warp_sum(su); sv = warp_sum(sv); R = warp_sum(R); dn = warp_sum(dn); se = warp_sum(se);
    if (lane == 0) {
        const float u = su / S, v = sv / S, d = dn / (R + 1e-5f);
        float* o = xyz + (size_t)item * 3;
        o[0] = (u - w * 0.5f) / (w / 300.0f);
        o[1] = (v - h * 0.5f) / (h / 300.0f);
        o[2] = d * depth_scale_inv;
        if (sse) atomicAdd(&sse[n < Ns ? 0 : 1], (double)se);
        if (aux) {          // what the backward needs per (sample, joint): softmax max / 1/sum, the soft-argmax, 1/(sum relu + 1e-5), depth
            float* a = aux + (size_t)item * 8;
            *reinterpret_cast<float4*>(a) = make_float4(mx, 1.f / S, u, v);
            *reinterpret_cast<float4*>(a + 4) = make_float4(1.f / (R + 1e-5f), d, 0.f, 0.f);
        }
    }
}

// The same backward as softargmax_bwd_kernel, PIXEL-major and written straight into the layout the network's backward pass
// consumes: bf16 NHWC [N,h,w,Cp] (channels [0,2J) = d loss / d score, the padding channels zero), from the per-(sample, joint)
// scalars the forward left in `aux`.  It replaces softargmax_bwd_kernel + nchw_to_nhwc_kernel (an fp32 NCHW round trip of the
// score gradient: 86 MB written and read again per stack at the BASELINE shape).  A thread owns one pixel: the loads of one
// channel are coalesced across the warp, the thread's 2J values are packed in registers and leave as 16-byte stores.
template <int J, int CP>
__global__ void __launch_bounds__(128) softargmax_bwd_nhwc_kernel(const float* __restrict__ score, const float* __restrict__ gxyz,
                                                                  const float* __restrict__ aux, int Ns, int C, int h, int w,
                                                                  float depth_scale_inv, const float* __restrict__ target_uv,
                                                                  float c_synt, float c_real, __nv_bfloat16* __restrict__ dscore) {
    __shared__ __align__(16) float s_aux[J * 8];       // mx, 1/S, u, v, 1/(R+eps), d, -, -
    __shared__ float s_g[J * 3];                       // gu * 20 * 300/w, gv * 20 * 300/h, gz * depth_scale_inv
    const int n = blockIdx.y;
    const int hw = h * w;
    for (int i = threadIdx.x; i < J * 8; i += 128) s_aux[i] = aux[(size_t)n * J * 8 + i];
    for (int i = threadIdx.x; i < J * 3; i += 128) {
        const float g = gxyz[(size_t)n * J * 3 + i];
        const int c = i % 3;
        s_g[i] = c == 0 ? g * (300.0f / w) * 20.f : (c == 1 ? g * (300.0f / h) * 20.f : g * depth_scale_inv);
    }
    __syncthreads();
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= hw) return;
    const int vv = p / w, uu = p - vv * w;
    const float fu = (float)uu, fv = (float)vv;
    const bool synt = n < Ns;
    const float cm = synt ? c_synt : c_real;
    const float* sc = score + (size_t)n * C * hw + p;
    const float* tg = (target_uv && synt) ? target_uv + (size_t)n * J * hw + p : nullptr;
    float out[2 * J];
#pragma unroll
    for (int k = 0; k < J; ++k) {
        const float x = sc[(size_t)k * hw];
        const float dhv = sc[(size_t)(J + k) * hw];
        const float4 a0 = *reinterpret_cast<const float4*>(s_aux + k * 8);
        const float rinv = s_aux[k * 8 + 4], d = s_aux[k * 8 + 5];
        const float gz = s_g[k * 3 + 2];
        const float pr = __expf(x * 20.f - a0.x) * a0.y;
        float gx = pr * (s_g[k * 3] * (fu - a0.z) + s_g[k * 3 + 1] * (fv - a0.w));
        if (x > 0.f) gx += gz * (dhv - d) * rinv;
        gx += cm * (tg ? x - tg[(size_t)k * hw] : x);
        out[k] = gx;
        out[J + k] = gz * fmaxf(x, 0.f) * rinv;
    }
    uint32_t pk[CP / 2];
#pragma unroll
    for (int c = 0; c < CP / 2; ++c) {
        const float lo = 2 * c < 2 * J ? out[2 * c < 2 * J ? 2 * c : 0] : 0.f;
        const float hi = 2 * c + 1 < 2 * J ? out[2 * c + 1 < 2 * J ? 2 * c + 1 : 0] : 0.f;
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(lo, hi);
        pk[c] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    uint4* dst = reinterpret_cast<uint4*>(dscore + ((size_t)n * hw + p) * CP);
#pragma unroll
    for (int q = 0; q < CP / 8; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// gscore[uv] = soft-argmax backward + c_synt*(uv - target) (n < Ns) or c_real*uv (n >= Ns); gscore[d] likewise.
__global__ void __launch_bounds__(128) softargmax_bwd_kernel(const float* __restrict__ score, const float* __restrict__ gxyz,
                                                             int N, int Ns, int J, int C, int h, int w,
                                                             float depth_scale_inv, const float* __restrict__ target_uv,
                                                             float c_synt, float c_real, float* __restrict__ gscore) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= N * J) return;
    const int n = item / J, k = item - n * J;
    const int hw = h * w;
    const size_t ouv = ((size_t)n * C + k) * hw, od = ((size_t)n * C + J + k) * hw;
    const float* uv = score + ouv;
    const float* dh = score + od;
    const float* tg = (target_uv && n < Ns) ? target_uv + ((size_t)n * J + k) * hw : nullptr;
    const float cm = n < Ns ? c_synt : c_real;
    float mx = -3.4e38f;
    for (int i = lane; i < hw; i += 32) mx = fmaxf(mx, uv[i] * 20.f);
    mx = warp_max(mx);
    float S = 0.f, su = 0.f, sv = 0.f, R = 0.f, dn = 0.f;
    for (int i = lane; i < hw; i += 32) {
        const float x = uv[i];
        const float e = __expf(x * 20.f - mx);
        const int vv = i / w, uu = i - vv * w;
        S += e; su += e * (float)uu; sv += e * (float)vv;
        const float rl = fmaxf(x, 0.f);
        R += rl; dn += dh[i] * rl;
    }
    S = warp_sum(S); su = warp_sum(su); sv = warp_sum(sv); R = warp_sum(R); dn = warp_sum(dn);
    const float u = su / S, v = sv / S, rinv = 1.f / (R + 1e-5f), d = dn * rinv;
    const float* g = gxyz + (size_t)item * 3;
    const float gu = g[0] * (300.0f / w) * 20.f, gv = g[1] * (300.0f / h) * 20.f, gz = g[2] * depth_scale_inv;
    const float sinv = 1.f / S;
    for (int i = lane; i < hw; i += 32) {
        const float x = uv[i];
        const float p = __expf(x * 20.f - mx) * sinv;
        const int vv = i / w, uu = i - vv * w;
        float gx = p * (gu * ((float)uu - u) + gv * ((float)vv - v));
        if (x > 0.f) gx += gz * (dh[i] - d) * rinv;
        gx += cm * (tg ? x - tg[i] : x);
        gscore[ouv + i] = gx;
        gscore[od + i] = gz * fmaxf(x, 0.f) * rinv;
    }
}

}  // namespace

static int softargmax_fwd_impl(const void* score, int N, int Ns, int J, int C, int h, int w, float depth_scale_inv,
                               const void* target_uv, void* xyz, void* sse2, void* aux, void* stream) {
    SH_REQUIRE(score && xyz, "sh_softargmax_fwd: null pointer");
    SH_REQUIRE(N >= 0 && J >= 1 && C >= 2 * J && h >= 1 && w >= 1 && Ns >= 0 && Ns <= N, "sh_softargmax_fwd: bad shape");
    if (N == 0) return SH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (sse2) SH_CUDA(cudaMemsetAsync(sse2, 0, 16, st));
    softargmax_fwd_kernel<<<sh_div_up((long)N * J, 4), 128, 0, st>>>((const float*)score, N, Ns, J, C, h, w,
                                                                     depth_scale_inv, (const float*)target_uv,
                                                                     (float*)xyz, (double*)sse2, (float*)aux);
    SH_CHECK_LAUNCH("softargmax_fwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_softargmax_fwd(const void* score, int N, int Ns, int J, int C, int h, int w, float depth_scale_inv,
                                 const void* target_uv, void* xyz, void* sse2, void* stream) {
    return softargmax_fwd_impl(score, N, Ns, J, C, h, w, depth_scale_inv, target_uv, xyz, sse2, nullptr, stream);
}

SH_EXPORT int sh_softargmax_fwd_aux(const void* score, int N, int Ns, int J, int C, int h, int w, float depth_scale_inv,
                                     const void* target_uv, void* xyz, void* sse2, void* aux, void* stream) {
    SH_REQUIRE(aux && ((uintptr_t)aux & 15) == 0, "sh_softargmax_fwd_aux: aux must be a 16-byte aligned fp32 [N,J,8] buffer");
    return softargmax_fwd_impl(score, N, Ns, J, C, h, w, depth_scale_inv, target_uv, xyz, sse2, aux, stream);
}

SH_EXPORT int sh_softargmax_bwd_nhwc(const void* score, const void* gxyz, const void* aux, int N, int Ns, int J, int C, int h, int w,
                                      float depth_scale_inv, const void* target_uv, float c_synt, float c_real, void* dscore,
                                      int Cp, void* stream) {
    SH_REQUIRE(score && gxyz && aux && dscore, "sh_softargmax_bwd_nhwc: null pointer");
    SH_REQUIRE(N >= 0 && C >= 2 * J && h >= 1 && w >= 1 && Ns >= 0 && Ns <= N, "sh_softargmax_bwd_nhwc: bad shape");
    SH_REQUIRE(J == 41 && Cp == 128, "sh_softargmax_bwd_nhwc: built for J = 41 joints and 128 padded channels (got J=%d Cp=%d); "
               "use sh_softargmax_bwd + sh_nchw_to_nhwc otherwise", J, Cp);
    SH_REQUIRE(((uintptr_t)dscore & 15) == 0 && ((uintptr_t)aux & 15) == 0, "sh_softargmax_bwd_nhwc: dscore / aux must be 16-byte aligned");
    if (N == 0) return SH_OK;
    dim3 grid(sh_div_up((long)h * w, 128), N);
    softargmax_bwd_nhwc_kernel<41, 128><<<grid, 128, 0, (cudaStream_t)stream>>>(
        (const float*)score, (const float*)gxyz, (const float*)aux, Ns, C, h, w, depth_scale_inv, (const float*)target_uv, c_synt,
        c_real, (__nv_bfloat16*)dscore);
    SH_CHECK_LAUNCH("softargmax_bwd_nhwc_kernel");
    return SH_OK;
}

SH_EXPORT int sh_softargmax_bwd(const void* score, const void* gxyz, int N, int Ns, int J, int C, int h, int w,
                                 float depth_scale_inv, const void* target_uv, float c_synt, float c_real,
                                 void* gscore, void* stream) {
    SH_REQUIRE(score && gxyz && gscore, "sh_softargmax_bwd: null pointer");
    SH_REQUIRE(N >= 0 && J >= 1 && C >= 2 * J && h >= 1 && w >= 1 && Ns >= 0 && Ns <= N, "sh_softargmax_bwd: bad shape");
    if (N == 0) return SH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    softargmax_bwd_kernel<<<sh_div_up((long)N * J, 4), 128, 0, st>>>((const float*)score, (const float*)gxyz, N, Ns, J,
                                                                     C, h, w, depth_scale_inv, (const float*)target_uv,
                                                                     c_synt, c_real, (float*)gscore);
    SH_CHECK_LAUNCH("softargmax_bwd_kernel");
    return SH_OK;
}
