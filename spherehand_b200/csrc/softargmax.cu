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
                                                             float* __restrict__ xyz, double* __restrict__ sse) {
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
    }
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

SH_EXPORT int sh_softargmax_fwd(const void* score, int N, int Ns, int J, int C, int h, int w, float depth_scale_inv,
                                 const void* target_uv, void* xyz, void* sse2, void* stream) {
    SH_REQUIRE(score && xyz, "sh_softargmax_fwd: null pointer");
    SH_REQUIRE(N >= 0 && J >= 1 && C >= 2 * J && h >= 1 && w >= 1 && Ns >= 0 && Ns <= N, "sh_softargmax_fwd: bad shape");
    if (N == 0) return SH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (sse2) SH_CUDA(cudaMemsetAsync(sse2, 0, 16, st));
    softargmax_fwd_kernel<<<sh_div_up((long)N * J, 4), 128, 0, st>>>((const float*)score, N, Ns, J, C, h, w,
                                                                     depth_scale_inv, (const float*)target_uv,
                                                                     (float*)xyz, (double*)sse2);
    SH_CHECK_LAUNCH("softargmax_fwd_kernel");
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
