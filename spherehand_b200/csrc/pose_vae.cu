// PoseVae.prior_loss — frozen VAE pose prior, forward value + gradient w.r.t. the pose in one launch (sm_100a).
//
// Replaces  PoseVae.prior_loss / _reparameterize / _likelihood
//     /root/reference/network/pose_vae.py:25-62, 81-89
// (123 -> 256 -> 256 -> (32,32) -> 256 -> 256 -> 123 MLP with GroupNorm(16,256)+ReLU, z = mu + eps*0.1*exp(logvar/2),
//  loss = MSE(x, recon) + (-1/2) sum(1 + logvar - mu^2 - e^logvar)).  The weights are frozen (pose_vae.py:22-23), so
// only d(loss)/dx is produced.  The reference runs 7 small GEMMs + 4 GroupNorms + ~20 elementwise kernels and their
// backward (launch bound at M = B*V = 192 rows); here each CTA owns 4 rows and keeps every activation in shared
// memory.  fp32 throughout (the 1e-4 tolerance of the loss heads rules out bf16 tensor cores for a 0.2 MFLOP/row MLP).
//
// HBM layout: x fp32 [M,123]; eps fp32 [M,32] (host-drawn N(0,1)); weights: one packed fp32 blob,
// every matrix stored twice (W [out,in] for the backward, W^T [in,out] for the forward) — see VaeWeights.
#include "common.cuh"

namespace {

constexpr int RB = 4;        // rows per CTA
constexpr int HID = 256;
constexpr int POSE = 123;
constexpr int LAT = 32;
constexpr int kThreads = 256;

struct VaeWeights {          // pointers into the packed blob
    const float *w1, *w1t, *b1, *g1, *be1;     // 123 -> 256
    const float *w2, *w2t, *b2, *g2, *be2;     // 256 -> 256
    const float *wm, *wmt, *bm;                // 256 -> 32
    const float *wl, *wlt, *bl;                // 256 -> 32
    const float *w3, *w3t, *b3, *g3, *be3;     // 32 -> 256
    const float *w4, *w4t, *b4, *g4, *be4;     // 256 -> 256
    const float *w5, *w5t, *b5;                // 256 -> 123
};

// out[r][t] = bias[t] + sum_k in[r][k] * Wt[k][t]   (thread t, coalesced over t)
__device__ __forceinline__ void dense_fwd(const float* __restrict__ wt, const float* __restrict__ bias, int n_in, int n_out,
                                          const float* in, int in_stride, float* out, int out_stride) {
    const int t = threadIdx.x;
    if (t < n_out) {
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = bias[t];
#pragma unroll 32                // 32 independent weight loads in flight: the loop is a serial L2-latency chain otherwise
        for (int k = 0; k < n_in; ++k) {
            const float w = __ldg(wt + (size_t)k * n_out + t);
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[r] += in[r * in_stride + k] * w;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) out[r * out_stride + t] = acc[r];
    }
    __syncthreads();
}

// din[r][k] (+)= sum_t dout[r][t] * W[t][k]   (thread k, coalesced over k)
__device__ __forceinline__ void dense_bwd(const float* __restrict__ w, int n_in, int n_out, const float* dout,
                                          int dout_stride, float* din, int din_stride, bool accumulate) {
    const int k = threadIdx.x;
    if (k < n_in) {
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.f;
#pragma unroll 32
        for (int t = 0; t < n_out; ++t) {
            const float wv = __ldg(w + (size_t)t * n_in + k);
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[r] += dout[r * dout_stride + t] * wv;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (accumulate) din[r * din_stride + k] += acc[r];
            else din[r * din_stride + k] = acc[r];
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// GroupNorm(16 groups of 16 channels) + ReLU over HID features, in place: a -> h; xhat, rstd saved for the backward
__device__ __forceinline__ void gn_relu_fwd(const float* __restrict__ gamma, const float* __restrict__ beta, const float* a,
                                            float* h, float* xhat, float* rstd) {
    const int t = threadIdx.x;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const float v = a[r * HID + t];
        const float mean = group16_sum(v) * (1.f / 16.f);
        const float d = v - mean;
        const float var = group16_sum(d * d) * (1.f / 16.f);
        const float rs = rsqrtf(var + 1e-5f);
        const float xh = d * rs;
        xhat[r * HID + t] = xh;
        if ((t & 15) == 0) rstd[r * 16 + (t >> 4)] = rs;
        h[r * HID + t] = fmaxf(xh * gamma[t] + beta[t], 0.f);
    }
    __syncthreads();
}

// dh (gradient w.r.t. the ReLU output h) -> da (gradient w.r.t. the pre-norm activation), in place in `d`
__device__ __forceinline__ void gn_relu_bwd(const float* __restrict__ gamma, const float* h, const float* xhat,
                                            const float* rstd, float* d) {
    const int t = threadIdx.x;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const float dy = h[r * HID + t] > 0.f ? d[r * HID + t] : 0.f;
        const float dxh = dy * gamma[t];
        const float xh = xhat[r * HID + t];
        const float m1 = group16_sum(dxh) * (1.f / 16.f);
        const float m2 = group16_sum(dxh * xh) * (1.f / 16.f);
        d[r * HID + t] = rstd[r * 16 + (t >> 4)] * (dxh - m1 - xh * m2);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) vae_prior_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                                             const VaeWeights W, int M, int M_mean, double* __restrict__ acc,
                                                             float* __restrict__ grad_x) {
    extern __shared__ float sm[];
    float* s_x = sm;                       // [RB][128]
    float* s_a = s_x + RB * 128;           // scratch pre-norm [RB][HID]
    float* s_h[4];
    float* s_xh[4];
    float* p = s_a + RB * HID;
    for (int i = 0; i < 4; ++i) { s_h[i] = p; p += RB * HID; s_xh[i] = p; p += RB * HID; }
    float* s_rs = p; p += 4 * RB * 16;     // rstd per layer
    float* s_mu = p; p += RB * LAT;
    float* s_lv = p; p += RB * LAT;
    float* s_z = p; p += RB * LAT;
    float* s_rec = p; p += RB * 128;
    float* s_d = p; p += RB * HID;         // gradient buffer A
    float* s_d2 = p; p += RB * HID;        // gradient buffer B
    float* s_red = p;                      // [2]

    const int t = threadIdx.x;
    const int row0 = blockIdx.x * RB;
    for (int i = t; i < RB * 128; i += kThreads) {
        const int r = i / 128, k = i % 128;
        s_x[i] = (row0 + r < M && k < POSE) ? x[(size_t)(row0 + r) * POSE + k] : 0.f;
    }
    if (t < 2) s_red[t] = 0.f;
    __syncthreads();

    // ---------------- forward
    dense_fwd(W.w1t, W.b1, POSE, HID, s_x, 128, s_a, HID);
    gn_relu_fwd(W.g1, W.be1, s_a, s_h[0], s_xh[0], s_rs + 0 * RB * 16);
    dense_fwd(W.w2t, W.b2, HID, HID, s_h[0], HID, s_a, HID);
    gn_relu_fwd(W.g2, W.be2, s_a, s_h[1], s_xh[1], s_rs + 1 * RB * 16);
    dense_fwd(W.wmt, W.bm, HID, LAT, s_h[1], HID, s_mu, LAT);
    dense_fwd(W.wlt, W.bl, HID, LAT, s_h[1], HID, s_lv, LAT);
    float kld = 0.f;
    if (t < RB * LAT) {
        const int r = t / LAT, k = t % LAT;
        const float mu = s_mu[t], lv = s_lv[t];
        const float e = (row0 + r < M) ? eps[(size_t)(row0 + r) * LAT + k] : 0.f;
        s_z[t] = e * (expf(0.5f * lv) * 0.1f) + mu;                         // pose_vae.py:50-52
        if (row0 + r < M) kld = -0.5f * (1.f + lv - mu * mu - expf(lv));    // pose_vae.py:61
    }
    __syncthreads();
    dense_fwd(W.w3t, W.b3, LAT, HID, s_z, LAT, s_a, HID);
    gn_relu_fwd(W.g3, W.be3, s_a, s_h[2], s_xh[2], s_rs + 2 * RB * 16);
    dense_fwd(W.w4t, W.b4, HID, HID, s_h[2], HID, s_a, HID);
    gn_relu_fwd(W.g4, W.be4, s_a, s_h[3], s_xh[3], s_rs + 3 * RB * 16);
    dense_fwd(W.w5t, W.b5, HID, POSE, s_h[3], HID, s_rec, 128);

    // ---------------- loss + seed gradient  (mean over M_mean*123 elements: M_mean = rows of the GLOBAL batch)
    const float n_inv = 1.f / ((float)M_mean * POSE);
    float sse = 0.f;
    for (int i = t; i < RB * 128; i += kThreads) {
        const int r = i / 128, k = i % 128;
        float g = 0.f;
        if (row0 + r < M && k < POSE) {
            const float df = s_x[i] - s_rec[i];
            sse += df * df;
            g = 2.f * df * n_inv;
        }
        s_rec[i] = -g;      // d loss / d recon ; the direct d/dx term (+g) is added at the end
    }
    sse = warp_sum(sse);
    kld = warp_sum(kld);
    if ((t & 31) == 0) { atomicAdd(&s_red[0], sse); atomicAdd(&s_red[1], kld); }
    __syncthreads();

    // ---------------- backward
    dense_bwd(W.w5, HID, POSE, s_rec, 128, s_d, HID, false);
    gn_relu_bwd(W.g4, s_h[3], s_xh[3], s_rs + 3 * RB * 16, s_d);
    dense_bwd(W.w4, HID, HID, s_d, HID, s_d2, HID, false);
    gn_relu_bwd(W.g3, s_h[2], s_xh[2], s_rs + 2 * RB * 16, s_d2);
    dense_bwd(W.w3, LAT, HID, s_d2, HID, s_z, LAT, false);                 // s_z now holds dL/dz
    if (t < RB * LAT) {
        const int r = t / LAT, k = t % LAT;
        const float dz = s_z[t], mu = s_mu[t], lv = s_lv[t];
        const float e = (row0 + r < M) ? eps[(size_t)(row0 + r) * LAT + k] : 0.f;
        const bool live = row0 + r < M;
        // z = mu + eps*0.1*exp(lv/2);  KLD = -1/2 (1 + lv - mu^2 - e^lv)
        s_mu[t] = dz + (live ? mu : 0.f);
        s_lv[t] = dz * e * 0.05f * expf(0.5f * lv) + (live ? -0.5f * (1.f - expf(lv)) : 0.f);
    }
    __syncthreads();
    dense_bwd(W.wm, HID, LAT, s_mu, LAT, s_d, HID, false);
    dense_bwd(W.wl, HID, LAT, s_lv, LAT, s_d, HID, true);
    gn_relu_bwd(W.g2, s_h[1], s_xh[1], s_rs + 1 * RB * 16, s_d);
    dense_bwd(W.w2, HID, HID, s_d, HID, s_d2, HID, false);
    gn_relu_bwd(W.g1, s_h[0], s_xh[0], s_rs + 0 * RB * 16, s_d2);
    dense_bwd(W.w1, POSE, HID, s_d2, HID, s_a, 128, false);               // s_a[r*128+k] = dL/dx via the encoder
    for (int i = t; i < RB * 128; i += kThreads) {
        const int r = i / 128, k = i % 128;
        if (row0 + r < M && k < POSE) grad_x[(size_t)(row0 + r) * POSE + k] = s_a[i] - s_rec[i];   // + direct term
    }
    if (t < 2) atomicAdd(&acc[t], (double)s_red[t]);
}

__global__ void vae_finish_kernel(const double* __restrict__ acc, int M, float* __restrict__ loss3) {
    if (threadIdx.x == 0) {
        const double recon = acc[0] / ((double)M * POSE);
        loss3[0] = (float)(recon + acc[1]);
        loss3[1] = (float)recon;
        loss3[2] = (float)acc[1];
    }
}

constexpr size_t kSmemFloats = RB * 128 + RB * HID + 8 * RB * HID + 4 * RB * 16 + 3 * RB * LAT + RB * 128 + 2 * RB * HID + 8;

}  // namespace

// Packed weight blob layout (floats), built once by the host from the state_dict (pose_vae.py:26-46):
//   for each of the 7 Linear layers in order base.0, base.3, mu, logvar, decoder.0, decoder.3, decoder.6:
//       W [out,in], W^T [in,out], bias [out]; and, for the 4 layers followed by GroupNorm, gamma [256], beta [256].
SH_EXPORT size_t sh_vae_blob_floats(void) {
    size_t n = 0;
    const int dims[7][2] = {{POSE, HID}, {HID, HID}, {HID, LAT}, {HID, LAT}, {LAT, HID}, {HID, HID}, {HID, POSE}};
    const int has_gn[7] = {1, 1, 0, 0, 1, 1, 0};
    for (int i = 0; i < 7; ++i) n += 2 * (size_t)dims[i][0] * dims[i][1] + dims[i][1] + (has_gn[i] ? 2 * HID : 0);
    return n;
}

SH_EXPORT int sh_vae_prior_fwdbwd(const void* x, const void* eps, const void* weight_blob, int M, int M_mean, void* loss3,
                                   void* grad_x, void* scratch, void* stream) {
    SH_REQUIRE(x && eps && weight_blob && loss3 && grad_x && scratch, "sh_vae_prior_fwdbwd: null pointer");
    SH_REQUIRE(M >= 1 && M_mean >= M, "sh_vae_prior_fwdbwd: bad M / M_mean");
    cudaStream_t st = (cudaStream_t)stream;
    const float* p = (const float*)weight_blob;
    VaeWeights W;
    auto take = [&](size_t n) { const float* q = p; p += n; return q; };
    W.w1 = take(HID * POSE); W.w1t = take(HID * POSE); W.b1 = take(HID); W.g1 = take(HID); W.be1 = take(HID);
    W.w2 = take(HID * HID); W.w2t = take(HID * HID); W.b2 = take(HID); W.g2 = take(HID); W.be2 = take(HID);
    W.wm = take(LAT * HID); W.wmt = take(LAT * HID); W.bm = take(LAT);
    W.wl = take(LAT * HID); W.wlt = take(LAT * HID); W.bl = take(LAT);
    W.w3 = take(HID * LAT); W.w3t = take(HID * LAT); W.b3 = take(HID); W.g3 = take(HID); W.be3 = take(HID);
    W.w4 = take(HID * HID); W.w4t = take(HID * HID); W.b4 = take(HID); W.g4 = take(HID); W.be4 = take(HID);
    W.w5 = take(POSE * HID); W.w5t = take(POSE * HID); W.b5 = take(POSE);
    static bool attr_set = false;
    const size_t smem = kSmemFloats * sizeof(float);
    if (!attr_set) {
        SH_CUDA(cudaFuncSetAttribute(vae_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    SH_CUDA(cudaMemsetAsync(scratch, 0, 16, st));
    vae_prior_kernel<<<sh_div_up(M, RB), kThreads, smem, st>>>((const float*)x, (const float*)eps, W, M, M_mean, (double*)scratch,
                                                              (float*)grad_x);
    SH_CHECK_LAUNCH("vae_prior_kernel");
    vae_finish_kernel<<<1, 32, 0, st>>>((const double*)scratch, M_mean, (float*)loss3);
    SH_CHECK_LAUNCH("vae_finish_kernel");
    return SH_OK;
}

// ------------------------------------------------------------------------------------------------ PoseDenoiser (eval)
// PoseDenoiser.forward in eval mode (/root/reference/network/pose_denoiser.py:56-73): gather 112 of the 123 joint coordinates,
// x0.01, Linear(112,256)+GroupNorm(16,256)+ReLU, Linear(256,256)+GroupNorm+ReLU, Linear(256,33), /0.01, scatter into a copy of
// the input.  The reference runs 3 GEMMs, 2 GroupNorms and ~10 index / elementwise kernels for a handful of rows; here a CTA
// owns 4 rows and keeps every activation in shared memory (same dense / GroupNorm helpers as the VAE prior above).
namespace {

struct DenWeights {
    const float *w1t, *b1, *g1, *be1;      // n_in -> 256  (W^T [in,out])
    const float *w2t, *b2, *g2, *be2;      // 256  -> 256
    const float *w3t, *b3;                 // 256  -> n_out
};

__global__ void __launch_bounds__(kThreads) pose_denoiser_kernel(const float* __restrict__ fea, const int* __restrict__ in_idx,
                                                                 const int* __restrict__ out_idx, const DenWeights W, int M, int n_fea,
                                                                 int n_in, int n_out, float scale, float* __restrict__ out) {
    __shared__ float s_x[RB * HID];        // gathered, scaled input rows (n_in <= 256)
    __shared__ float s_a[RB * HID];
    __shared__ float s_h[RB * HID];
    __shared__ float s_xh[RB * HID];
    __shared__ float s_rs[RB * 16];
    __shared__ float s_o[RB * HID];
    const int t = threadIdx.x;
    const int row0 = blockIdx.x * RB;
    for (int i = t; i < RB * n_in; i += kThreads) {
        const int r = i / n_in, k = i - r * n_in;
        s_x[r * HID + k] = row0 + r < M ? __fmul_rn(fea[(size_t)(row0 + r) * n_fea + in_idx[k]], scale) : 0.f;
    }
    for (int i = t; i < RB * n_fea; i += kThreads) {                       // denoised_fea = fea.clone()
        const int r = i / n_fea, k = i - r * n_fea;
        if (row0 + r < M) out[(size_t)(row0 + r) * n_fea + k] = fea[(size_t)(row0 + r) * n_fea + k];
    }
    __syncthreads();
    dense_fwd(W.w1t, W.b1, n_in, HID, s_x, HID, s_a, HID);
    gn_relu_fwd(W.g1, W.be1, s_a, s_h, s_xh, s_rs);
    dense_fwd(W.w2t, W.b2, HID, HID, s_h, HID, s_a, HID);
    gn_relu_fwd(W.g2, W.be2, s_a, s_h, s_xh, s_rs);
    dense_fwd(W.w3t, W.b3, HID, n_out, s_h, HID, s_o, HID);
    for (int i = t; i < RB * n_out; i += kThreads) {
        const int r = i / n_out, k = i - r * n_out;
        if (row0 + r < M) out[(size_t)(row0 + r) * n_fea + out_idx[k]] = __fdiv_rn(s_o[r * HID + k], scale);
    }
}

}  // namespace

// Packed blob (floats): W1^T [n_in,256], b1, gamma1, beta1 [256 each], W2^T [256,256], b2, gamma2, beta2, W3^T [256,n_out], b3 [n_out].
SH_EXPORT size_t sh_pose_denoiser_blob_floats(int n_in, int n_out) {
    return (size_t)n_in * HID + 3 * HID + (size_t)HID * HID + 3 * HID + (size_t)HID * n_out + n_out;
}

// fea fp32 [M,n_fea]; in_idx int32 [n_in] (<= 256), out_idx int32 [n_out] (<= 256), entries in [0,n_fea); out fp32 [M,n_fea]
SH_EXPORT int sh_pose_denoiser_fwd(const void* fea, const void* in_idx, const void* out_idx, const void* blob, int M, int n_fea,
                                    int n_in, int n_out, float scale, void* out, void* stream) {
    SH_REQUIRE(M >= 0 && n_fea >= 1 && n_in >= 1 && n_in <= HID && n_out >= 1 && n_out <= HID, "sh_pose_denoiser_fwd: bad sizes");
    if (M == 0) return SH_OK;
    SH_REQUIRE(fea && in_idx && out_idx && blob && out, "sh_pose_denoiser_fwd: null pointer");
    SH_REQUIRE(scale != 0.f, "sh_pose_denoiser_fwd: scale must be non-zero");
    const float* p = (const float*)blob;
    auto take = [&](size_t n) { const float* q = p; p += n; return q; };
    DenWeights W;
    W.w1t = take((size_t)n_in * HID); W.b1 = take(HID); W.g1 = take(HID); W.be1 = take(HID);
    W.w2t = take((size_t)HID * HID); W.b2 = take(HID); W.g2 = take(HID); W.be2 = take(HID);
    W.w3t = take((size_t)HID * n_out); W.b3 = take(n_out);
    pose_denoiser_kernel<<<sh_div_up(M, RB), kThreads, 0, (cudaStream_t)stream>>>((const float*)fea, (const int*)in_idx, (const int*)out_idx, W, M,
                                                                                   n_fea, n_in, n_out, scale, (float*)out);
    SH_CHECK_LAUNCH("pose_denoiser_kernel");
    return SH_OK;
}
