// Train-step glue (sm_100a): weighting / routing of the loss-head gradients into the soft-argmax backward, the scalar
// loss terms, and a CUDA-graph-replayable Adam step.
//
// Replaces the autograd bookkeeping of
//   MultiTaskLoss.forward (weights + per-term sums)   /root/reference/network/create_network_and_criterion.py:171-181, 183-263
//   Engine.sum_loss_terms                              /root/reference/network/engine.py:144-148
//   torch.optim.Adam(lr, weight_decay=1e-5).step()     /root/reference/network/engine.py:95-97, 376
// All of it is O(N*J) work: one launch each, nothing here is bandwidth relevant.
#include "common.cuh"

namespace {

struct CombineW {
    float synt_hm, synt_pt, proj, cons, hm_mean, prior, col, bone;   // MultiTaskLoss.weights (:171-181)
};

// gxyz[n] for n < Ns (synthetic rows): d/dxyz of w_pt * MSE(z, target_z)              (:194-199)
// gxyz[Ns+m] (real rows): sum_t w_t * d loss_t / d xyz                                (:208-255)
// terms[8] += (synt_uv, synt_d, mv_projection, mv_consistency, uv_hm_mean, pose_prior, collision, bone_length); terms[8] += total
__global__ void __launch_bounds__(256) step_combine_kernel(
    const float* __restrict__ g_mvproj, const float* __restrict__ g_pose3, const float* __restrict__ g_prior,
    const float* __restrict__ xyz, const float4* __restrict__ target_xyz4, const float* __restrict__ loss_mv3,
    const float* __restrict__ loss_pose3, const float* __restrict__ loss_prior3, const double* __restrict__ sse2, int Ns,
    int M, int J, int hw, CombineW w, float mean_scale, float* __restrict__ gxyz, float* __restrict__ terms) {
    __shared__ float s_red[8];
    const int real_elems = M * J * 3, synt_elems = Ns * J * 3;
    const size_t pose_stride = (size_t)real_elems;
    float sq = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < synt_elems + real_elems; i += gridDim.x * blockDim.x) {
        float g = 0.f;
        if (i < synt_elems) {
            if (i % 3 == 2) {
                const float d = xyz[i] - target_xyz4[i / 3].z;
                sq += d * d;
                g = w.synt_pt * mean_scale * 2.f * d / (float)(Ns * J);
            }
        } else {
            const int r = i - synt_elems;
            if (g_mvproj) g += w.proj * mean_scale * g_mvproj[r];
            if (g_pose3) {
                g += w.cons * mean_scale * g_pose3[r];                      // mean over B*V*J*3
                g += w.col * g_pose3[pose_stride + r];                      // SUM over batch and pairs (render.py:176)
                g += w.bone * mean_scale * g_pose3[2 * pose_stride + r];    // mean over B*35
            }
            if (g_prior) g += w.prior * 0.01f * g_prior[r];                 // x = xyz / 100; mean/sum split done in the VAE kernel
        }
        gxyz[i] = g;
    }
    sq = warp_sum(sq);
    if (threadIdx.x < 8) s_red[threadIdx.x] = 0.f;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && sq != 0.f) atomicAdd(&s_red[0], sq);
    __syncthreads();
    if (threadIdx.x == 0) {
        const float synt_d = Ns > 0 ? w.synt_pt * mean_scale * s_red[0] / (float)(Ns * J) : 0.f;
        if (synt_d != 0.f) { atomicAdd(&terms[1], synt_d); atomicAdd(&terms[8], synt_d); }
        if (blockIdx.x == 0) {
            float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (sse2) {
                if (Ns > 0) t[0] = w.synt_hm * mean_scale * (float)(sse2[0] / ((double)Ns * J * hw));
                if (M > 0) t[4] = w.hm_mean * mean_scale * (float)(sse2[1] / ((double)M * J * hw));
            }
            // batch-MEAN terms carry mean_scale = 1 / world_size (like their gradients), batch-SUM terms (collision; the VAE's KLD,
            // split inside its kernel) do not: the SUM over ranks of terms[] is then the loss of the global batch
            if (loss_mv3) t[2] = w.proj * mean_scale * loss_mv3[0];
            if (loss_pose3) { t[3] = w.cons * mean_scale * loss_pose3[0]; t[6] = w.col * loss_pose3[1]; t[7] = w.bone * mean_scale * loss_pose3[2]; }
            if (loss_prior3) t[5] = w.prior * loss_prior3[0];
            float tot = 0.f;
            for (int k = 0; k < 8; ++k) {
                if (k != 1 && t[k] != 0.f) atomicAdd(&terms[k], t[k]);
                tot += (k != 1) ? t[k] : 0.f;
            }
            atomicAdd(&terms[8], tot);
        }
    }
}

// Adam with L2 weight decay folded into the gradient (torch.optim.Adam semantics), learning rate and step count read
// from device memory so a captured CUDA graph can be replayed while the host-side scheduler changes them.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long n, const float* __restrict__ lr_dev, const int* __restrict__ step_dev, float b1, float b2,
                                float eps, float wd, float gscale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lr = *lr_dev;
    const float step = (float)(*step_dev);
    const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
    const float pi = p[i];
    const float gi = g[i] * gscale + wd * pi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

__global__ void step_increment_kernel(int* step_dev) { *step_dev += 1; }

__global__ void scale_kernel(const float4* __restrict__ x, float s, long n4, float4* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = x[i];
    y[i] = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
}
__global__ void scale_scalar_kernel(const float* __restrict__ x, float s, long n, float* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] * s;
}

// xyz[m][j][0] /= u[m], xyz[m][j][1] /= v[m] (IEEE division, as torch's): the undo of the scale augmentation on the real views'
// joints, create_network_and_criterion.py:124-126 -- and, applied to the gradient, its backward (d(x/u)/dx = 1/u).
__global__ void unscale_xy_kernel(float* __restrict__ xyz, const float* __restrict__ u, const float* __restrict__ v, int M, int J) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * J) return;
    const int m = i / J;
    xyz[3 * i] = __fdiv_rn(xyz[3 * i], u[m]);
    xyz[3 * i + 1] = __fdiv_rn(xyz[3 * i + 1], v[m]);
}

}  // namespace

SH_EXPORT int sh_unscale_xy(void* xyz, const void* u_scales, const void* v_scales, int M, int J, void* stream) {
    SH_REQUIRE(M >= 0 && J >= 1, "sh_unscale_xy: bad arguments");
    if (M == 0) return SH_OK;
    SH_REQUIRE(xyz && u_scales && v_scales, "sh_unscale_xy: null pointer");
    unscale_xy_kernel<<<sh_div_up(M * J, 256), 256, 0, (cudaStream_t)stream>>>((float*)xyz, (const float*)u_scales, (const float*)v_scales, M, J);
    SH_CHECK_LAUNCH("unscale_xy_kernel");
    return SH_OK;
}

SH_EXPORT int sh_step_combine(const void* g_mvproj, const void* g_pose3, const void* g_prior, const void* xyz,
                              const void* target_xyz4, const void* loss_mv3, const void* loss_pose3,
                              const void* loss_prior3, const void* sse2, int Ns, int M, int J, int hw,
                              const float* weights8, float mean_scale, void* gxyz, void* terms9, void* stream) {
    SH_REQUIRE(xyz && gxyz && terms9 && weights8, "sh_step_combine: null pointer");
    SH_REQUIRE(Ns >= 0 && M >= 0 && J >= 1 && hw >= 1 && (Ns == 0 || target_xyz4), "sh_step_combine: bad arguments");
    if (Ns + M == 0) return SH_OK;
    CombineW w{weights8[0], weights8[1], weights8[2], weights8[3], weights8[4], weights8[5], weights8[6], weights8[7]};
    const long total = (long)(Ns + M) * J * 3;
    const int blocks = (int)(sh_div_up(total, 256) < 64 ? sh_div_up(total, 256) : 64);
    step_combine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float*)g_mvproj, (const float*)g_pose3, (const float*)g_prior, (const float*)xyz, (const float4*)target_xyz4,
        (const float*)loss_mv3, (const float*)loss_pose3, (const float*)loss_prior3, (const double*)sse2, Ns, M, J, hw, w,
        mean_scale, (float*)gxyz, (float*)terms9);
    SH_CHECK_LAUNCH("step_combine_kernel");
    return SH_OK;
}

SH_EXPORT int sh_adam_step_dev(void* p, const void* g, void* m, void* v, long n, const void* lr_dev, void* step_dev,
                               float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
    SH_REQUIRE(p && g && m && v && lr_dev && step_dev && n >= 0, "sh_adam_step_dev: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    step_increment_kernel<<<1, 1, 0, st>>>((int*)step_dev);
    if (n > 0)
        adam_dev_kernel<<<sh_div_up(n, 256), 256, 0, st>>>((float*)p, (const float*)g, (float*)m, (float*)v, n,
                                                           (const float*)lr_dev, (const int*)step_dev, beta1, beta2, eps,
                                                           weight_decay, grad_scale);
    SH_CHECK_LAUNCH("adam_dev_kernel");
    return SH_OK;
}

SH_EXPORT int sh_scale(const void* x, float s, long n, void* y, void* stream) {
    SH_REQUIRE((x && y) || n == 0, "sh_scale: null pointer");
    SH_REQUIRE(n >= 0, "sh_scale: bad n");
    if (n == 0) return SH_OK;
    if (n % 4 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0)
        scale_kernel<<<sh_div_up(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, s, n / 4, (float4*)y);
    else
        scale_scalar_kernel<<<sh_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, s, n, (float*)y);
    SH_CHECK_LAUNCH("scale_kernel");
    return SH_OK;
}
