// Stand-alone entries behind the reference's nn.Module surfaces that the fused train step does not need separately
// (sm_100a).  The fused step folds these into larger kernels (mvproj_loss.cu, synth.cu); the drop-in modules of
// spherehand_b200/mesh and spherehand_b200/network call them one by one, like the reference does.
//
// Replaces
//   DataToModelLoss.forward                 /root/reference/mesh/render.py:123-142      (+ autograd backward)
//   OthographicalProjection.forward         /root/reference/mesh/pointTransformation.py:84-99
//   InverseOthographicalProjection.forward  /root/reference/mesh/pointTransformation.py:118-124
//   RandScale.forward (matrix product part) /root/reference/mesh/pointTransformation.py:144-148
//   torch.clamp(depth_maps, max=100.0)      /root/reference/mesh/render.py:286
#include "sphere_common.cuh"

namespace {

constexpr int kThreads = 128;

// One block = a strip of rows of one image; spheres of the image staged in shared memory (TMA bulk copy).
// loss_acc[0] += sum over pixels of clamp(min_k | |P - c_k| - r_k |, 0, 50); gsph[n,k] += d(sum)/dc_k (unscaled).
__global__ void __launch_bounds__(kThreads) data_to_model_kernel(const float4* __restrict__ spheres, const float* __restrict__ dms,
                                                                 int J, int H, int W, int rows_per_block,
                                                                 float4* __restrict__ gsph, double* __restrict__ loss_acc) {
    __shared__ __align__(128) float4 s_sph[kMaxJ];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ float s_acc[kMaxJ * 3];
    __shared__ float s_loss;
    const int n = blockIdx.y;
    for (int t = threadIdx.x; t < kMaxJ * 3; t += kThreads) s_acc[t] = 0.f;
    if (threadIdx.x == 0) s_loss = 0.f;
    stage_spheres(s_sph, &s_bar, spheres + (size_t)n * J, J);
    const float halfw = W * 0.5f, halfh = H * 0.5f, fw = (float)W, fh = (float)H;
    const int r_begin = blockIdx.x * rows_per_block, r_end = min(r_begin + rows_per_block, H);
    const float* img = dms + (size_t)n * H * W;
    float l = 0.f;
    for (int p = r_begin * W + threadIdx.x; p < r_end * W; p += kThreads) {
        const int r = p / W, c = p - r * W;
        const float z = img[p];
        if (z > 99.f) continue;                                   // background pixels contribute 0 (render.py:137-139)
        const float xg = sh_grid_mm(c, halfw, fw), yg = sh_grid_mm(r, halfh, fh);
        float e_best = 3.4e38f, dist_b = 1.f, sgn_b = 0.f;
        int kb = 0;
        for (int k = 0; k < J; ++k) {
            const float4 s = s_sph[k];
            const float dx = xg - s.x, dy = yg - s.y, dz = z - s.z;
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
            const float sd = dist - s.w;
            const float e = fabsf(sd);
            if (e < e_best) { e_best = e; kb = k; dist_b = dist; sgn_b = sd; }
        }
        l += fminf(e_best, 50.f);
        if (e_best > 0.f && e_best <= 50.f && dist_b > 0.f) {
            const float4 s = s_sph[kb];
            const float coef = (sgn_b > 0.f ? 1.f : -1.f) / dist_b;      // d|d-r|/dc = sign * (c - P) / dist
            atomicAdd(&s_acc[kb * 3 + 0], coef * (s.x - xg));
            atomicAdd(&s_acc[kb * 3 + 1], coef * (s.y - yg));
            atomicAdd(&s_acc[kb * 3 + 2], coef * (s.z - z));
        }
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_loss, l);
    __syncthreads();
    float* out = reinterpret_cast<float*>(gsph + (size_t)n * J);
    for (int t = threadIdx.x; t < J * 3; t += kThreads) {
        const float v = s_acc[t];
        if (v != 0.f) atomicAdd(out + (t / 3) * 4 + (t % 3), v);
    }
    if (threadIdx.x == 0) atomicAdd(loss_acc, (double)s_loss);
}

__global__ void data_to_model_pack_kernel(const float* __restrict__ joints, const float* __restrict__ radii, int J, long total,
                                          float4* __restrict__ spheres, float4* __restrict__ gsph) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    spheres[i] = make_float4(joints[i * 3], joints[i * 3 + 1], joints[i * 3 + 2], radii[i % J]);
    gsph[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void data_to_model_finish_kernel(const float4* __restrict__ gsph, const double* __restrict__ loss_acc, long total,
                                            double inv_count, float* __restrict__ loss, float* __restrict__ grad_joints) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        const float4 g = gsph[i];
        grad_joints[i * 3] = (float)(g.x * inv_count);
        grad_joints[i * 3 + 1] = (float)(g.y * inv_count);
        grad_joints[i * 3 + 2] = (float)(g.z * inv_count);
    }
    if (i == 0) loss[0] = (float)(loss_acc[0] * inv_count);
}

// mode 1: u = x*rand_f[b]*fx + cx, v = y*rand_f[b]*fy + cy, (z, 1)   (pointTransformation.py:90-97: w is set to ONE)
// mode 2: K * p with K = [[fx,0,0,cx],[0,fy,0,cy],[0,0,1,0],[0,0,0,1]]   (:88-89: w is propagated)
// mode 3: K^-1 * p                                                       (InverseOthographicalProjection, :118-124)
__global__ void ortho_project_kernel(const float4* __restrict__ in, int nv, long total, int mode, float cx, float cy, float fx,
                                     float fy, const float* __restrict__ rand_f, float4* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float4 p = in[i];
    float4 o;
    if (mode == 1) {
        const float f = rand_f[i / nv];
        o = make_float4(__fadd_rn(__fmul_rn(__fmul_rn(p.x, f), fx), cx), __fadd_rn(__fmul_rn(__fmul_rn(p.y, f), fy), cy), p.z, 1.f);
    } else if (mode == 2) {
        o = make_float4(fmaf(fx, p.x, cx * p.w), fmaf(fy, p.y, cy * p.w), p.z, p.w);
    } else {
        const float ix = 1.f / fx, iy = 1.f / fy;
        o = make_float4(fmaf(ix, p.x, -cx * ix * p.w), fmaf(iy, p.y, -cy * iy * p.w), p.z, p.w);
    }
    out[i] = o;
}

// out[b,m] = diag(sx,sy,sz,1) * mats[b,m]: rows 0..2 of every 4x4 scaled
__global__ void rand_scale_kernel(const float* __restrict__ mats, const float* __restrict__ scales, int nmat, long total,
                                  float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // element index
    if (i >= total) return;
    const int row = (int)((i >> 2) & 3);
    const long b = i / (16L * nmat);
    out[i] = row < 3 ? scales[b * 3 + row] * mats[i] : mats[i];
}

// ResizeCropImage.forward (network/util_modules.py:388-424): nearest-neighbour resize by (v_scale, u_scale) pasted into the
// centre of an all-ones canvas, all images in one launch (the reference loops over images in Python, one interpolate
// each).  The integer geometry is evaluated with the reference's fp32 expressions: int(size*scale + 0.5), int(size*scale),
// floor(dst * (in/out)) of F.interpolate(mode='nearest').  An image with v_scale > 1 stays all ones (the reference's paste
// sits inside the `else` of `if v_scale > 1`).
__global__ void resize_crop_kernel(const float* __restrict__ in, const float* __restrict__ us, const float* __restrict__ vs, int H, int W,
                                   float* __restrict__ out) {
    const int n = blockIdx.y;
    const float u = us[n], v = vs[n];
    const int new_w = (int)__fadd_rn(__fmul_rn((float)W, u), 0.5f), new_h = (int)__fadd_rn(__fmul_rn((float)H, v), 0.5f);
    int u_start, u_cnt, ou_start;
    if (u > 1.f) { u_start = 0; u_cnt = W; ou_start = (new_w - W) / 2; }
    else { ou_start = 0; u_cnt = (int)__fmul_rn((float)W, u); u_start = (W - new_w) / 2; }
    const bool paste = !(v > 1.f);
    const int v_cnt = (int)__fmul_rn((float)H, v), v_start = (H - new_h) / 2;
    const float sx = new_w > 0 ? (float)W / (float)new_w : 0.f, sy = new_h > 0 ? (float)H / (float)new_h : 0.f;
    const float* img = in + (size_t)n * H * W;
    float* dst = out + (size_t)n * H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
        const int y = p / W, x = p - y * W;
        float val = 1.f;
        if (paste && y >= v_start && y < v_start + v_cnt && x >= u_start && x < u_start + u_cnt) {
            const int ry = y - v_start, rx = ou_start + (x - u_start);                 // position in the resized image
            const int iy = min((int)floorf(__fmul_rn((float)ry, sy)), H - 1), ix = min((int)floorf(__fmul_rn((float)rx, sx)), W - 1);
            val = img[iy * W + ix];
        }
        dst[p] = val;
    }
}

__global__ void clamp_max_kernel(const float* __restrict__ x, long n, float mx, float* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = x[i]; y[i] = (v != v) ? v : fminf(v, mx); }   // NaN propagates like torch.clamp
}

}  // namespace

SH_EXPORT size_t sh_data_to_model_scratch_bytes(int N, int J) { return (size_t)N * J * 32 + 16; }

// dms [N,H,W] (mm, background > 99), joints [N,J,3], radii [J] -> loss[1] = mean over ALL N*H*W pixels of the clamped
// point-to-surface distance, grad_joints [N,J,3] = d loss / d joints.  scratch: sh_data_to_model_scratch_bytes, 16-byte aligned.
SH_EXPORT int sh_data_to_model_fwdbwd(const void* dms, const void* joints, const void* radii, int N, int J, int H, int W,
                                       void* loss, void* grad_joints, void* scratch, void* stream) {
    SH_REQUIRE(dms && joints && radii && loss && grad_joints && scratch, "sh_data_to_model_fwdbwd: null pointer");
    SH_REQUIRE(J >= 1 && J <= kMaxJ, "sh_data_to_model_fwdbwd: J=%d outside [1,%d]", J, kMaxJ);
    SH_REQUIRE(N >= 1 && N <= 65535 && H >= 1 && W >= 1, "sh_data_to_model_fwdbwd: bad N/H/W");
    SH_REQUIRE(((uintptr_t)scratch & 15) == 0, "sh_data_to_model_fwdbwd: scratch must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const long total = (long)N * J;
    float4* spheres = (float4*)scratch;
    float4* gsph = spheres + total;
    double* acc = (double*)(gsph + total);
    SH_CUDA(cudaMemsetAsync(acc, 0, 8, st));
    data_to_model_pack_kernel<<<sh_div_up(total, 256), 256, 0, st>>>((const float*)joints, (const float*)radii, J, total, spheres, gsph);
    SH_CHECK_LAUNCH("data_to_model_pack_kernel");
    int rpb = H;
    while (rpb > 4 && (long)N * ((H + rpb - 1) / rpb) < 4L * SH_NUM_SMS) rpb = (rpb + 1) / 2;
    dim3 grid(sh_div_up(H, rpb), N);
    data_to_model_kernel<<<grid, kThreads, 0, st>>>(spheres, (const float*)dms, J, H, W, rpb, gsph, acc);
    SH_CHECK_LAUNCH("data_to_model_kernel");
    data_to_model_finish_kernel<<<sh_div_up(total, 256), 256, 0, st>>>(gsph, acc, total, 1.0 / ((double)N * H * W), (float*)loss,
                                                                      (float*)grad_joints);
    SH_CHECK_LAUNCH("data_to_model_finish_kernel");
    return SH_OK;
}

// points float4 [B,Nv] -> out float4 [B,Nv]; mode 1 needs rand_f [B].
SH_EXPORT int sh_ortho_project(const void* points, int B, int Nv, int mode, float cx, float cy, float fx, float fy,
                                const void* rand_f, void* out, void* stream) {
    SH_REQUIRE(points && out && mode >= 1 && mode <= 3 && (mode != 1 || rand_f), "sh_ortho_project: bad arguments");
    const long total = (long)B * Nv;
    if (total == 0) return SH_OK;
    ortho_project_kernel<<<sh_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)points, Nv, total, mode, cx, cy, fx, fy,
                                                                                 (const float*)rand_f, (float4*)out);
    SH_CHECK_LAUNCH("ortho_project_kernel");
    return SH_OK;
}

// mats [B,nmat,4,4], scales [B,3] -> out = diag(scale,1) * mats
SH_EXPORT int sh_rand_scale_apply(const void* mats, const void* scales, int B, int nmat, void* out, void* stream) {
    SH_REQUIRE(mats && scales && out, "sh_rand_scale_apply: null pointer");
    const long total = (long)B * nmat * 16;
    if (total == 0) return SH_OK;
    rand_scale_kernel<<<sh_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)mats, (const float*)scales, nmat, total,
                                                                              (float*)out);
    SH_CHECK_LAUNCH("rand_scale_kernel");
    return SH_OK;
}

SH_EXPORT int sh_clamp_max(const void* x, long n, float max_value, void* y, void* stream) {
    SH_REQUIRE((x && y) || n == 0, "sh_clamp_max: null pointer");
    if (n <= 0) return SH_OK;
    clamp_max_kernel<<<sh_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, n, max_value, (float*)y);
    SH_CHECK_LAUNCH("clamp_max_kernel");
    return SH_OK;
}

// depth_maps [N,H,W], u_scales / v_scales [N] -> out [N,H,W]
SH_EXPORT int sh_resize_crop(const void* depth_maps, const void* u_scales, const void* v_scales, int N, int H, int W, void* out,
                              void* stream) {
    SH_REQUIRE(N >= 0 && H >= 1 && W >= 1 && N <= 65535, "sh_resize_crop: bad N/H/W");
    if (N == 0) return SH_OK;
    SH_REQUIRE(depth_maps && u_scales && v_scales && out, "sh_resize_crop: null pointer");
    dim3 grid(sh_div_up((long)H * W, 1024) < 16 ? sh_div_up((long)H * W, 1024) : 16, N);
    resize_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)depth_maps, (const float*)u_scales, (const float*)v_scales,
                                                                H, W, (float*)out);
    SH_CHECK_LAUNCH("resize_crop_kernel");
    return SH_OK;
}
