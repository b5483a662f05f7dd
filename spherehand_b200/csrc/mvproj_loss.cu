// Fused MutualProjectionLoss: forward value + analytic gradient in one pass (sm_100a).
//
// Replaces, per step, the chain
//   MutualTransformation      /root/reference/mesh/multiview_utility.py:13-30   T_ij = inv_cam_j * cam_i
//   MutualProjection          /root/reference/mesh/multiview_utility.py:55-77   p_ijk = R_ij joint_ik + t_ij, BallRender, min over J
//   MutualProjectionLoss      /root/reference/mesh/multiview_utility.py:90-130  9*MSE(proj, real_j) + 500*9*DataToModel
//   DataToModelLoss           /root/reference/mesh/render.py:123-142            mean clamp(min_k | |P-c_k| - r_k |, 0, 50)
// and the backward autograd derives for them.  The loss is a scalar whose
// upstream gradient is a scalar, so the kernel emits d(loss)/d(joints) for upstream 1 and the
// autograd.Function scales it; nothing but projected_dms (part of the API) is written per pixel.
//
// HBM layout: cam, inv_cam fp32 [B,V,4,4]; joints fp32 [B,V,J,3] (mm); real fp32 [B,V,H,W] (mm, background
// 100.0); radii fp32 [J]; projected_dms fp32 [B,V,V,H,W]; scratch: spheres float4 [B*V*V*J] (p_ij, r),
// xf fp32 [B*V*V,12] (rows of [R|t]), gsph float4 [B*V*V*J], acc double[2]; out: loss fp32[3]
// (total, model_to_data, data_to_model), grad_joints fp32 [B,V,J,3].
// Algorithmic bytes per (b,i,j) pair: 4HW (real in) + 4HW (projected out) + 28J + 96.
#include "common.cuh"
#include "sphere_common.cuh"

namespace {

constexpr int kThreads = 128;

__global__ void mvproj_prep_kernel(const float* __restrict__ cam, const float* __restrict__ inv_cam,
                                   const float* __restrict__ joints, const float* __restrict__ radii, int V, int J,
                                   float4* __restrict__ spheres, float* __restrict__ xf) {
    __shared__ float T[12];
    const int pair = blockIdx.x;              // (b*V + i)*V + j
    const int j = pair % V, i = (pair / V) % V, b = pair / (V * V);
    const float* A = inv_cam + ((size_t)b * V + j) * 16;
    const float* C = cam + ((size_t)b * V + i) * 16;
    if (threadIdx.x < 12) {
        const int r = threadIdx.x / 4, c = threadIdx.x % 4;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += A[r * 4 + k] * C[k * 4 + c];
        T[threadIdx.x] = s;
        xf[(size_t)pair * 12 + threadIdx.x] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < J; k += blockDim.x) {
        const float* q = joints + (((size_t)b * V + i) * J + k) * 3;
        const float x = q[0], y = q[1], z = q[2];
        float4 o;
        o.x = T[0] * x + T[1] * y + T[2] * z + T[3];
        o.y = T[4] * x + T[5] * y + T[6] * z + T[7];
        o.z = T[8] * x + T[9] * y + T[10] * z + T[11];
        o.w = radii[k];
        spheres[(size_t)pair * J + k] = o;
    }
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int PX>
__global__ void __launch_bounds__(kThreads) mvproj_main_kernel(
    const float4* __restrict__ spheres, const float* __restrict__ real, int V, int J, int H, int W,
    int tiles_per_block, float w_mv, float w_diag, int is_mv, float* __restrict__ projected,
    float4* __restrict__ gsph, double* __restrict__ acc) {
    __shared__ __align__(128) float4 s_sph[kMaxJ];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) float s_acc[kThreads / 32][kMaxJ * 4];      // one private gradient table per warp (dcx, dcy, dcz, -)
    __shared__ float s_loss[2];
    const int pair = blockIdx.y;
    const int j = pair % V, i = (pair / V) % V, b = pair / (V * V);
    for (int t = threadIdx.x; t < (kThreads / 32) * kMaxJ * 4; t += kThreads) (&s_acc[0][0])[t] = 0.f;
    if (threadIdx.x < 2) s_loss[threadIdx.x] = 0.f;
    stage_spheres(s_sph, &s_bar, spheres + (size_t)pair * J, J);

    const float wpx = is_mv ? w_mv : (i == j ? w_diag : 0.f);   // weight of one pixel of this pair in both means
    const TileGeom g = make_geom(W, H, PX);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = g.tiles_x * g.tiles_y;
    const int t_begin = blockIdx.x * tiles_per_block;
    const int t_end = min(t_begin + tiles_per_block, n_tiles);
    const float halfw = W * 0.5f, halfh = H * 0.5f, fw = (float)W, fh = (float)H;
    const float* real_img = real + ((size_t)b * V + j) * H * W;
    float* proj_img = projected + (size_t)pair * H * W;
    float l_m2d = 0.f, l_d2m = 0.f;
    float* acc_w = s_acc[threadIdx.x >> 5];

    for (int t = t_begin + warp; t < t_end; t += kThreads / 32) {
        const int ty = t / g.tiles_x, tx = t - ty * g.tiles_x;
        const int row0 = ty * g.trows, col0 = tx * g.tq * PX;
        const int row1 = min(row0 + g.trows, H) - 1, col1 = min(col0 + g.tq * PX, W) - 1;
        const float bx0 = sh_grid_mm(col0, halfw, fw) - kCullMargin, bx1 = sh_grid_mm(col1, halfw, fw) + kCullMargin;
        const float by0 = sh_grid_mm(row0, halfh, fh) - kCullMargin, by1 = sh_grid_mm(row1, halfh, fh) + kCullMargin;
        uint32_t m[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            bool keep = false;
            if (k < J) {
                const float4 s = s_sph[k];
                const bool miss = (s.x + s.w < bx0) || (s.x - s.w > bx1) || (s.y + s.w < by0) || (s.y - s.w > by1);
                keep = !miss;
            }
            m[h] = __ballot_sync(0xffffffffu, keep);
        }
        const int r = row0 + lane / g.tq;
        const int c = col0 + (lane % g.tq) * PX;
        const bool active = (r < H) && (c < W);
        float best[PX], bsq[PX], xg[PX];
        int bidx[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            best[p] = SH_BACKGROUND;
            bidx[p] = -1;
            bsq[p] = 1.f;
            xg[p] = sh_grid_mm(c + p, halfw, fw);
        }
        const float yg = sh_grid_mm(r, halfh, fh);
        // ---- render view i's spheres in view j's frame (same arithmetic as sphere_render_fwd_kernel)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t mm = m[h];
            while (mm) {
                const int k = __ffs(mm) - 1 + 32 * h;
                mm &= mm - 1;
                const float4 s = s_sph[k];
                const float dy = __fsub_rn(yg, s.y);
                const float r2 = __fmul_rn(s.w, s.w);
                const float ys = __fmul_rn(dy, dy);
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const float dx = __fsub_rn(xg[p], s.x);
                    const float sv = __fsub_rn(__fsub_rn(r2, __fmul_rn(dx, dx)), ys);
                    if (sv > SH_S_MIN) {
                        const float sq = __fsqrt_rn(sv);
                        const float d = __fsub_rn(s.z, sq);
                        if (d < best[p]) {
                            best[p] = d;
                            bidx[p] = k;
                            bsq[p] = sq;
                        }
                    }
                }
            }
        }
        const size_t o = (size_t)r * W + c;
        float z[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) z[p] = SH_BACKGROUND;
        if (active) {
            if (PX == 4) {
                *reinterpret_cast<float4*>(proj_img + o) = make_float4(best[0], best[1 % PX], best[2 % PX], best[3 % PX]);
                const float4 zv = *reinterpret_cast<const float4*>(real_img + o);
                z[0] = zv.x; z[1 % PX] = zv.y; z[2 % PX] = zv.z; z[3 % PX] = zv.w;
            } else {
                proj_img[o] = best[0];
                z[0] = real_img[o];
            }
        }
        if (wpx == 0.f) continue;                          // block-uniform: off-diagonal pair of a single-view step
        // Gradients go to the warp's private table through warp_accumulate (shuffle sums per distinct sphere, one plain
        // read-modify-write): the shared-memory float atomics this replaces are CAS loops, 32-way contended inside a sphere.
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            // ---- model -> data: (proj - real)^2, gradient through the arg-min sphere
            const float diff = active ? best[p] - z[p] : 0.f;
            l_m2d += diff * diff;
            const bool has1 = active && bidx[p] >= 0;
            float gx = 0.f, gy = 0.f, gz = 0.f;
            if (has1) {
                const float4 s = s_sph[bidx[p]];
                const float gd = 2.f * wpx * diff;
                const float inv = gd / bsq[p];
                gx = -(xg[p] - s.x) * inv;
                gy = -(yg - s.y) * inv;
                gz = gd;
            }
            warp_accumulate(acc_w, has1, bidx[p], gx, gy, gz, 0.f, lane);
        }
        // ---- data -> model: distance of the observed point to the nearest sphere surface.  The lane's PX pixels
        // share ONE pass over the J spheres (one broadcast load per sphere, PX independent dependency chains); the search uses
        // the one-instruction approximate square root (this loop is what the kernel spends its time in), then the winner is
        // evaluated with the IEEE one: value and gradient are exact for the sphere found, and the choice can differ from an
        // exact search only between candidates closer than ~1e-7 relative.
        bool fg[PX];
        bool any_fg = false;
        float e_search[PX];
        int kb[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            fg[p] = active && !(z[p] > 99.f);
            any_fg |= fg[p];
            e_search[p] = 3.4e38f;
            kb[p] = 0;
        }
        if (any_fg) {
            for (int k = 0; k < J; ++k) {
                const float4 s = s_sph[k];
                const float dy = yg - s.y;
                const float dy2 = dy * dy;
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    const float dx = xg[p] - s.x, dz = z[p] - s.z;
                    const float e = fabsf(sqrt_approx(dx * dx + dy2 + dz * dz) - s.w);
                    if (e < e_search[p]) { e_search[p] = e; kb[p] = k; }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            bool has2 = false;
            float hx = 0.f, hy = 0.f, hz = 0.f;
            if (fg[p]) {
                const float4 sb = s_sph[kb[p]];
                const float bx = xg[p] - sb.x, by = yg - sb.y, bz = z[p] - sb.z;
                const float dist_b = sqrtf(bx * bx + by * by + bz * bz);
                const float sgn_b = dist_b - sb.w;
                const float e_best = fabsf(sgn_b);
                l_d2m += fminf(e_best, 50.f);
                if (e_best > 0.f && e_best <= 50.f && dist_b > 0.f) {
                    const float coef = 500.f * wpx * (sgn_b > 0.f ? 1.f : -1.f) / dist_b;   // d|d-r|/dc = sign*(c-P)/dist
                    has2 = true;
                    hx = coef * (sb.x - xg[p]);
                    hy = coef * (sb.y - yg);
                    hz = coef * (sb.z - z[p]);
                }
            }
            warp_accumulate(acc_w, has2, kb[p], hx, hy, hz, 0.f, lane);
        }
    }
    l_m2d = warp_sum(l_m2d);
    l_d2m = warp_sum(l_d2m);
    if (lane == 0 && wpx != 0.f) {
        atomicAdd(&s_loss[0], l_m2d);
        atomicAdd(&s_loss[1], l_d2m);
    }
    __syncthreads();
    if (wpx == 0.f) return;
    float* out = reinterpret_cast<float*>(gsph + (size_t)pair * J);
    for (int t = threadIdx.x; t < J * 3; t += kThreads) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += s_acc[w][(t / 3) * 4 + (t % 3)];
        if (v != 0.f) atomicAdd(out + (t / 3) * 4 + (t % 3), v);
    }
    if (threadIdx.x < 2) atomicAdd(&acc[threadIdx.x], (double)s_loss[threadIdx.x] * (double)wpx);
}

__global__ void mvproj_finish_kernel(const float4* __restrict__ gsph, const float* __restrict__ xf, int V, int J,
                                     const double* __restrict__ acc, float* __restrict__ loss,
                                     float* __restrict__ grad_joints) {
    const int bi = blockIdx.x;   // b*V + i
    for (int t = threadIdx.x; t < J * 3; t += blockDim.x) {
        const int k = t / 3, c = t % 3;
        float s = 0.f;
        for (int j = 0; j < V; ++j) {
            const size_t pair = (size_t)bi * V + j;
            const float4 gv = gsph[pair * J + k];
            const float* T = xf + pair * 12;
            s += T[0 + c] * gv.x + T[4 + c] * gv.y + T[8 + c] * gv.z;    // R^T g
        }
        grad_joints[((size_t)bi * J + k) * 3 + c] = s;
    }
    if (bi == 0 && threadIdx.x == 0) {
        const double m2d = acc[0], d2m = acc[1];
        loss[0] = (float)(m2d + 500.0 * d2m);
        loss[1] = (float)m2d;
        loss[2] = (float)d2m;
    }
}

}  // namespace

SH_EXPORT size_t sh_mvproj_scratch_bytes(int B, int V, int J) {
    const size_t pairs = (size_t)B * V * V;
    return pairs * J * 16 * 2 + pairs * 12 * 4 + 64;
}

SH_EXPORT int sh_mvproj_loss_fwdbwd(const void* cam, const void* inv_cam, const void* joints, const void* real,
                                     const void* radii, int B, int V, int J, int H, int W, int is_mv,
                                     void* projected_dms, void* loss3, void* grad_joints, void* scratch,
                                     void* stream) {
    SH_REQUIRE(cam && inv_cam && joints && real && radii && projected_dms && loss3 && grad_joints && scratch,
               "sh_mvproj_loss_fwdbwd: null pointer");
    SH_REQUIRE(J >= 1 && J <= kMaxJ, "sh_mvproj_loss_fwdbwd: J=%d outside [1,%d]", J, kMaxJ);
    SH_REQUIRE(B >= 1 && V >= 1 && V <= 8 && H >= 1 && W >= 1, "sh_mvproj_loss_fwdbwd: bad B/V/H/W");
    SH_REQUIRE((long)B * V * V <= 65535, "sh_mvproj_loss_fwdbwd: B*V*V > 65535");
    SH_REQUIRE(((uintptr_t)scratch & 15) == 0, "sh_mvproj_loss_fwdbwd: scratch must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int pairs = B * V * V;
    char* sc = (char*)scratch;
    float4* spheres = (float4*)sc;
    float4* gsph = (float4*)(sc + (size_t)pairs * J * 16);
    float* xf = (float*)(sc + (size_t)pairs * J * 32);
    double* acc = (double*)(sc + (size_t)pairs * J * 32 + (size_t)pairs * 48);
    SH_REQUIRE(((uintptr_t)acc & 7) == 0, "sh_mvproj_loss_fwdbwd: internal alignment");
    SH_CUDA(cudaMemsetAsync(gsph, 0, (size_t)pairs * J * 16, st));
    SH_CUDA(cudaMemsetAsync(acc, 0, 16, st));
    mvproj_prep_kernel<<<pairs, 64, 0, st>>>((const float*)cam, (const float*)inv_cam, (const float*)joints,
                                             (const float*)radii, V, J, spheres, xf);
    SH_CHECK_LAUNCH("mvproj_prep_kernel");
    // per-pixel weights: is_mv -> 9/(B V^2 H W) on every pair; else 3/(B H W) on the diagonal (multiview_utility.py:100-127)
    const float w_mv = (float)(9.0 / ((double)B * V * V * H * W));
    const float w_diag = (float)(3.0 / ((double)B * H * W));
    const int px = (W % 4 == 0 && ((uintptr_t)real & 15) == 0 && ((uintptr_t)projected_dms & 15) == 0) ? 4 : 1;
    const TileGeom g = make_geom(W, H, px);
    const int n_tiles = g.tiles_x * g.tiles_y;
    int tpb = 16;
    while (tpb > 4 && (long)pairs * ((n_tiles + tpb - 1) / tpb) < 4L * SH_NUM_SMS * 8) tpb >>= 1;
    dim3 grid(sh_div_up(n_tiles, tpb), pairs);
    if (px == 4)
        mvproj_main_kernel<4><<<grid, kThreads, 0, st>>>(spheres, (const float*)real, V, J, H, W, tpb, w_mv, w_diag,
                                                         is_mv, (float*)projected_dms, gsph, acc);
    else
        mvproj_main_kernel<1><<<grid, kThreads, 0, st>>>(spheres, (const float*)real, V, J, H, W, tpb, w_mv, w_diag,
                                                         is_mv, (float*)projected_dms, gsph, acc);
    SH_CHECK_LAUNCH("mvproj_main_kernel");
    mvproj_finish_kernel<<<B * V, 128, 0, st>>>(gsph, xf, V, J, acc, (float*)loss3, (float*)grad_joints);
    SH_CHECK_LAUNCH("mvproj_finish_kernel");
    return SH_OK;
}
