// R2 — differentiable sphere depth renderer, forward + analytic backward (sm_100a).
//
// Replaces BallRender.forward (/root/reference/mesh/render.py:26-53) followed by the min over the J
// spheres of one image (/root/reference/mesh/render.py:89, mesh/multiview_utility.py:76), and the
// gradient torch autograd derives from them.  One launch renders N images from N*J
// spheres without materialising the reference's [N*J,H,W] per-sphere maps.
//
// Data layout in HBM: spheres float4 [N*J] = (cx, cy, cz, r) (16 B aligned, 16*J B per image, staged
// into shared memory with one cp.async.bulk per CTA); depth fp32 [N,H,W]; idx uint8 [N,H,W]
// (winning sphere, 255 = background); grad_depth fp32 [N,H,W]; grad_spheres float4 [N*J]
// (dL/dcx, dL/dcy, dL/dcz, dL/dr).
//
// Arithmetic follows the reference operation by operation with explicit round-to-nearest intrinsics
// (no FMA contraction), so depth is bit-identical to an IEEE evaluation of render.py:31-52 and the
// arg-min index is exact.  Work is cut by a conservative warp-level bounding-box cull: a sphere is
// skipped for a 32x4 pixel tile only when it provably covers no pixel of the tile.
#include "common.cuh"
#include "sphere_common.cuh"

namespace {

constexpr int kThreads = 128;

template <int PX>
__global__ void __launch_bounds__(kThreads) sphere_render_fwd_kernel(
    const float4* __restrict__ spheres, int J, int H, int W, int tiles_per_block,
    float* __restrict__ depth, uint8_t* __restrict__ idx) {
    __shared__ __align__(128) float4 s_sph[kMaxJ];
    __shared__ __align__(8) uint64_t s_bar;
    const int n = blockIdx.y;
    stage_spheres(s_sph, &s_bar, spheres + (size_t)n * J, J);

    const TileGeom g = make_geom(W, H, PX);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = g.tiles_x * g.tiles_y;
    const int t_begin = blockIdx.x * tiles_per_block;
    const int t_end = min(t_begin + tiles_per_block, n_tiles);
    const float halfw = W * 0.5f, halfh = H * 0.5f, fw = (float)W, fh = (float)H;

    for (int t = t_begin + warp; t < t_end; t += kThreads / 32) {
        const int ty = t / g.tiles_x, tx = t - ty * g.tiles_x;
        const int row0 = ty * g.trows, col0 = tx * g.tq * PX;
        const int row1 = min(row0 + g.trows, H) - 1, col1 = min(col0 + g.tq * PX, W) - 1;
        // tile bounding box in mm (grid is monotone in u, v)
        const float bx0 = sh_grid_mm(col0, halfw, fw) - kCullMargin, bx1 = sh_grid_mm(col1, halfw, fw) + kCullMargin;
        const float by0 = sh_grid_mm(row0, halfh, fh) - kCullMargin, by1 = sh_grid_mm(row1, halfh, fh) + kCullMargin;
        uint32_t m[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            bool keep = false;
            if (k < J) {
                const float4 s = s_sph[k];
                // skip only if the sphere's |d| < r box misses the tile on some axis (NaN compares false -> kept)
                const bool miss = (s.x + s.w < bx0) || (s.x - s.w > bx1) || (s.y + s.w < by0) || (s.y - s.w > by1);
                keep = !miss;
            }
            m[h] = __ballot_sync(0xffffffffu, keep);
        }
        const int r = row0 + lane / g.tq;
        const int c = col0 + (lane % g.tq) * PX;
        const bool active = (r < H) && (c < W);
        float best[PX];
        int bidx[PX];
        float xg[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            best[p] = SH_BACKGROUND;
            bidx[p] = 255;
            xg[p] = sh_grid_mm(c + p, halfw, fw);
        }
        const float yg = sh_grid_mm(r, halfh, fh);
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
                    const float sv = __fsub_rn(__fsub_rn(r2, __fmul_rn(dx, dx)), ys);   // (r^2 - dx^2) - dy^2
                    if (sv > SH_S_MIN) {                                                // clamp(min=1e-2) != 1e-2
                        const float d = __fsub_rn(s.z, __fsqrt_rn(sv));
                        if (d < best[p]) {                                              // first minimum wins
                            best[p] = d;
                            bidx[p] = k;
                        }
                    }
                }
            }
        }
        if (active) {
            const size_t o = ((size_t)n * H + r) * W + c;
            if (PX == 4) {
                *reinterpret_cast<float4*>(depth + o) = make_float4(best[0], best[1], best[2], best[3]);
                *reinterpret_cast<uchar4*>(idx + o) =
                    make_uchar4((uint8_t)bidx[0], (uint8_t)bidx[1], (uint8_t)bidx[2], (uint8_t)bidx[3]);
            } else {
                depth[o] = best[0];
                idx[o] = (uint8_t)bidx[0];
            }
        }
    }
}

template <int PX>
__global__ void __launch_bounds__(kThreads) sphere_render_bwd_kernel(
    const float4* __restrict__ spheres, int J, int H, int W, int tiles_per_block,
    const float* __restrict__ grad_depth, const uint8_t* __restrict__ idx, float4* __restrict__ grad_spheres) {
    __shared__ __align__(128) float4 s_sph[kMaxJ];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ float s_acc[kMaxJ * 4];
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < kMaxJ * 4; i += kThreads) s_acc[i] = 0.f;
    stage_spheres(s_sph, &s_bar, spheres + (size_t)n * J, J);   // contains a __syncthreads()

    const TileGeom g = make_geom(W, H, PX);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = g.tiles_x * g.tiles_y;
    const int t_begin = blockIdx.x * tiles_per_block;
    const int t_end = min(t_begin + tiles_per_block, n_tiles);
    const float halfw = W * 0.5f, halfh = H * 0.5f, fw = (float)W, fh = (float)H;

    for (int t = t_begin + warp; t < t_end; t += kThreads / 32) {
        const int ty = t / g.tiles_x, tx = t - ty * g.tiles_x;
        const int r = ty * g.trows + lane / g.tq;
        const int c = tx * g.tq * PX + (lane % g.tq) * PX;
        if (r >= H || c >= W) continue;
        const size_t o = ((size_t)n * H + r) * W + c;
        float gd[PX];
        int ki[PX];
        if (PX == 4) {
            const float4 gv = *reinterpret_cast<const float4*>(grad_depth + o);
            const uchar4 kv = *reinterpret_cast<const uchar4*>(idx + o);
            gd[0] = gv.x; gd[1 % PX] = gv.y; gd[2 % PX] = gv.z; gd[3 % PX] = gv.w;
            ki[0] = kv.x; ki[1 % PX] = kv.y; ki[2 % PX] = kv.z; ki[3 % PX] = kv.w;
        } else {
            gd[0] = grad_depth[o];
            ki[0] = idx[o];
        }
        const float yg = sh_grid_mm(r, halfh, fh);
        // run-length merge over the lane's pixels before touching shared atomics
        int cur = -1;
        float ax = 0.f, ay = 0.f, az = 0.f, ar = 0.f;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int k = ki[p];
            if (k >= J) continue;   // background (255)
            if (k != cur) {
                if (cur >= 0) {
                    atomicAdd(&s_acc[cur * 4 + 0], ax); atomicAdd(&s_acc[cur * 4 + 1], ay);
                    atomicAdd(&s_acc[cur * 4 + 2], az); atomicAdd(&s_acc[cur * 4 + 3], ar);
                }
                cur = k; ax = ay = az = ar = 0.f;
            }
            const float4 s = s_sph[k];
            const float dx = __fsub_rn(sh_grid_mm(c + p, halfw, fw), s.x);
            const float dy = __fsub_rn(yg, s.y);
            const float sv = __fsub_rn(__fsub_rn(__fmul_rn(s.w, s.w), __fmul_rn(dx, dx)), __fmul_rn(dy, dy));
            const float inv = gd[p] / __fsqrt_rn(fmaxf(sv, SH_S_MIN));
            ax -= dx * inv;      // dd/dcx = -(xg-cx)/sqrt(s)
            ay -= dy * inv;
            az += gd[p];         // dd/dcz = 1
            ar -= s.w * inv;     // dd/dr  = -r/sqrt(s)
        }
        if (cur >= 0) {
            atomicAdd(&s_acc[cur * 4 + 0], ax); atomicAdd(&s_acc[cur * 4 + 1], ay);
            atomicAdd(&s_acc[cur * 4 + 2], az); atomicAdd(&s_acc[cur * 4 + 3], ar);
        }
    }
    __syncthreads();
    float* out = reinterpret_cast<float*>(grad_spheres + (size_t)n * J);
    for (int i = threadIdx.x; i < J * 4; i += kThreads) {
        const float v = s_acc[i];
        if (v != 0.f) atomicAdd(out + i, v);
    }
}

int pick_tiles_per_block(int n_tiles, int N) {
    // enough CTAs for >= 4 waves of 148 SMs x 8 resident CTAs when the batch allows, >= 4 tiles per warp otherwise
    int tpb = 16;
    while (tpb > 4 && (long)N * ((n_tiles + tpb - 1) / tpb) < 4L * SH_NUM_SMS * 8) tpb >>= 1;
    return tpb;
}

}  // namespace

SH_EXPORT int sh_sphere_render_fwd(const void* spheres, int N, int J, int H, int W, void* depth, void* idx,
                                    void* stream) {
    SH_REQUIRE(J >= 1 && J <= kMaxJ, "sh_sphere_render_fwd: J=%d outside [1,%d]", J, kMaxJ);
    SH_REQUIRE(N >= 0 && H >= 1 && W >= 1 && N <= 65535 * 64, "sh_sphere_render_fwd: bad N/H/W");
    if (N == 0) return SH_OK;   // empty batch: nothing to read or write (pointers may be null)
    SH_REQUIRE(spheres && depth && idx, "sh_sphere_render_fwd: null pointer");
    SH_REQUIRE(((uintptr_t)spheres & 15) == 0, "sh_sphere_render_fwd: spheres must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int px = (W % 4 == 0 && ((uintptr_t)depth & 15) == 0 && ((uintptr_t)idx & 3) == 0) ? 4 : 1;
    const TileGeom g = make_geom(W, H, px);
    const int n_tiles = g.tiles_x * g.tiles_y;
    const int tpb = pick_tiles_per_block(n_tiles, N);
    for (int n0 = 0; n0 < N; n0 += 65535) {
        const int nn = N - n0 < 65535 ? N - n0 : 65535;
        dim3 grid(sh_div_up(n_tiles, tpb), nn);
        const float4* sp = (const float4*)spheres + (size_t)n0 * J;
        float* dp = (float*)depth + (size_t)n0 * H * W;
        uint8_t* ip = (uint8_t*)idx + (size_t)n0 * H * W;
        if (px == 4)
            sphere_render_fwd_kernel<4><<<grid, kThreads, 0, st>>>(sp, J, H, W, tpb, dp, ip);
        else
            sphere_render_fwd_kernel<1><<<grid, kThreads, 0, st>>>(sp, J, H, W, tpb, dp, ip);
    }
    SH_CHECK_LAUNCH("sphere_render_fwd_kernel");
    return SH_OK;
}

SH_EXPORT int sh_sphere_render_bwd(const void* grad_depth, const void* idx, const void* spheres, int N, int J,
                                    int H, int W, void* grad_spheres, void* stream) {
    SH_REQUIRE(J >= 1 && J <= kMaxJ, "sh_sphere_render_bwd: J=%d outside [1,%d]", J, kMaxJ);
    SH_REQUIRE(N >= 0 && H >= 1 && W >= 1, "sh_sphere_render_bwd: bad N/H/W");
    if (N == 0) return SH_OK;
    SH_REQUIRE(grad_depth && idx && spheres && grad_spheres, "sh_sphere_render_bwd: null pointer");
    SH_REQUIRE(((uintptr_t)spheres & 15) == 0, "sh_sphere_render_bwd: spheres must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    SH_CUDA(cudaMemsetAsync(grad_spheres, 0, (size_t)N * J * 16, st));
    const int px = (W % 4 == 0 && ((uintptr_t)grad_depth & 15) == 0 && ((uintptr_t)idx & 3) == 0) ? 4 : 1;
    const TileGeom g = make_geom(W, H, px);
    const int n_tiles = g.tiles_x * g.tiles_y;
    const int tpb = pick_tiles_per_block(n_tiles, N);
    for (int n0 = 0; n0 < N; n0 += 65535) {
        const int nn = N - n0 < 65535 ? N - n0 : 65535;
        dim3 grid(sh_div_up(n_tiles, tpb), nn);
        const float4* sp = (const float4*)spheres + (size_t)n0 * J;
        const float* gp = (const float*)grad_depth + (size_t)n0 * H * W;
        const uint8_t* ip = (const uint8_t*)idx + (size_t)n0 * H * W;
        float4* op = (float4*)grad_spheres + (size_t)n0 * J;
        if (px == 4)
            sphere_render_bwd_kernel<4><<<grid, kThreads, 0, st>>>(sp, J, H, W, tpb, gp, ip, op);
        else
            sphere_render_bwd_kernel<1><<<grid, kThreads, 0, st>>>(sp, J, H, W, tpb, gp, ip, op);
    }
    SH_CHECK_LAUNCH("sphere_render_bwd_kernel");
    return SH_OK;
}
