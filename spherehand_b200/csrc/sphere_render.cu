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

// Forward: sphere-major rasterisation into a shared-memory z-buffer.
// A pixel-major loop tests every sphere whose bounding box touches the pixel's TILE (6-7 per pixel at J = 48) although only
// ~1.5 cover the pixel: the kernel was instruction-bound at 17 % of the HBM roofline.  Here a CTA owns a band of image rows;
// each warp takes spheres, its lanes walk the sphere's own pixel bounding box (clipped to the band), and every covered pixel
// does ONE native 64-bit shared-memory atomicMin on the key (order-preserving depth bits << 32 | sphere index): the smallest
// depth wins and, among equal depths, the smallest index -- exactly torch.min's first minimum.  The per-pixel arithmetic is
// unchanged (explicit _rn operations in the reference's order), so depth stays bit-identical.  A final pass decodes the keys
// and writes depth / idx with coalesced 16-byte stores.
constexpr int kFwdThreads = 256;
constexpr int kZbufBytes = 32 * 1024;           // per CTA: band rows x W 64-bit keys

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

__global__ void __launch_bounds__(kFwdThreads) sphere_render_fwd_kernel(
    const float4* __restrict__ spheres, int J, int H, int W, int band_rows,
    float* __restrict__ depth, uint8_t* __restrict__ idx) {
    extern __shared__ __align__(16) unsigned long long s_z[];        // [band_rows][W] keys, then xg[W], yg[band_rows]
    __shared__ __align__(128) float4 s_sph[kMaxJ];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_next;
    const int n = blockIdx.y;
    const int row0 = blockIdx.x * band_rows;
    const int rows = min(band_rows, H - row0);
    const int npx = rows * W;
    float* s_xg = reinterpret_cast<float*>(s_z + (size_t)band_rows * W);
    float* s_yg = s_xg + W;
    const float halfw = W * 0.5f, halfh = H * 0.5f, fw = (float)W, fh = (float)H;
    const unsigned long long bg = ((unsigned long long)float_to_ordered(SH_BACKGROUND) << 32) | 255ull;
    for (int i = threadIdx.x; i < npx; i += kFwdThreads) s_z[i] = bg;
    // the reference pixel grid, one IEEE division per row / column instead of two per pixel test
    for (int i = threadIdx.x; i < W; i += kFwdThreads) s_xg[i] = sh_grid_mm(i, halfw, fw);
    for (int i = threadIdx.x; i < rows; i += kFwdThreads) s_yg[i] = sh_grid_mm(row0 + i, halfh, fh);
    if (threadIdx.x == 0) s_next = kFwdThreads / 32;
    stage_spheres(s_sph, &s_bar, spheres + (size_t)n * J, J);      // contains a __syncthreads()

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float px_per_mm_x = fw / 300.0f, px_per_mm_y = fh / 300.0f;
    int k = warp;                                                    // warps pull spheres from a shared counter: box sizes vary
    while (k < J) {
        const float4 s = s_sph[k];
        // pixel bounding box of |d| < r (exact test below; 1e-2 px of slack covers the rounding of this estimate), clipped to the band
        const float fu0 = (s.x - s.w) * px_per_mm_x + halfw - 1e-2f, fu1 = (s.x + s.w) * px_per_mm_x + halfw + 1e-2f;
        const float fv0 = (s.y - s.w) * px_per_mm_y + halfh - 1e-2f, fv1 = (s.y + s.w) * px_per_mm_y + halfh + 1e-2f;
        const bool hit = fu1 >= 0.f && fu0 <= fw - 1.f && fv1 >= (float)row0 && fv0 <= (float)(row0 + rows - 1);   // NaN -> false
        if (hit) {
            const int u0 = max((int)floorf(fu0), 0), u1 = min((int)ceilf(fu1), W - 1);
            const int v0 = max((int)floorf(fv0), row0), v1 = min((int)ceilf(fv1), row0 + rows - 1);
            const int bw = u1 - u0 + 1, total = bw * (v1 - v0 + 1);
            const float inv_bw = 1.0f / (float)bw;
            const float r2 = __fmul_rn(s.w, s.w);
            for (int i = lane; i < total; i += 32) {
                const int vy = (int)(((float)i + 0.5f) * inv_bw);   // exact for i < 2^20
                const int u = u0 + (i - vy * bw), v = v0 + vy - row0;
                const float dx = __fsub_rn(s_xg[u], s.x);
                const float dy = __fsub_rn(s_yg[v], s.y);
                const float sv = __fsub_rn(__fsub_rn(r2, __fmul_rn(dx, dx)), __fmul_rn(dy, dy));   // (r^2 - dx^2) - dy^2
                if (sv > SH_S_MIN) {                                                                // clamp(min=1e-2) != 1e-2
                    const float d = __fsub_rn(s.z, __fsqrt_rn(sv));
                    if (d < SH_BACKGROUND)                                                          // the reference's `d < best` against 100.0
                        atomicMin(&s_z[v * W + u], ((unsigned long long)float_to_ordered(d) << 32) | (unsigned long long)k);
                }
            }
        }
        if (lane == 0) k = atomicAdd(&s_next, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
    }
    __syncthreads();
    float* dp = depth + ((size_t)n * H + row0) * W;
    uint8_t* ip = idx + ((size_t)n * H + row0) * W;
    if ((W & 3) == 0 && (((uintptr_t)dp | (uintptr_t)ip) & 15) == 0) {
        for (int i = threadIdx.x * 4; i < npx; i += kFwdThreads * 4) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(s_z + i), b = *reinterpret_cast<const ulonglong2*>(s_z + i + 2);
            *reinterpret_cast<float4*>(dp + i) = make_float4(ordered_to_float((uint32_t)(a.x >> 32)), ordered_to_float((uint32_t)(a.y >> 32)),
                                                             ordered_to_float((uint32_t)(b.x >> 32)), ordered_to_float((uint32_t)(b.y >> 32)));
            *reinterpret_cast<uchar4*>(ip + i) = make_uchar4((uint8_t)a.x, (uint8_t)a.y, (uint8_t)b.x, (uint8_t)b.y);
        }
    } else {
        for (int i = threadIdx.x; i < npx; i += kFwdThreads) {
            const unsigned long long kk = s_z[i];
            dp[i] = ordered_to_float((uint32_t)(kk >> 32));
            ip[i] = (uint8_t)kk;
        }
    }
}

template <int PX>
__global__ void __launch_bounds__(kThreads) sphere_render_bwd_kernel(
    const float4* __restrict__ spheres, int J, int H, int W, int tiles_per_block,
    const float* __restrict__ grad_depth, const uint8_t* __restrict__ idx, float4* __restrict__ grad_spheres) {
    __shared__ __align__(128) float4 s_sph[kMaxJ];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) float s_acc[kThreads / 32][kMaxJ * 4];      // one private table per warp
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < (kThreads / 32) * kMaxJ * 4; i += kThreads) (&s_acc[0][0])[i] = 0.f;
    stage_spheres(s_sph, &s_bar, spheres + (size_t)n * J, J);   // contains a __syncthreads()

    const TileGeom g = make_geom(W, H, PX);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = g.tiles_x * g.tiles_y;
    const int t_begin = blockIdx.x * tiles_per_block;
    const int t_end = min(t_begin + tiles_per_block, n_tiles);
    const float halfw = W * 0.5f, halfh = H * 0.5f, fw = (float)W, fh = (float)H;
    float* acc_w = s_acc[warp];

    for (int t = t_begin + warp; t < t_end; t += kThreads / 32) {      // warp-uniform trip count
        const int ty = t / g.tiles_x, tx = t - ty * g.tiles_x;
        const int r = ty * g.trows + lane / g.tq;
        const int c = tx * g.tq * PX + (lane % g.tq) * PX;
        const bool inside = r < H && c < W;
        float gd[PX];
        int ki[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) { gd[p] = 0.f; ki[p] = 255; }
        if (inside) {
            const size_t o = ((size_t)n * H + r) * W + c;
            if (PX == 4) {
                const float4 gv = *reinterpret_cast<const float4*>(grad_depth + o);
                const uchar4 kv = *reinterpret_cast<const uchar4*>(idx + o);
                gd[0] = gv.x; gd[1 % PX] = gv.y; gd[2 % PX] = gv.z; gd[3 % PX] = gv.w;
                ki[0] = kv.x; ki[1 % PX] = kv.y; ki[2 % PX] = kv.z; ki[3 % PX] = kv.w;
            } else {
                gd[0] = grad_depth[o];
                ki[0] = idx[o];
            }
        }
        bool any_fg = false;
#pragma unroll
        for (int p = 0; p < PX; ++p) any_fg |= ki[p] < J;
        if (!__any_sync(0xffffffffu, any_fg)) continue;                // background tile
        const float yg = sh_grid_mm(r, halfh, fh);
        // run-length merge over the lane's pixels, one warp-collective flush per run
        int cur = -1;
        float ax = 0.f, ay = 0.f, az = 0.f, ar = 0.f;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int k = ki[p];
            const bool fg = k < J;
            const bool flush = fg && cur >= 0 && k != cur;             // this lane's run ends here
            if (__any_sync(0xffffffffu, flush)) {
                warp_accumulate(acc_w, flush, cur, ax, ay, az, ar, lane);
                if (flush) { cur = -1; ax = ay = az = ar = 0.f; }
            }
            if (fg) {
                cur = k;
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
        }
        warp_accumulate(acc_w, cur >= 0, cur, ax, ay, az, ar, lane);
    }
    __syncthreads();
    float* out = reinterpret_cast<float*>(grad_spheres + (size_t)n * J);
    for (int i = threadIdx.x; i < J * 4; i += kThreads) {
        float v = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < kThreads / 32; ++w4) v += s_acc[w4][i];
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
    SH_REQUIRE(W <= kZbufBytes / 8, "sh_sphere_render_fwd: W > %d", kZbufBytes / 8);
    if (N == 0) return SH_OK;   // empty batch: nothing to read or write (pointers may be null)
    SH_REQUIRE(spheres && depth && idx, "sh_sphere_render_fwd: null pointer");
    SH_REQUIRE(((uintptr_t)spheres & 15) == 0, "sh_sphere_render_fwd: spheres must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // band of rows per CTA: as many as fit the 32 KB z-buffer, fewer while the grid would leave SMs idle
    int band = kZbufBytes / 8 / W;
    if (band > H) band = H;
    while (band > 8 && (long)N * ((H + band - 1) / band) < 3L * SH_NUM_SMS) band = (band + 1) / 2;
    const size_t smem = (size_t)band * W * 8 + (size_t)(W + band) * 4;
    for (int n0 = 0; n0 < N; n0 += 65535) {
        const int nn = N - n0 < 65535 ? N - n0 : 65535;
        dim3 grid(sh_div_up(H, band), nn);
        sphere_render_fwd_kernel<<<grid, kFwdThreads, smem, st>>>((const float4*)spheres + (size_t)n0 * J, J, H, W, band,
                                                                 (float*)depth + (size_t)n0 * H * W, (uint8_t*)idx + (size_t)n0 * H * W);
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
