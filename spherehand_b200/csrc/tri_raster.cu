// R1 — triangle-mesh z-buffer depth rasteriser (forward only, non-differentiable) (sm_100a).
//
// Replaces the reference's only native code:
//   kernel / depth_rasterization_cuda_forward
//       /root/reference/mesh/cuda_kernel/depth_rasterization_cuda_kernel.cu:18-113, 115-134
// The reference launches <<<B*F, 1>>> (one thread per block, 1/32 lane utilisation) with a CAS-loop float
// atomicMin.  Here 8 lanes share a triangle (4 triangles per warp, columns strided across the lanes), the
// z-test is a single native integer atomic (sign-split ordering of IEEE floats) and the 1000.0 fill is one
// vectorised pass.  The per-column / per-pixel arithmetic reproduces the reference's rounding sequence operation by
// operation (float storage, double-promoted max/min/ceil, truncating int conversion, and the exact FMA fusions of
// the reference binary) so that pixel coverage — an integer decision — and depth are bit-identical to it.
//
// A second entry rasterises only on the sample lattice the reference's 640->S bilinear resize actually
// reads (S=128 -> pixel (5i+2,5j+2); S=64 -> {10i+4,10i+5}^2), which is exact, not an
// approximation, because the resize is applied after clamp(max=100) (mesh/render.py:286,311).
//
// HBM layout: face_vertices fp32 [B,F,3,3]; z-buffer fp32 [B,H,W] (1000.0 where empty).
#include "common.cuh"

namespace {

constexpr int kLanesPerTri = 8;

// z-test: IEEE floats order like sign-magnitude integers.  NaN never wins (fminf semantics of the
// reference's CAS loop, .cu:6-16).
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v != v) return;
    const int bits = __float_as_int(v);
    if (bits >= 0)
        atomicMin(reinterpret_cast<int*>(addr), bits);
    else
        atomicMax(reinterpret_cast<unsigned int*>(addr), (unsigned int)bits);
}

struct Lattice {
    int step;   // 1 = every pixel
    int off0;   // first sampled residue
    int noff;   // consecutive sampled residues per period
    int ow;     // output width  (= W for step 1)
    int oh;
};

// Index of the first lattice sample with coordinate >= x (x >= 0), and coordinate of lattice sample q.
__device__ __forceinline__ int lat_coord(const Lattice& L, int q) { return (q / L.noff) * L.step + L.off0 + (q % L.noff); }
__device__ __forceinline__ int lat_first(const Lattice& L, int x) {
    int c = x / L.step;
    int q = c * L.noff;
    while (lat_coord(L, q) < x) ++q;
    return q;
}

// Per-triangle constants.  Every floating-point operation below is an explicit round-to-nearest intrinsic in the
// exact sequence (which products are fused into an FMA, which are rounded first) that the reference binary executes
// — read off `cuobjdump -sass` of the reference kernel built by oracle/build_ref.py (nvcc 12.9, default -fmad=true) —
// so coverage AND depth are bit-identical to it, for the full-image and the lattice variant alike, whatever this
// translation unit's own contraction choices would have been.
struct TriSetup {
    float ax, ay, bx, by, cx;
    float az, bz, cz;
    float e[9];                 // rows of the inverse edge-function matrix (.cu:57-65)
    float s_lo, s_hi, s_long;   // slopes of edges lo-mid, mid-hi, lo-hi
    bool lo_vertical, hi_vertical;
};

__device__ __forceinline__ bool tri_setup(const float* __restrict__ f, TriSetup& t) {
    // back-face cull on the raw winding (.cu:33)
    if (__fmul_rn(__fsub_rn(f[7], f[1]), __fsub_rn(f[3], f[0])) < __fmul_rn(__fsub_rn(f[4], f[1]), __fsub_rn(f[6], f[0])))
        return false;
    // vertex order by x: lo / mid / hi, ties resolved as the reference does (.cu:36-45)
    int lo, hi;
    if (f[0] < f[3]) {
        lo = (f[6] < f[0]) ? 2 : 0;
        hi = (f[3] < f[6]) ? 2 : 1;
    } else {
        lo = (f[6] < f[3]) ? 2 : 1;
        hi = (f[0] < f[6]) ? 2 : 0;
    }
    const int mid = 3 - lo - hi;
    const float ax = f[3 * lo], ay = f[3 * lo + 1];
    const float bx = f[3 * mid], by = f[3 * mid + 1];
    const float cx = f[3 * hi], cy = f[3 * hi + 1];
    if (ax == cx) return false;   // zero width (.cu:54)
    t.ax = ax; t.ay = ay; t.bx = bx; t.by = by; t.cx = cx;
    t.az = f[3 * lo + 2]; t.bz = f[3 * mid + 2]; t.cz = f[3 * hi + 2];
    const float by_cy = __fsub_rn(by, cy), cy_ay = __fsub_rn(cy, ay), ay_by = __fsub_rn(ay, by);
    const float cx_bx = __fsub_rn(cx, bx), ax_cx = __fsub_rn(ax, cx), bx_ax = __fsub_rn(bx, ax);
    // den = cx(ay-by) + ax(by-cy) + bx(cy-ay): middle product rounded, the outer two fused (.cu:61-64)
    const float den = __fmaf_rn(bx, cy_ay, __fmaf_rn(cx, ay_by, __fmul_rn(ax, by_cy)));
    t.e[0] = __fdiv_rn(by_cy, den);
    t.e[1] = __fdiv_rn(cx_bx, den);
    t.e[2] = __fdiv_rn(__fmaf_rn(bx, cy, -__fmul_rn(cx, by)), den);
    t.e[3] = __fdiv_rn(cy_ay, den);
    t.e[4] = __fdiv_rn(ax_cx, den);
    t.e[5] = __fdiv_rn(__fmaf_rn(cx, ay, -__fmul_rn(ax, cy)), den);
    t.e[6] = __fdiv_rn(ay_by, den);
    t.e[7] = __fdiv_rn(bx_ax, den);
    t.e[8] = __fdiv_rn(__fmaf_rn(ax, by, -__fmul_rn(bx, ay)), den);
    t.lo_vertical = !(bx_ax != 0.f);
    t.hi_vertical = !(cx_bx != 0.f);
    t.s_lo = __fdiv_rn(__fsub_rn(by, ay), bx_ax);
    t.s_hi = __fdiv_rn(__fsub_rn(cy, by), cx_bx);
    t.s_long = __fdiv_rn(cy_ay, __fsub_rn(cx, ax));
    return true;
}

// Row span [yi_min, yi_max] of column xi (.cu:72-90): bounds are taken in double and truncated, as the reference does.
__device__ __forceinline__ void tri_column_span(const TriSetup& t, int32_t xi, int height, int32_t& yi_min, int32_t& yi_max) {
    const float xf = (float)xi;
    float y_edge;
    if (xf <= t.bx)
        y_edge = t.lo_vertical ? t.by : __fmaf_rn(__fsub_rn(xf, t.ax), t.s_lo, t.ay);
    else
        y_edge = t.hi_vertical ? t.by : __fmaf_rn(__fsub_rn(xf, t.bx), t.s_hi, t.by);
    const float y_long = __fmaf_rn(__fsub_rn(xf, t.ax), t.s_long, t.ay);
    yi_min = max(0., ceil(min(y_edge, y_long)));
    yi_max = min(max(y_edge, y_long), height - 1.);
}

// Depth of the triangle's plane at pixel (xi, yi): clamped + renormalised barycentrics, 1/z blend (.cu:97-109).
__device__ __forceinline__ float tri_pixel_depth(const TriSetup& t, int32_t xi, int32_t yi) {
    const float xf = (float)xi, yf = (float)yi;
    // the reference binary fuses row 0 on the x product and rows 1, 2 on the y product (its x products are hoisted)
    float w0 = __fadd_rn(__fmaf_rn(t.e[0], xf, __fmul_rn(t.e[1], yf)), t.e[2]);
    float w1 = __fadd_rn(__fmaf_rn(yf, t.e[4], __fmul_rn(xf, t.e[3])), t.e[5]);
    float w2 = __fadd_rn(__fmaf_rn(yf, t.e[7], __fmul_rn(xf, t.e[6])), t.e[8]);
    w0 = min(max(w0, 0.), 1.);
    w1 = min(max(w1, 0.), 1.);
    w2 = min(max(w2, 0.), 1.);
    const float total = __fadd_rn(__fadd_rn(__fadd_rn(0.f, w0), w1), w2);
    w0 = __fdiv_rn(w0, total);
    w1 = __fdiv_rn(w1, total);
    w2 = __fdiv_rn(w2, total);
    const float q = __fadd_rn(__fadd_rn(__fdiv_rn(w0, t.az), __fdiv_rn(w1, t.bz)), __fdiv_rn(w2, t.cz));
    return __frcp_rn(q);   // == (float)(1. / (double)q): the double quotient rounds to the correctly rounded float
}

template <bool LATTICE>
__global__ void __launch_bounds__(128) tri_raster_kernel(const int num_faces, const long total_faces, const int width,
                                                         const int height, const float* __restrict__ face_vertices,
                                                         float* __restrict__ depth_map, const Lattice L) {
    const long gid = (long)blockIdx.x * (blockDim.x / kLanesPerTri) + threadIdx.x / kLanesPerTri;
    const int sub = threadIdx.x % kLanesPerTri;
    if (gid >= total_faces) return;
    const long image = gid / num_faces;
    TriSetup t;
    if (!tri_setup(&face_vertices[gid * 9], t)) return;

    const int32_t xi_min = max(ceil(t.ax), 0.);        // float -> double -> truncate (.cu:68)
    const int32_t xi_max = min(t.cx, width - 1.);      // (.cu:69)
    float* out = depth_map + image * (LATTICE ? (long)L.ow * L.oh : (long)width * height);

    // the 8 lanes of a triangle take columns round-robin
    for (int q = (LATTICE ? lat_first(L, xi_min) : xi_min) + sub;; q += kLanesPerTri) {
        const int32_t xi = LATTICE ? lat_coord(L, q) : q;
        if (xi > xi_max) break;
        int32_t yi_min, yi_max;
        tri_column_span(t, xi, height, yi_min, yi_max);
        for (int r = LATTICE ? lat_first(L, yi_min) : yi_min;; ++r) {
            const int32_t yi = LATTICE ? lat_coord(L, r) : r;
            if (yi > yi_max) break;
            const float zp = tri_pixel_depth(t, xi, yi);
            atomic_min_float(&out[LATTICE ? (long)r * L.ow + q : (long)yi * width + xi], zp);
        }
    }
}

__global__ void fill_kernel(float4* __restrict__ p4, long n4, float* __restrict__ tail, int ntail, float v) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    const float4 vv = make_float4(v, v, v, v);
    for (long k = i; k < n4; k += stride) p4[k] = vv;
    if (i < ntail) tail[i] = v;
}

int fill_value(float* p, long n, float v, cudaStream_t st) {
    // p is at least 4-byte aligned; peel to 16 B
    long head = 0;
    while (head < n && (((uintptr_t)(p + head)) & 15)) ++head;
    const long n4 = (n - head) / 4;
    const long tail0 = head + n4 * 4;
    if (head) fill_kernel<<<1, 32, 0, st>>>(nullptr, 0, p, (int)head, v);
    const int blocks = (int)((n4 + 255) / 256 < 148L * 16 ? (n4 + 255) / 256 : 148L * 16);
    fill_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>((float4*)(p + head), n4, p + tail0, (int)(n - tail0), v);
    return SH_OK;
}

}  // namespace

// Drop-in for the pybind entry `depth_rasterization.forward(width, height, vertices)`
// (/root/reference/mesh/cuda_kernel/depth_rasterization_cuda.cpp:15-25): fills `out` with 1000.0 and rasterises.
SH_EXPORT int sh_tri_raster_fwd(const void* face_vertices, int B, int F, int W, int H, void* out, void* stream) {
    SH_REQUIRE(out && (face_vertices || (long)B * F == 0), "sh_tri_raster_fwd: null pointer");
    SH_REQUIRE(B >= 0 && F >= 0 && W >= 1 && H >= 1, "sh_tri_raster_fwd: bad B/F/W/H");
    if (B == 0) return SH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    fill_value((float*)out, (long)B * H * W, 1000.0f, st);
    SH_CHECK_LAUNCH("fill_kernel");
    const long total = (long)B * F;
    if (total > 0) {
        Lattice L{1, 0, 1, W, H};
        const int tris_per_block = 128 / kLanesPerTri;
        tri_raster_kernel<false><<<sh_div_up(total, tris_per_block), 128, 0, st>>>(F, total, W, H, (const float*)face_vertices,
                                                                               (float*)out, L);
        SH_CHECK_LAUNCH("tri_raster_kernel");
    }
    return SH_OK;
}

// Rasterise only the pixels {c*step + off0 + o : o < noff}^2 of a virtual W x H image into a compact
// [B, oh, ow] buffer (oh = ow = (number of lattice samples below H / W)).
SH_EXPORT int sh_tri_raster_lattice_fwd(const void* face_vertices, int B, int F, int W, int H, int step, int off0,
                                         int noff, void* out, int oh, int ow, void* stream) {
    SH_REQUIRE(out && face_vertices, "sh_tri_raster_lattice_fwd: null pointer");
    SH_REQUIRE(B >= 0 && F >= 0 && W >= 1 && H >= 1 && step >= 1 && noff >= 1 && off0 >= 0 && off0 + noff <= step,
               "sh_tri_raster_lattice_fwd: bad arguments");
    SH_REQUIRE(ow == (W / step) * noff && oh == (H / step) * noff && W % step == 0 && H % step == 0,
               "sh_tri_raster_lattice_fwd: output size does not match the lattice");
    if (B == 0) return SH_OK;
    cudaStream_t st = (cudaStream_t)stream;
    fill_value((float*)out, (long)B * oh * ow, 1000.0f, st);
    SH_CHECK_LAUNCH("fill_kernel");
    const long total = (long)B * F;
    if (total > 0) {
        Lattice L{step, off0, noff, ow, oh};
        const int tris_per_block = 128 / kLanesPerTri;
        tri_raster_kernel<true><<<sh_div_up(total, tris_per_block), 128, 0, st>>>(F, total, W, H, (const float*)face_vertices,
                                                                              (float*)out, L);
        SH_CHECK_LAUNCH("tri_raster_kernel<lattice>");
    }
    return SH_OK;
}
