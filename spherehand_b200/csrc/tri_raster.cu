// R1 — triangle-mesh z-buffer depth rasteriser (forward only, non-differentiable) (sm_100a).
//
// Replaces the reference's only native code:
//   kernel / depth_rasterization_cuda_forward
//       /root/reference/mesh/cuda_kernel/depth_rasterization_cuda_kernel.cu:18-113, 115-134
// The reference launches <<<B*F, 1>>> (one thread per block, 1/32 lane utilisation) with a CAS-loop float
// atomicMin.  Here 8 lanes share a triangle (4 triangles per warp, columns strided across the lanes), the
// z-test is a single native integer atomic (sign-split ordering of IEEE floats) and the 1000.0 fill is one
// vectorised pass.  The per-column / per-pixel expressions are written exactly as in the reference
// (float storage, double-promoted max/min/ceil, truncating int conversion, 1./(...) in double) so that pixel
// coverage — an integer decision — is taken on identical values.
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

template <bool LATTICE>
__global__ void __launch_bounds__(128) tri_raster_kernel(const int num_faces, const long total_faces, const int width,
                                                         const int height, const float* __restrict__ face_vertices,
                                                         float* __restrict__ depth_map, const Lattice L) {
    const long gid = (long)blockIdx.x * (blockDim.x / kLanesPerTri) + threadIdx.x / kLanesPerTri;
    const int sub = threadIdx.x % kLanesPerTri;
    if (gid >= total_faces) return;
    const long image = gid / num_faces;
    const float* f = &face_vertices[gid * 9];

    // back-face cull on the raw winding (.cu:33)
    if ((f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0])) return;

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
    const float ax = f[3 * lo], ay = f[3 * lo + 1], az = f[3 * lo + 2];
    const float bx = f[3 * mid], by = f[3 * mid + 1], bz = f[3 * mid + 2];
    const float cx = f[3 * hi], cy = f[3 * hi + 1], cz = f[3 * hi + 2];
    if (ax == cx) return;   // zero width (.cu:54)

    // rows of the inverse edge-function matrix (.cu:57-65); expression shapes kept so nvcc contracts the same FMAs
    float e[9] = {by - cy, cx - bx, bx * cy - cx * by,
                  cy - ay, ax - cx, cx * ay - ax * cy,
                  ay - by, bx - ax, ax * by - bx * ay};
    float e_den = (cx * (ay - by) + ax * (by - cy) + bx * (cy - ay));
#pragma unroll
    for (int k = 0; k < 9; k++) e[k] /= e_den;

    const int32_t xi_min = max(ceil(ax), 0.);          // float -> double -> truncate (.cu:68)
    const int32_t xi_max = min(cx, width - 1.);        // (.cu:69)
    float* out = depth_map + image * (LATTICE ? (long)L.ow * L.oh : (long)width * height);

    int q = LATTICE ? lat_first(L, xi_min) + sub : xi_min + sub;
    for (;; q += kLanesPerTri) {
        const int32_t xi = LATTICE ? lat_coord(L, q) : q;
        if (xi > xi_max) break;
        float y_edge, y_long;
        if (xi <= bx) {                                 // left part: edge lo-mid (.cu:73-79)
            if (bx - ax != 0) {
                y_edge = (by - ay) / (bx - ax) * (xi - ax) + ay;
            } else {
                y_edge = by;
            }
        } else {                                        // right part: edge mid-hi (.cu:80-86)
            if (cx - bx != 0) {
                y_edge = (cy - by) / (cx - bx) * (xi - bx) + by;
            } else {
                y_edge = by;
            }
        }
        y_long = (cy - ay) / (cx - ax) * (xi - ax) + ay; // long edge lo-hi (.cu:87)

        const int32_t yi_min = max(0., ceil(min(y_edge, y_long)));   // (.cu:89)
        const int32_t yi_max = min(max(y_edge, y_long), height - 1.); // (.cu:90)
        int r = LATTICE ? lat_first(L, yi_min) : yi_min;
        for (;; ++r) {
            const int32_t yi = LATTICE ? lat_coord(L, r) : r;
            if (yi > yi_max) break;
            float bary[3];
#pragma unroll
            for (int k = 0; k < 3; k++) bary[k] = e[3 * k + 0] * xi + e[3 * k + 1] * yi + e[3 * k + 2];
            float total = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                bary[k] = min(max(bary[k], 0.), 1.);    // clamp in double, store float (.cu:103)
                total += bary[k];
            }
#pragma unroll
            for (int k = 0; k < 3; k++) bary[k] /= total;
            const float zp = 1. / (bary[0] / az + bary[1] / bz + bary[2] / cz);   // perspective-style 1/z blend (.cu:109)
            const long index = LATTICE ? (long)r * L.ow + q : (long)yi * width + xi;
            atomic_min_float(&out[index], zp);
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
