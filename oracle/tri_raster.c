/* CPU oracle for R1, the triangle z-buffer rasteriser.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A restatement (not a copy) of the algorithm in
 *   /root/reference/mesh/cuda_kernel/depth_rasterization_cuda_kernel.cu:18-113  (per-triangle kernel)
 *   /root/reference/mesh/cuda_kernel/depth_rasterization_cuda_kernel.cu:115-134 (z-buffer init 1000.0)
 * including the parts that decide pixel coverage: float storage with double-promoted
 * max/min/ceil sub-expressions (:68-69, :89-90, :103), truncating int32 conversion of the bounds,
 * the vertex sort tie rules (:38-45), the degenerate-triangle early-out (:54), the vertical-edge special
 * cases (:74-84), clamp-then-renormalise of the barycentrics (:101-106) and the double-precision
 * reciprocal of the 1/z blend (:109).  Min-combination is order independent, so a serial loop is exact.
 *
 * `use_fma` selects how a*b+c sub-expressions are rounded: 0 = every operation rounded separately
 * (ISO C), 1 = exactly the fused multiply-adds that nvcc 12.9's default contraction (-fmad=true) emitted for the
 * reference binary (read off `cuobjdump -sass oracle/_ref/depth_rasterization_ref.so`; with it this oracle is
 * bit-identical to the reference kernel run on the GPU box — tests/test_gpu_kernels.py checks all three ways).
 * Additionally records the winning face per pixel (the "integer z-buffer argmin index" of
 * BASELINE.json, which the reference itself never materialises).
 */
#include <math.h>
#include <stdint.h>

static float mad(float a, float b, float c, int use_fma) { return use_fma ? fmaf(a, b, c) : a * b + c; }
/* a*b - c*d as nvcc contracts it: the second product rounded, the first fused */
static float msub2(float a, float b, float c, float d, int use_fma) {
    float t = c * d;
    return use_fma ? fmaf(a, b, -t) : a * b - t;
}

static void raster_one(const float *f, int width, int height, float *zbuf, int32_t *fbuf, int32_t fid, int use_fma) {
    /* back-face cull, .cu:33 */
    if ((f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0])) return;

    /* order the vertices by x: lo, mid, hi with the reference's tie rules, .cu:36-45 */
    int lo, hi, mid = 0;
    if (f[0] < f[3]) {
        lo = (f[6] < f[0]) ? 2 : 0;
        hi = (f[3] < f[6]) ? 2 : 1;
    } else {
        lo = (f[6] < f[3]) ? 2 : 1;
        hi = (f[0] < f[6]) ? 2 : 0;
    }
    for (int k = 0; k < 3; k++)
        if (k != lo && k != hi) mid = k;
    const float ax = f[3 * lo], ay = f[3 * lo + 1], az = f[3 * lo + 2];
    const float bx = f[3 * mid], by = f[3 * mid + 1], bz = f[3 * mid + 2];
    const float cx = f[3 * hi], cy = f[3 * hi + 1], cz = f[3 * hi + 2];
    if (ax == cx) return; /* .cu:54 */

    /* inverse of the edge-function matrix, .cu:57-65 */
    float inv[9];
    inv[0] = by - cy;  inv[1] = cx - bx;  inv[2] = msub2(bx, cy, cx, by, use_fma);
    inv[3] = cy - ay;  inv[4] = ax - cx;  inv[5] = msub2(cx, ay, ax, cy, use_fma);
    inv[6] = ay - by;  inv[7] = bx - ax;  inv[8] = msub2(ax, by, bx, ay, use_fma);
    float den;
    if (use_fma) {
        den = fmaf(bx, cy - ay, fmaf(cx, ay - by, ax * (by - cy)));
    } else {
        den = cx * (ay - by) + ax * (by - cy) + bx * (cy - ay);
    }
    for (int k = 0; k < 9; k++) inv[k] /= den;

    const int32_t xi_min = (int32_t)fmax((double)ceilf(ax), 0.);          /* .cu:68 */
    const int32_t xi_max = (int32_t)fmin((double)cx, width - 1.);         /* .cu:69 */
    for (int32_t xi = xi_min; xi <= xi_max; xi++) {
        const float xf = (float)xi;
        float y1, y2;
        if (xf <= bx) {                                                   /* .cu:73-79 */
            y1 = (bx - ax != 0) ? mad((by - ay) / (bx - ax), xf - ax, ay, use_fma) : by;
        } else {                                                          /* .cu:80-86 */
            y1 = (cx - bx != 0) ? mad((cy - by) / (cx - bx), xf - bx, by, use_fma) : by;
        }
        y2 = mad((cy - ay) / (cx - ax), xf - ax, ay, use_fma);            /* .cu:87 */

        const int32_t yi_min = (int32_t)fmax(0., (double)ceilf(fminf(y1, y2)));      /* .cu:89 */
        const int32_t yi_max = (int32_t)fmin((double)fmaxf(y1, y2), height - 1.);    /* .cu:90 */
        for (int32_t yi = yi_min; yi <= yi_max; yi++) {
            const float yf = (float)yi;
            float w[3], wsum = 0;
            for (int k = 0; k < 3; k++) {
                float v;                                                                       /* .cu:99 */
                if (!use_fma)
                    v = inv[3 * k] * xf + inv[3 * k + 1] * yf + inv[3 * k + 2];
                else if (k == 0) /* reference binary: row 0 fuses the x product ... */
                    v = fmaf(inv[0], xf, inv[1] * yf) + inv[2];
                else             /* ... rows 1, 2 the y product (their x products are hoisted out of the row loop) */
                    v = fmaf(yf, inv[3 * k + 1], xf * inv[3 * k]) + inv[3 * k + 2];
                v = (float)fmin(fmax((double)v, 0.), 1.);                                    /* .cu:103 */
                w[k] = v;
                wsum += v;
            }
            for (int k = 0; k < 3; k++) w[k] /= wsum;
            const float s = w[0] / az + w[1] / bz + w[2] / cz;
            const float zp = (float)(1. / (double)s);                                         /* .cu:109 */
            const long idx = (long)yi * width + xi;
            if (fminf(zp, zbuf[idx]) != zbuf[idx]) {   /* fminf ignores a NaN zp, like the CAS loop :6-16 */
                zbuf[idx] = zp;
                fbuf[idx] = fid;
            }
        }
    }
}

void oracle_tri_raster(const float *face_vertices, int num_batches, int num_faces, int width, int height,
                       float *depth, int32_t *face_id, int use_fma) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < num_batches; b++) {
        float *z = depth + (long)b * width * height;
        int32_t *fb = face_id + (long)b * width * height;
        for (long i = 0; i < (long)width * height; i++) {
            z[i] = 1000.0f;   /* .cu:122 */
            fb[i] = -1;
        }
        for (int fi = 0; fi < num_faces; fi++)
            raster_one(face_vertices + ((long)b * num_faces + fi) * 9, width, height, z, fb, fi, use_fma);
    }
}
