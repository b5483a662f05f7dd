// Synthetic-data branch: forward kinematics, skinning, projection, heat-map targets, depth noise (sm_100a).
//
// Replaces (all forward-only; the reference runs them detached, network/util_modules.py:122)
//   HandTransformationMat.forward (+ AxisRotationMatrix, TranslationMatrix, FingerJoint, Finger, Palm)
//       /root/reference/mesh/kinematicsTransformation.py:29-54, 61-68, 92-112, 123-127, 145-155, 169-177
//   RandScale.forward                  /root/reference/mesh/pointTransformation.py:135-148
//   LinearBlendSkinning.forward        /root/reference/mesh/pointTransformation.py:39-46
//   OthographicalProjection.forward    /root/reference/mesh/pointTransformation.py:84-99
//   InverseOthographicalProjection     /root/reference/mesh/pointTransformation.py:118-124
//   HeatmapRender.forward              /root/reference/mesh/render.py:226-248
//   DepthNoise.forward                 /root/reference/network/util_modules.py:60-84
//   the clamp/resize/scale tail of DepthRasterization/HandSynthesizer
//       /root/reference/mesh/render.py:286,311 ; network/util_modules.py:112
// The reference spends several hundred tiny launches on the 4x4 chains; here FK is one launch (one thread per
// (pose, finger)), skinning visits only the non-zero bone weights (CSR, <= 5 per vertex instead of 17).
//
// HBM layout: params fp32 [B,26]; mats fp32 [B,17,4,4] row-major; skin CSR: row_ptr int32 [Nv+1], bone int32
// [nnz], wv float4 [nnz] (= w * rest vertex, fp32 product as pointTransformation.py:30); points float4 [B,Nv].
#include "common.cuh"

namespace {

struct M4 { float m[16]; };

__device__ __forceinline__ M4 mul44(const M4& a, const M4& b) {
    M4 c;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += a.m[i * 4 + k] * b.m[k * 4 + j];
            c.m[i * 4 + j] = s;
        }
    return c;
}

// Rodrigues rotation about a fixed unit axis (kinematicsTransformation.py:39-53)
__device__ __forceinline__ M4 axis_rot(float x, float y, float z, float ang) {
    float s, c;
    sincosf(ang, &s, &c);
    const float i = 1.f - c;
    M4 r;
    r.m[0] = x * x * i + c;     r.m[1] = x * y * i - z * s; r.m[2] = x * z * i + y * s;  r.m[3] = 0.f;
    r.m[4] = x * y * i + z * s; r.m[5] = y * y * i + c;     r.m[6] = y * z * i - x * s;  r.m[7] = 0.f;
    r.m[8] = x * z * i - y * s; r.m[9] = y * z * i + x * s; r.m[10] = z * z * i + c;     r.m[11] = 0.f;
    r.m[12] = 0.f; r.m[13] = 0.f; r.m[14] = 0.f; r.m[15] = 1.f;
    return r;
}

__device__ __forceinline__ M4 load44(const float* p) {
    M4 r;
#pragma unroll
    for (int i = 0; i < 16; ++i) r.m[i] = p[i];
    return r;
}

__device__ __forceinline__ void store_scaled(float* dst, const M4& m, const float* sc) {
    // RandScale: diag(sx,sy,sz,1) @ M  ==  scale rows 0..2 (pointTransformation.py:144-148)
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = (sc && i < 12) ? sc[i / 4] * m.m[i] : m.m[i];
}

__global__ void fk_kernel(const float* __restrict__ params, const float* __restrict__ scales,
                          const float* __restrict__ off, const float* __restrict__ inv_off, int B,
                          float* __restrict__ mats) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = t / 5, f = t % 5;
    if (b >= B) return;
    const float* p = params + (size_t)b * 26;
    const float* sc = scales ? scales + (size_t)b * 3 : nullptr;
    // palm = T * Rz * Ry * Rx  (kinematicsTransformation.py:148-152)
    M4 rot = axis_rot(1.f, 0.f, 0.f, p[0]);
    rot = mul44(axis_rot(0.f, 1.f, 0.f, p[1]), rot);
    rot = mul44(axis_rot(0.f, 0.f, 1.f, p[2]), rot);
    M4 tr;
#pragma unroll
    for (int i = 0; i < 16; ++i) tr.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    tr.m[3] = p[3]; tr.m[7] = p[4]; tr.m[11] = p[5];
    const M4 palm = mul44(tr, rot);
    float* out = mats + (size_t)b * 17 * 16;
    if (f == 0) {
        store_scaled(out, palm, sc);
        store_scaled(out + 16, palm, sc);      // carpals share the palm transform (:154)
    }
    // abduction axes z, z, -y, -y, z for bone groups finger4,3,2,1,5 (:162-166)
    const float axy = (f == 2 || f == 3) ? -1.f : 0.f;
    const float axz = (f == 2 || f == 3) ? 0.f : 1.f;
    const float* a = p + 6 + 4 * f;
    M4 parent = palm;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int bone = 2 + 3 * f + k;
        M4 local;
        if (k == 0) local = mul44(axis_rot(0.f, axy, axz, a[0]), axis_rot(1.f, 0.f, 0.f, a[1]));
        else local = axis_rot(1.f, 0.f, 0.f, a[k + 1]);
        M4 g = mul44(mul44(load44(inv_off + bone * 16), local), load44(off + bone * 16));   // O^-1 * R * O (:108-109)
        parent = mul44(parent, g);
        store_scaled(out + bone * 16, parent, sc);
    }
}

// mode 0: skinned homogeneous points (LinearBlendSkinning.forward)
// mode 1: + projection with per-sample focal jitter: (x*rand_f*fx + cx, y*rand_f*fy + cy, z, 1)   (:90-97)
// mode 2: + projection with the K matrix: (fx*x + cx*w, fy*y + cy*w, z, w)                        (:89)
__global__ void lbs_kernel(const float* __restrict__ mats, const int* __restrict__ row_ptr, const int* __restrict__ bone,
                           const float4* __restrict__ wv, int B, int Nv, int right_hand, int mode, float cx, float cy,
                           float fx, float fy, const float* __restrict__ rand_f, float4* __restrict__ out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * Nv) return;
    const int b = (int)(t / Nv), v = (int)(t % Nv);
    const float* M = mats + (size_t)b * 17 * 16;
    float x = 0.f, y = 0.f, z = 0.f, w = 0.f;
    for (int e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
        const float* T = M + bone[e] * 16;
        const float4 p = wv[e];
        x += T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3] * p.w;
        y += T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7] * p.w;
        z += T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11] * p.w;
        w += T[12] * p.x + T[13] * p.y + T[14] * p.z + T[15] * p.w;
    }
    if (right_hand) x = -x;
    if (mode == 1) {
        const float f = rand_f[b];
        x = x * f * fx + cx;
        y = y * f * fy + cy;
        w = 1.f;
    } else if (mode == 2) {
        x = fx * x + cx * w;
        y = fy * y + cy * w;
    }
    out[t] = make_float4(x, y, z, w);
}

// face_vertices[b,f,3,3] = points[b, faces[f,:], 0:3]   (mesh/render.py:308-309)
__global__ void gather_faces_kernel(const float4* __restrict__ pts, const int* __restrict__ faces, int B, int Nv, int F,
                                    float* __restrict__ fv) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * F * 3) return;
    const int b = (int)(t / (F * 3)), fi = (int)(t % (F * 3));
    const float4 p = pts[(size_t)b * Nv + faces[fi]];
    float* o = fv + t * 3;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
}

// z lattice [B, S*noff, S*noff] -> depth image [B,S,S]: clamp(max=100) (render.py:286), mean of the noff^2
// samples the bilinear resize reads (render.py:311), * depth_scale (util_modules.py:112)
__global__ void lattice_to_depth_kernel(const float* __restrict__ z, int B, int S, int noff, float depth_scale,
                                        float* __restrict__ dm) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * S * S) return;
    const int x = (int)(t % S), y = (int)((t / S) % S);
    const long b = t / ((long)S * S);
    const int ow = S * noff;
    const float* zi = z + b * ow * ow;
    float v;
    if (noff == 1) {
        v = fminf(zi[(long)y * ow + x], 100.f);
    } else {
        const float a = fminf(zi[(long)(2 * y) * ow + 2 * x], 100.f), c = fminf(zi[(long)(2 * y) * ow + 2 * x + 1], 100.f);
        const float d = fminf(zi[(long)(2 * y + 1) * ow + 2 * x], 100.f), e = fminf(zi[(long)(2 * y + 1) * ow + 2 * x + 1], 100.f);
        // F.interpolate bilinear: horizontal lerp of each row (weights .5/.5), then vertical
        v = 0.5f * (0.5f * a + 0.5f * c) + 0.5f * (0.5f * d + 0.5f * e);
    }
    dm[t] = v * depth_scale;
}

// DepthNoise with the three N(0,1) draws supplied by the host (util_modules.py:64-83)
__global__ void depth_noise_kernel(const float* __restrict__ dm, const float* __restrict__ nx, const float* __restrict__ ny,
                                   const float* __restrict__ nz, int B, int H, int W, float sx, float sy, float sz,
                                   float* __restrict__ out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * H * W) return;
    const int x = (int)(t % W), y = (int)((t / W) % H);
    const long b = t / ((long)H * W);
    int u = (int)(nx[t] * sx + 0.5f) + x;     // .long() truncates toward zero
    int v = (int)(ny[t] * sy + 0.5f) + y;
    u = min(max(u, 0), W - 1);
    v = min(max(v, 0), H - 1);
    float d = dm[b * H * W + (long)v * W + u];
    if (d < 1.0f) d += nz[t] * sz;
    out[t] = d;
}

// uvd float4 [B,J] -> uv_hm, d_hm [B,J,h,w] and xyz float4 [B,J] (inverse K)
__global__ void heatmap_kernel(const float4* __restrict__ uvd, int B, int J, int hm, float sigma, float uv_scale,
                               float depth_scale, float cx, float cy, float fx, float fy, float* __restrict__ uv_hms,
                               float* __restrict__ d_hms, float4* __restrict__ xyz) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * J * hm * hm;
    if (t >= total) return;
    const int x = (int)(t % hm), y = (int)((t / hm) % hm);
    const long bj = t / ((long)hm * hm);
    const float4 p = uvd[bj];
    const float du = (float)x - p.x, dv = (float)y - p.y;
    const float g = expf(-0.5f * sigma * (du * du + dv * dv));
    uv_hms[t] = g * uv_scale;
    d_hms[t] = (g > 0.05f ? p.z : 0.f) * depth_scale;
    if (x == 0 && y == 0) xyz[bj] = make_float4((p.x - cx * p.w) / fx, (p.y - cy * p.w) / fy, p.z, p.w);
}

}  // namespace

SH_EXPORT int sh_fk_fwd(const void* params, const void* scales, const void* offset_mats, const void* inv_offset_mats,
                         int B, void* mats, void* stream) {
    SH_REQUIRE(params && offset_mats && inv_offset_mats && mats, "sh_fk_fwd: null pointer");
    SH_REQUIRE(B >= 0, "sh_fk_fwd: bad B");
    if (B == 0) return SH_OK;
    fk_kernel<<<sh_div_up((long)B * 5, 128), 128, 0, (cudaStream_t)stream>>>((const float*)params, (const float*)scales,
                                                                            (const float*)offset_mats,
                                                                            (const float*)inv_offset_mats, B, (float*)mats);
    SH_CHECK_LAUNCH("fk_kernel");
    return SH_OK;
}

SH_EXPORT int sh_lbs_fwd(const void* mats, const void* row_ptr, const void* bone, const void* wv, int B, int Nv,
                          int right_hand, int mode, float cx, float cy, float fx, float fy, const void* rand_f,
                          void* out_points, void* stream) {
    SH_REQUIRE(mats && row_ptr && bone && wv && out_points, "sh_lbs_fwd: null pointer");
    SH_REQUIRE(B >= 0 && Nv >= 1 && mode >= 0 && mode <= 2 && (mode != 1 || rand_f), "sh_lbs_fwd: bad arguments");
    if (B == 0) return SH_OK;
    lbs_kernel<<<sh_div_up((long)B * Nv, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)mats, (const int*)row_ptr, (const int*)bone, (const float4*)wv, B, Nv, right_hand, mode, cx, cy, fx,
        fy, (const float*)rand_f, (float4*)out_points);
    SH_CHECK_LAUNCH("lbs_kernel");
    return SH_OK;
}

SH_EXPORT int sh_gather_faces(const void* points, const void* faces, int B, int Nv, int F, void* face_vertices,
                               void* stream) {
    SH_REQUIRE(points && faces && face_vertices, "sh_gather_faces: null pointer");
    if (B == 0 || F == 0) return SH_OK;
    gather_faces_kernel<<<sh_div_up((long)B * F * 3, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)points, (const int*)faces, B, Nv, F, (float*)face_vertices);
    SH_CHECK_LAUNCH("gather_faces_kernel");
    return SH_OK;
}

SH_EXPORT int sh_lattice_to_depth(const void* z, int B, int S, int noff, float depth_scale, void* dm, void* stream) {
    SH_REQUIRE(z && dm && (noff == 1 || noff == 2), "sh_lattice_to_depth: bad arguments");
    if (B == 0) return SH_OK;
    lattice_to_depth_kernel<<<sh_div_up((long)B * S * S, 256), 256, 0, (cudaStream_t)stream>>>((const float*)z, B, S, noff,
                                                                                            depth_scale, (float*)dm);
    SH_CHECK_LAUNCH("lattice_to_depth_kernel");
    return SH_OK;
}

SH_EXPORT int sh_depth_noise(const void* dm, const void* nx, const void* ny, const void* nz, int B, int H, int W,
                              float sx, float sy, float sz, void* out, void* stream) {
    SH_REQUIRE(dm && nx && ny && nz && out && dm != out, "sh_depth_noise: bad arguments");
    if (B == 0) return SH_OK;
    depth_noise_kernel<<<sh_div_up((long)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)dm, (const float*)nx, (const float*)ny, (const float*)nz, B, H, W, sx, sy, sz, (float*)out);
    SH_CHECK_LAUNCH("depth_noise_kernel");
    return SH_OK;
}

SH_EXPORT int sh_heatmap_render(const void* uvd, int B, int J, int hm, float sigma, float uv_scale, float depth_scale,
                                 float cx, float cy, float fx, float fy, void* uv_hms, void* d_hms, void* xyz,
                                 void* stream) {
    SH_REQUIRE(uvd && uv_hms && d_hms && xyz, "sh_heatmap_render: null pointer");
    if (B == 0) return SH_OK;
    heatmap_kernel<<<sh_div_up((long)B * J * hm * hm, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)uvd, B, J, hm, sigma, uv_scale, depth_scale, cx, cy, fx, fy, (float*)uv_hms, (float*)d_hms,
        (float4*)xyz);
    SH_CHECK_LAUNCH("heatmap_kernel");
    return SH_OK;
}

// ------------------------------------------------------------------------------------------------ pose sampler
// JointAngleDataset.__getitem__ (/root/reference/dataset/joint_angle.py:21-233) for n poses in one launch.  The reference
// draws one pose with ~35 separate `torch.rand(1)` calls and as many tiny tensor ops; here a thread owns a pose and consumes
// ITS slice of a pre-drawn uniform stream in exactly the reference's order (palm 6, spread + 4 abductions, thumb 4, flexion
// mode + per-finger draws: between 30 and 44 per pose), with the reference's fp32 operation sequence (explicit _rn
// intrinsics: no FMA contraction), so that fed the same uniforms it is bit-identical to the reference.  offsets[i] is the
// index of pose i's first uniform (sequential stream: the host walks the mode draws; independent streams: i * 48).
namespace {

struct USrc {
    const float* u;
    int p;
    __device__ __forceinline__ float next() { return u[p++]; }
};
constexpr float kPiF = 3.14159274101257324f;              // float32(math.pi): the scalar is cast to the tensor's dtype

__device__ __forceinline__ float ja_deg(float a) { return __fdiv_rn(__fmul_rn(a, kPiF), 180.f); }       // a * pi / 180
// curr_flex = (rand * 30 + base) * pi / 180 + (rand * 20 - 10) * pi / 180
__device__ __forceinline__ float ja_curr(USrc& s, float base) {
    const float a = ja_deg(__fadd_rn(__fmul_rn(s.next(), 30.f), base));
    const float b = ja_deg(__fsub_rn(__fmul_rn(s.next(), 20.f), 10.f));
    return __fadd_rn(a, b);
}
// the three "curled" finger shapes (:42-103) differ only in the additive constants of their three flexion draws
__device__ __forceinline__ void ja_curled(USrc& s, float b1, float b2, float b3, float* f) {
    float f1 = -0.2f, f2 = -0.4f, f3 = -0.34f;
    float c = ja_curr(s, b1);
    f1 = __fadd_rn(f1, c);                                // 1.0 * curr_flex
    f2 = __fadd_rn(f2, __fmul_rn(0.2f, c));
    c = ja_curr(s, b2);
    f1 = __fadd_rn(f1, __fmul_rn(0.2f, c));
    f2 = __fadd_rn(f2, c);
    f3 = __fadd_rn(f3, __fmul_rn(0.7f, c));
    c = ja_curr(s, b3);
    f2 = __fadd_rn(f2, __fmul_rn(0.2f, c));
    f3 = __fadd_rn(f3, c);
    f[0] = f1; f[1] = f2; f[2] = f3;
}
// finger shape by index: 0 straight, 1 open, 2 half open, 3 pinching, 4 closed (_set_rand_flex's numbering, :145-158)
__device__ __forceinline__ void ja_finger(USrc& s, int shape, float* f) {
    if (shape == 0) {
        f[0] = __fsub_rn(__fmul_rn(s.next(), 0.25f), 0.25f);
        f[1] = __fsub_rn(__fmul_rn(s.next(), 0.4f), 0.4f);
        f[2] = __fsub_rn(__fmul_rn(s.next(), 0.34f), 0.34f);
    } else if (shape == 1) {
        f[0] = __fsub_rn(__fmul_rn(s.next(), 0.25f), 0.1f);
        f[1] = __fsub_rn(__fmul_rn(s.next(), 0.4f), 0.1f);
        f[2] = __fsub_rn(__fmul_rn(s.next(), 0.34f), 0.1f);
    } else if (shape == 2) {
        ja_curled(s, 0.f, 60.f, 60.f, f);                 // (rand*30)*pi/180 == (rand*30 + 0)*pi/180 bit for bit
    } else if (shape == 3) {
        ja_curled(s, 60.f, 5.f, 5.f, f);
    } else {
        ja_curled(s, 60.f, 60.f, 60.f, f);
    }
}
__device__ __forceinline__ int ja_mode(USrc& s, float n) { return (int)__fmul_rn(s.next(), n); }        // int(torch.rand(1) * n)

__global__ void pose_sample_kernel(const float* __restrict__ u, const int* __restrict__ offsets, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    USrc s{u, offsets[i]};
    float* p = out + (size_t)i * 26;
    // _set_palm (:21-29)
    p[0] = __fsub_rn(__fmul_rn(s.next(), 6.28f), 3.14f);
    p[1] = __fmul_rn(-s.next(), 3.14f);
    p[2] = __fsub_rn(__fmul_rn(s.next(), 6.28f), 3.14f);
    p[3] = __fsub_rn(__fmul_rn(s.next(), 30.f), 15.f);
    p[4] = __fsub_rn(__fmul_rn(s.next(), 30.f), 15.f);
    p[5] = __fsub_rn(__fmul_rn(s.next(), 50.f), 35.f);
    // _set_abduct (:31-39): index, middle, ring, pinky
    const float spread = __fdiv_rn(__fsub_rn(s.next(), 0.35f), 1.55f);
    const float ka[4] = {1.55f, 0.75f, -0.75f, -2.2f};
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const float r = ja_deg(__fsub_rn(__fmul_rn(s.next(), 10.f), 5.f));
        p[6 + 4 * f] = __fmul_rn(ka[f], __fadd_rn(spread, r));
    }
    // _set_thumb (:117-128) -> parameter[22:26] = (abduct, flex, 0.25 flex, flex_3)
    {
        const float sel = s.next();
        const float v = s.next();
        const float flex = sel < 0.5f ? __fsub_rn(__fmul_rn(v, 0.35f), 0.25f) : __fadd_rn(__fmul_rn(v, 0.6f), 0.1f);
        const float f3 = __fsub_rn(__fmul_rn(s.next(), 2.f), 1.7f);
        p[22] = __fsub_rn(s.next(), 0.5f);
        p[23] = flex;
        p[24] = __fmul_rn(0.25f, flex);
        p[25] = f3;
    }
    // _set_flex (:160-214).  Per-finger rule: >= 0 a fixed shape, -1 rand_open (:130-137), -2 rand_close (:139-144), -3 rand_flex.
    // The second `mode == 8` branch of the reference (index & pinky open) is unreachable; mode 10 (rand*10 rounding up to 10.0)
    // leaves `flex` undefined there and is mapped to 9 here.
    const int mode = min(ja_mode(s, 10.f), 9);
    int rule[4];
    if (mode <= 4) { rule[0] = rule[1] = rule[2] = rule[3] = mode; }
    else if (mode == 5) { rule[0] = -1; rule[1] = -2; rule[2] = -2; rule[3] = -2; }
    else if (mode == 6) { rule[0] = -2; rule[1] = -2; rule[2] = -2; rule[3] = -1; }
    else if (mode == 7) { rule[0] = -1; rule[1] = -1; rule[2] = -2; rule[3] = -2; }
    else if (mode == 8) { rule[0] = -2; rule[1] = -1; rule[2] = -1; rule[3] = -1; }
    else { rule[0] = rule[1] = rule[2] = rule[3] = -3; }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        int shape = rule[f];
        if (shape == -1) shape = min(ja_mode(s, 3.f), 2);                 // straight / open / half open
        else if (shape == -2) shape = 3 + min(ja_mode(s, 2.f), 1);        // pinching / closed
        else if (shape == -3) shape = min(ja_mode(s, 5.f), 4);
        ja_finger(s, shape, p + 7 + 4 * f);
    }
}

}  // namespace

// u: fp32 uniforms in [0,1); offsets int32 [n] (first uniform of every pose; a pose reads at most 44); out fp32 [n,26]
SH_EXPORT int sh_sample_poses(const void* u, const void* offsets, int n, void* out, void* stream) {
    SH_REQUIRE(n >= 0, "sh_sample_poses: bad n");
    if (n == 0) return SH_OK;
    SH_REQUIRE(u && offsets && out, "sh_sample_poses: null pointer");
    pose_sample_kernel<<<sh_div_up(n, 64), 64, 0, (cudaStream_t)stream>>>((const float*)u, (const int*)offsets, n, (float*)out);
    SH_CHECK_LAUNCH("pose_sample_kernel");
    return SH_OK;
}
