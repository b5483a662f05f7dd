// Small pose-space loss heads, forward value + analytic gradient in one launch (sm_100a).
//
// Replaces
//   MultiviewConsistencyLoss.forward  /root/reference/mesh/multiview_utility.py:138-167 (hm_weight=None)
//   CollisionLoss.forward             /root/reference/mesh/render.py:168-176  (690 pairs, render.py:153-162)
//   BoneLengthLoss.forward            /root/reference/mesh/render.py:196-206  (35 pairs, mesh/bone_length.py:36-56)
// and their autograd backward.  Reference quirk reproduced on purpose: collision and
// bone-length flatten [B,V,41,3] to [B,123,3] and index 0..40, i.e. they only see view 0 of every tuple.
//
// HBM layout: cam fp32 [B,V,4,4]; joints fp32 [B,V,J,3]; out losses fp32[3] = (consistency, collision,
// bone_length); grads fp32 [3,B,V,J,3] (d loss_t / d joints, upstream 1).  One CTA per tuple b.
#include "common.cuh"

namespace {

constexpr int kMaxV = 8;
constexpr int kMaxJ = 64;
constexpr int kColPairs = 690;
constexpr int kBonePairs = 35;

__constant__ uint8_t c_col_a[kColPairs];
__constant__ uint8_t c_col_b[kColPairs];
__constant__ uint8_t c_bone_a[kBonePairs];
__constant__ uint8_t c_bone_b[kBonePairs];
__constant__ float c_bone_lo[kBonePairs];   // (0.80 L)^2
__constant__ float c_bone_hi[kBonePairs];   // (1.05 L)^2
bool g_tables_ready = false;

const float kBoneLen[kBonePairs] = {
    25.212656021118164f, 18.249488830566406f, 27.5742244720459f, 38.532264709472656f, 25.10819435119629f,
    31.173757553100586f, 18.329626083374023f, 19.15080451965332f, 16.209327697753906f, 21.52261734008789f,
    32.740535736083984f, 30.58920669555664f, 33.205970764160156f, 11.672294616699219f, 17.084707260131836f,
    17.084720611572266f, 16.697546005249023f, 23.92103385925293f, 20.87999725341797f, 22.58038330078125f,
    27.55999755859375f, 15.471183776855469f, 13.214692115783691f, 21.748210906982422f, 13.021653175354004f,
    16.643720626831055f, 18.83765983581543f, 12.724685668945312f, 16.238431930541992f, 18.04928970336914f,
    11.045844078063965f, 11.320968627929688f, 30.078536987304688f, 16.255985260009766f, 19.434825897216797f};

int upload_tables() {
    if (g_tables_ready) return SH_OK;
    uint8_t a[kColPairs], b[kColPairs];
    int n = 0;
    for (int i = 0; i < 11; ++i)
        for (int j = 11; j < 41; ++j) { a[n] = (uint8_t)i; b[n] = (uint8_t)j; ++n; }
    for (int i = 11; i < 41; ++i)
        for (int j = i + 1; j < 41; ++j)
            if ((i - 11) / 6 != (j - 11) / 6) { a[n] = (uint8_t)i; b[n] = (uint8_t)j; ++n; }
    if (n != kColPairs) return SH_ERR_INVALID;
    const uint8_t ba0[20] = {3, 2, 3, 8, 2, 2, 9, 8, 4, 8, 7, 4, 6, 7, 0, 5, 7, 7, 6, 6};
    const uint8_t bb0[20] = {2, 9, 8, 2, 4, 10, 10, 4, 10, 7, 4, 6, 10, 6, 5, 1, 0, 5, 5, 1};
    uint8_t ba[kBonePairs], bb[kBonePairs];
    float lo[kBonePairs], hi[kBonePairs];
    for (int i = 0; i < 20; ++i) { ba[i] = ba0[i]; bb[i] = bb0[i]; }
    for (int f = 0; f < 5; ++f)
        for (int k = 0; k < 3; ++k) { ba[20 + 3 * f + k] = (uint8_t)(11 + 2 * k + 6 * f); bb[20 + 3 * f + k] = (uint8_t)(12 + 2 * k + 6 * f); }
    for (int i = 0; i < kBonePairs; ++i) {
        const float l0 = kBoneLen[i] * 0.80f, l1 = kBoneLen[i] * 1.05f;
        lo[i] = l0 * l0;
        hi[i] = l1 * l1;
    }
    if (cudaMemcpyToSymbol(c_col_a, a, sizeof(a)) != cudaSuccess) return SH_ERR_CUDA;
    if (cudaMemcpyToSymbol(c_col_b, b, sizeof(b)) != cudaSuccess) return SH_ERR_CUDA;
    if (cudaMemcpyToSymbol(c_bone_a, ba, sizeof(ba)) != cudaSuccess) return SH_ERR_CUDA;
    if (cudaMemcpyToSymbol(c_bone_b, bb, sizeof(bb)) != cudaSuccess) return SH_ERR_CUDA;
    if (cudaMemcpyToSymbol(c_bone_lo, lo, sizeof(lo)) != cudaSuccess) return SH_ERR_CUDA;
    if (cudaMemcpyToSymbol(c_bone_hi, hi, sizeof(hi)) != cudaSuccess) return SH_ERR_CUDA;
    g_tables_ready = true;
    return SH_OK;
}

__global__ void __launch_bounds__(128) pose_losses_kernel(const float* __restrict__ cam, const float* __restrict__ joints,
                                                          int B, int V, int J, int flags, float min_sq_dist,
                                                          double* __restrict__ acc, float* __restrict__ grads) {
    __shared__ float s_j[kMaxV * kMaxJ * 3];     // joints of this tuple
    __shared__ float s_q[kMaxV * kMaxJ * 3];     // canonical-frame joints, then d/dq
    __shared__ float s_cam[kMaxV * 12];
    __shared__ float s_g[2][41 * 3];             // collision / bone-length grads (view 0)
    __shared__ float s_loss[3];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nj = V * J * 3;
    for (int t = tid; t < nj; t += blockDim.x) s_j[t] = joints[(size_t)b * nj + t];
    for (int t = tid; t < V * 12; t += blockDim.x) s_cam[t] = cam[((size_t)b * V + t / 12) * 16 + t % 12];
    for (int t = tid; t < 2 * 41 * 3; t += blockDim.x) (&s_g[0][0])[t] = 0.f;
    if (tid < 3) s_loss[tid] = 0.f;
    __syncthreads();
    const size_t gstride = (size_t)B * nj;

    // ---------------- multi-view consistency
    if (flags & 1) {
        for (int t = tid; t < nj; t += blockDim.x) {
            const int v = t / (J * 3), k = (t / 3) % J, x = t % 3;
            const float* T = s_cam + v * 12 + x * 4;
            const float* p = s_j + (v * J + k) * 3;
            s_q[t] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3];
        }
        __syncthreads();
        const float n_inv = 1.f / ((float)B * V * J * 3);
        float lsum = 0.f;
        float dq[kMaxV];
        for (int t = tid; t < J * 3; t += blockDim.x) {
            float q[kMaxV];
            for (int v = 0; v < V; ++v) q[v] = s_q[v * J * 3 + t];
            // torch.median: the lower median, i.e. rank (V-1)/2 in ascending order
            int mi = 0;
            for (int v = 0; v < V; ++v) {
                int rank = 0;
                for (int u = 0; u < V; ++u) rank += (q[u] < q[v]) || (q[u] == q[v] && u < v);
                if (rank == (V - 1) / 2) mi = v;
            }
            const float med = q[mi];
            float tot = 0.f;
            for (int v = 0; v < V; ++v) {
                const float d = med - q[v];
                lsum += d * d;
                dq[v] = -2.f * d * n_inv;
                tot += 2.f * d * n_inv;
            }
            dq[mi] += tot;
            for (int v = 0; v < V; ++v) s_q[v * J * 3 + t] = dq[v];   // only this thread touches column t
        }
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) atomicAdd(&s_loss[0], lsum * n_inv);
        __syncthreads();
        for (int t = tid; t < nj; t += blockDim.x) {
            const int v = t / (J * 3), k = (t / 3) % J, c = t % 3;
            const float* T = s_cam + v * 12;
            const float* g = s_q + (v * J + k) * 3;
            grads[(size_t)b * nj + t] = T[0 + c] * g[0] + T[4 + c] * g[1] + T[8 + c] * g[2];   // R^T dq
        }
    } else {
        for (int t = tid; t < nj; t += blockDim.x) grads[(size_t)b * nj + t] = 0.f;
    }

    // ---------------- collision (sum over batch and pairs; view 0 only)
    if ((flags & 2) && J >= 41) {
        float lsum = 0.f;
        for (int p = tid; p < kColPairs; p += blockDim.x) {
            const int a = c_col_a[p], c = c_col_b[p];
            const float dx = s_j[a * 3] - s_j[c * 3], dy = s_j[a * 3 + 1] - s_j[c * 3 + 1], dz = s_j[a * 3 + 2] - s_j[c * 3 + 2];
            const float v = min_sq_dist - (dx * dx + dy * dy + dz * dz);
            if (v > 0.f) {
                lsum += v;
                atomicAdd(&s_g[0][a * 3 + 0], -2.f * dx); atomicAdd(&s_g[0][a * 3 + 1], -2.f * dy); atomicAdd(&s_g[0][a * 3 + 2], -2.f * dz);
                atomicAdd(&s_g[0][c * 3 + 0], 2.f * dx); atomicAdd(&s_g[0][c * 3 + 1], 2.f * dy); atomicAdd(&s_g[0][c * 3 + 2], 2.f * dz);
            }
        }
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) atomicAdd(&s_loss[1], lsum);
    }
    // ---------------- bone length (two means over B*35; view 0 only)
    if ((flags & 4) && J >= 41) {
        const float n_inv = 1.f / ((float)B * kBonePairs);
        float lsum = 0.f;
        for (int p = tid; p < kBonePairs; p += blockDim.x) {
            const int a = c_bone_a[p], c = c_bone_b[p];
            const float dx = s_j[a * 3] - s_j[c * 3], dy = s_j[a * 3 + 1] - s_j[c * 3 + 1], dz = s_j[a * 3 + 2] - s_j[c * 3 + 2];
            const float sq = dx * dx + dy * dy + dz * dz;
            float coef = 0.f;
            if (c_bone_lo[p] - sq > 0.f) { lsum += (c_bone_lo[p] - sq) * n_inv; coef -= 2.f * n_inv; }
            if (sq - c_bone_hi[p] > 0.f) { lsum += (sq - c_bone_hi[p]) * n_inv; coef += 2.f * n_inv; }
            if (coef != 0.f) {
                atomicAdd(&s_g[1][a * 3 + 0], coef * dx); atomicAdd(&s_g[1][a * 3 + 1], coef * dy); atomicAdd(&s_g[1][a * 3 + 2], coef * dz);
                atomicAdd(&s_g[1][c * 3 + 0], -coef * dx); atomicAdd(&s_g[1][c * 3 + 1], -coef * dy); atomicAdd(&s_g[1][c * 3 + 2], -coef * dz);
            }
        }
        lsum = warp_sum(lsum);
        if ((tid & 31) == 0) atomicAdd(&s_loss[2], lsum);
    }
    __syncthreads();
    for (int t = tid; t < nj; t += blockDim.x) {
        const bool v0 = t < 41 * 3 && J >= 41;      // flattened (V*J) index < 41 <=> view 0, joints 0..40
        grads[gstride + (size_t)b * nj + t] = v0 ? s_g[0][t] : 0.f;
        grads[2 * gstride + (size_t)b * nj + t] = v0 ? s_g[1][t] : 0.f;
    }
    if (tid < 3) atomicAdd(&acc[tid], (double)s_loss[tid]);
}

__global__ void pose_losses_finish_kernel(const double* __restrict__ acc, float* __restrict__ losses) {
    if (threadIdx.x < 3) losses[threadIdx.x] = (float)acc[threadIdx.x];
}

}  // namespace

// flags: bit0 consistency, bit1 collision, bit2 bone length.  scratch: >= 32 bytes, 8-byte aligned.
SH_EXPORT int sh_pose_losses_fwdbwd(const void* cam, const void* joints, int B, int V, int J, int flags,
                                     float min_dist, void* losses3, void* grads3, void* scratch, void* stream) {
    SH_REQUIRE(cam && joints && losses3 && grads3 && scratch, "sh_pose_losses_fwdbwd: null pointer");
    SH_REQUIRE(B >= 1 && V >= 1 && V <= kMaxV && J >= 1 && J <= kMaxJ, "sh_pose_losses_fwdbwd: bad B/V/J");
    SH_REQUIRE(!(flags & 6) || J == 41, "sh_pose_losses_fwdbwd: collision/bone-length tables need J == 41 (got %d)", J);
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = upload_tables();
    if (rc != SH_OK) { snprintf(g_sh_last_error, sizeof(g_sh_last_error), "sh_pose_losses_fwdbwd: table upload failed"); return rc; }
    SH_CUDA(cudaMemsetAsync(scratch, 0, 32, st));
    pose_losses_kernel<<<B, 128, 0, st>>>((const float*)cam, (const float*)joints, B, V, J, flags, min_dist * min_dist,
                                          (double*)scratch, (float*)grads3);
    SH_CHECK_LAUNCH("pose_losses_kernel");
    pose_losses_finish_kernel<<<1, 32, 0, st>>>((const double*)scratch, (float*)losses3);
    SH_CHECK_LAUNCH("pose_losses_finish_kernel");
    return SH_OK;
}
