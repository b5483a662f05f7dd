// Result on B200 (gpurun, round 1): K-major SWIZZLE_128B descriptors with base_offset 0 address correctly for EVERY row shift and
// for 8-row group pitches of 8, 10, 16 and 18 rows (SBO 1024 / 1280 / 2048 / 2304 B); base_offset != 0 is wrong for shifted starts.
// Probe: does a tcgen05 shared-memory descriptor whose start address is shifted by whole 128-byte rows (not 1024-byte aligned)
// address the rows TMA wrote with SWIZZLE_128B?  Decides whether a 3x3 convolution can derive its nine taps from ONE halo
// tile in shared memory (start shifted by dw rows, 8-row groups `sbo` bytes apart) instead of nine TMA boxes.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../spherehand_b200/csrc umma_probe.cu ../spherehand_b200/csrc/build/conv_tc.o ../spherehand_b200/csrc/build/api.o -o umma_probe
#include "tc_common.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>

constexpr int kRows = 288;      // rows of the source matrix (each 64 bf16 = 128 B)

__device__ __forceinline__ uint64_t desc_k(uint32_t addr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// mode 0: K-major A = rows [shift + (m/8)*group_rows + m%8] of SRC, B = identity (K-major)  -> D[m][n] = SRC[row(m)][n]
// mode 1: MN-major B = rows [shift + k] of SRC (K index = row), A = identity (MN-major)      -> D[m][n] = SRC[shift+m][n], m < 64
__global__ void __launch_bounds__(192, 1) probe_kernel(const __grid_constant__ CUtensorMap tmSrc, const __grid_constant__ CUtensorMap tmId,
                                                       int mode, int shift, int group_rows, int use_base_off, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_src = smem;                                   // kRows x 128 B
    uint8_t* s_id = smem + 512 * 128;                        // 128 x 128 B identity (rows 64.. zero)
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_id + 128 * 128);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 512 * 128 + 128 * 128);           // full boxes count, out-of-bounds rows included
        tma_load_2d(s_src, &tmSrc, bar, 0, 0);
        tma_load_2d(s_src + 256 * 128, &tmSrc, bar, 0, 256);
        tma_load_2d(s_id, &tmId, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        if (mode == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
            const uint32_t a0 = smem_u32(s_src) + shift * 128;
            const uint32_t bo = use_base_off ? (uint32_t)(shift & 7) : 0u;
            for (int kk = 0; kk < 4; ++kk)
                umma_bf16(tm, desc_k(a0, group_rows * 128, bo) + 2 * kk, desc_k(smem_u32(s_id), 1024, 0) + 2 * kk, idesc, kk != 0);
        } else {
            const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
            const uint32_t b0 = smem_u32(s_src) + shift * 128;
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t ba = b0 + kk * 2048;
                const uint32_t bo = use_base_off ? (uint32_t)((ba >> 7) & 7) : 0u;
                // identity A: [k rows][m 128]: two 64-wide chunks, chunk 1 (m >= 64) all zero -> LBO = 64 rows * 128 B
                umma_bf16(tm, desc_mn(smem_u32(s_id) + kk * 2048, 64 * 128, 1024, 0), desc_mn(ba, 64 * 128, 1024, bo), idesc, kk != 0);
            }
        }
        umma_commit(done);
    }
    if (warp >= 2) {
        mbar_wait(done, 0);
        tc_fence_after();
        const int quad = warp & 3;
        uint32_t v[32];
        for (int c0 = 0; c0 < 64; c0 += 32) {
            tmem_ld32(tm + ((uint32_t)(quad * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[(quad * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 64);
}

static float src_val(int r, int c) { return (float)(((r * 3 + c * 5) % 61) - 30); }

int main() {
    std::vector<__nv_bfloat16> h_src(kRows * 64), h_id(128 * 64);
    for (int r = 0; r < kRows; ++r) for (int c = 0; c < 64; ++c) h_src[r * 64 + c] = __float2bfloat16(src_val(r, c));
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 64; ++c) h_id[r * 64 + c] = __float2bfloat16(r == c ? 1.f : 0.f);
    __nv_bfloat16 *d_src, *d_id; float* d_out;
    cudaMalloc(&d_src, h_src.size() * 2); cudaMalloc(&d_id, h_id.size() * 2); cudaMalloc(&d_out, 128 * 64 * 4);
    cudaMemcpy(d_src, h_src.data(), h_src.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_id, h_id.data(), h_id.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmSrc, tmId;
    const uint64_t dims[2] = {64, (uint64_t)kRows}, strides[1] = {128}; const uint32_t box[2] = {64, 256};
    if (sh_make_tmap_bf16(&tmSrc, d_src, 2, dims, strides, box)) { printf("tmap failed\n"); return 1; }
    const uint64_t dimi[2] = {64, 128}; const uint32_t boxi[2] = {64, 128};
    if (sh_make_tmap_bf16(&tmId, d_id, 2, dimi, strides, boxi)) { printf("tmap failed\n"); return 1; }
    const size_t smem = (size_t)(512 + 128) * 128 + 1024 + 256;     // second source box lands at row 256: reserve 512 rows
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> h_out(128 * 64);
    struct Case { int mode, shift, group_rows, bo; };
    std::vector<Case> cases;
    for (int bo = 0; bo < 2; ++bo) {
        for (int sh : {0, 1, 2, 3, 7, 8, 9, 17}) cases.push_back({0, sh, 8, bo});      // contiguous groups, shifted start
        for (int sh : {0, 1, 2, 17, 18, 34}) cases.push_back({0, sh, 16, bo});         // halo pitch 16 rows (SBO 2048)
        for (int sh : {0, 1, 2, 13}) cases.push_back({0, sh, 18, bo});                 // halo pitch 18 rows (SBO 2304, not 1024-aligned)
        for (int sh : {0, 1, 2, 10, 11, 12, 20, 21, 22}) cases.push_back({0, sh, 10, bo});   // halo pitch 10 rows (SBO 1280): the 8x16 tile of a 3x3 conv
        for (int sh : {0, 1, 2, 3, 9, 19}) cases.push_back({1, sh, 8, bo});            // MN-major, shifted K rows
    }
    for (const Case& c : cases) {
        cudaMemset(d_out, 0xff, 128 * 64 * 4);
        probe_kernel<<<1, 192, smem>>>(tmSrc, tmId, c.mode, c.shift, c.group_rows, c.bo, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d shift %d pitch %d bo %d: CUDA error %s\n", c.mode, c.shift, c.group_rows, c.bo, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(h_out.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            float exp;
            if (c.mode == 0) exp = src_val(c.shift + (m / 8) * c.group_rows + m % 8, n);
            else exp = m < 64 ? src_val(c.shift + m, n) : 0.f;
            if (h_out[m * 64 + n] != exp) { if (first < 0) first = m * 64 + n; ++bad; }
        }
        printf("mode %d shift %2d pitch %2d base_off %d : %s (%d mismatches%s)\n", c.mode, c.shift, c.group_rows, c.bo, bad ? "FAIL" : "ok", bad,
               bad ? "" : "");
        if (bad && first >= 0) printf("    first mismatch m=%d n=%d got %g\n", first / 64, first % 64, h_out[first]);
    }
    return 0;
}
