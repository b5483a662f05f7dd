/* spherehand_b200 — C ABI of the B200 (sm_100a) hot path of melonwan/sphereHand.
 *
 * libspherehand_b200.so exports exactly the functions declared here.  All of them:
 *   - take plain device pointers and sizes (no torch types), never allocate device memory, never synchronise;
 *   - launch on the CUDA stream passed last (a `cudaStream_t` cast to `void*`; NULL = legacy default stream, which is
 *     what the reference kernel uses, depth_rasterization_cuda_kernel.cu:125);
 *   - return 0 on success, 1 = invalid argument, 2 = CUDA error, 3 = unsupported shape; `sh_last_error()` holds the text;
 *   - borrow their inputs (never mutated) and write only the buffers documented as outputs, which the caller owns.
 * Tensors are contiguous, fp32 unless stated; "bf16" is __nv_bfloat16.  file:line citations are into the reference
 * tree (melonwan/sphereHand @ 4a5bae9) and name the interface each entry replaces.
 *
 * The reference's only FFI is the pybind function `depth_rasterization.forward(width, height, vertices)`
 * (mesh/cuda_kernel/depth_rasterization_cuda.cpp:15-25); `sh_tri_raster_fwd` is its drop-in.  Every other entry
 * is what a native binding for the corresponding Python module of the hot path would bind.
 */
#ifndef SPHEREHAND_B200_H
#define SPHEREHAND_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ library */
int sh_abi_version(void);          /* 3 */
const char* sh_build_arch(void);      /* "sm_100a" */
const char* sh_last_error(void);
long sh_launch_count(void);          /* kernel launches issued through this library since load */

/* ------------------------------------------------------------------------------------------------ R1: triangle rasteriser
 * Replaces depth_rasterization_cuda_forward + kernel, mesh/cuda_kernel/depth_rasterization_cuda_kernel.cu:18-134.
 * face_vertices [B,F,3,3]; out [B,H,W], written with 1000.0 where nothing is drawn. */
int sh_tri_raster_fwd(const void* face_vertices, int B, int F, int W, int H, void* out, void* stream);
/* Same rasteriser evaluated only at the pixels {c*step + off0 + o, o < noff}^2 of the virtual W x H image, compact
 * output [B,oh,ow] (oh = H/step*noff): exactly the samples DepthRasterization.forward's bilinear resize reads
 * (mesh/render.py:310-311; step 5/off 2/noff 1 for 640->128, step 10/off 4/noff 2 for 640->64). */
int sh_tri_raster_lattice_fwd(const void* face_vertices, int B, int F, int W, int H, int step, int off0, int noff,
                              void* out, int oh, int ow, void* stream);

/* ------------------------------------------------------------------------------------------------ R2: sphere renderer
 * Replaces BallRender.forward (mesh/render.py:26-53) + min over spheres (mesh/render.py:89,
 * mesh/multiview_utility.py:76) and their autograd backward.
 * spheres float4 [N*J] = (cx,cy,cz,r), 16-byte aligned; depth [N,H,W]; idx uint8 [N,H,W] (255 = background);
 * grad_spheres float4 [N*J] = dL/d(cx,cy,cz,r) (overwritten).  1 <= J <= 64. */
int sh_sphere_render_fwd(const void* spheres, int N, int J, int H, int W, void* depth, void* idx, void* stream);
int sh_sphere_render_bwd(const void* grad_depth, const void* idx, const void* spheres, int N, int J, int H, int W,
                         void* grad_spheres, void* stream);

/* ------------------------------------------------------------------------------------------------ fused MutualProjectionLoss
 * Replaces MutualProjectionLoss.forward (mesh/multiview_utility.py:90-130) incl. MutualTransformation (:13-30),
 * MutualProjection (:55-77), DataToModelLoss (mesh/render.py:123-142) and the backward to `joints`.
 * cam, inv_cam [B,V,4,4]; joints [B,V,J,3]; real [B,V,H,W]; radii [J];
 * out: projected_dms [B,V,V,H,W]; loss3 = (loss, model_to_data, data_to_model); grad_joints [B,V,J,3] = dloss/djoints.
 * scratch: sh_mvproj_scratch_bytes(B,V,J) bytes, 16-byte aligned.  V <= 8, B*V*V <= 65535. */
size_t sh_mvproj_scratch_bytes(int B, int V, int J);
int sh_mvproj_loss_fwdbwd(const void* cam, const void* inv_cam, const void* joints, const void* real, const void* radii,
                          int B, int V, int J, int H, int W, int is_mv, void* projected_dms, void* loss3,
                          void* grad_joints, void* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------ small pose-space heads
 * Replaces MultiviewConsistencyLoss.forward (mesh/multiview_utility.py:138-167, hm_weight=None), CollisionLoss.forward
 * (mesh/render.py:168-176) and BoneLengthLoss.forward (mesh/render.py:196-206) with their backward.
 * flags: bit0 consistency, bit1 collision, bit2 bone length (the latter two need J == 41 and only see view 0, as the
 * reference does).  losses3 = (consistency, collision, bone_length); grads3 [3,B,V,J,3]; scratch >= 32 bytes. */
int sh_pose_losses_fwdbwd(const void* cam, const void* joints, int B, int V, int J, int flags, float min_dist,
                          void* losses3, void* grads3, void* scratch, void* stream);

/* Replaces PoseVae.prior_loss (network/pose_vae.py:81-89).  x [M,123] (= xyz/100), eps [M,32] ~ N(0,1) drawn by the
 * host, weight_blob: sh_vae_blob_floats() floats packed as documented in csrc/pose_vae.cu.
 * loss3 = (loss, recon_mse, kld); grad_x [M,123]; scratch >= 16 bytes, 8-byte aligned.
 * M_mean >= M: number of rows the reconstruction MSE is averaged over (= M on one GPU; = the GLOBAL batch rows when the
 * batch is sharded over ranks, so that summing rank gradients reproduces the single-GPU gradient; the KLD is a sum). */
size_t sh_vae_blob_floats(void);
int sh_vae_prior_fwdbwd(const void* x, const void* eps, const void* weight_blob, int M, int M_mean, void* loss3,
                        void* grad_x, void* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------ heat-map heads
 * Replaces RecoverXYZCoordinateFromHeatmap.forward (network/util_modules.py:182-201), the uv/d channel split of
 * HeatmapEstimationNetwork.forward (network/create_network_and_criterion.py:111-123) and the heat-map MSE terms of
 * MultiTaskLoss.forward (:188-193, :231-235).  score [N,C,h,w] NCHW (uv = channels [0,J), depth = [J,2J));
 * samples n < Ns are synthetic (MSE against target_uv [Ns,J,h,w]), the rest real (MSE against 0).
 * fwd: xyz [N,J,3] mm; sse2 double[2] = (sum (uv-target)^2, sum uv_real^2) or NULL.
 * bwd: gscore [N,C,h,w] = d/dscore of (xyz . gxyz) + c_synt/2*sse_synt + c_real/2*sse_real  (overwrites 2J channels). */
int sh_softargmax_fwd(const void* score, int N, int Ns, int J, int C, int h, int w, float depth_scale_inv,
                      const void* target_uv, void* xyz, void* sse2, void* stream);
int sh_softargmax_bwd(const void* score, const void* gxyz, int N, int Ns, int J, int C, int h, int w,
                      float depth_scale_inv, const void* target_uv, float c_synt, float c_real, void* gscore,
                      void* stream);
/* The pair the fused train step uses: the forward also leaves the per-(sample, joint) scalars of the soft-argmax in
 * aux fp32 [N,J,8] (softmax max, 1/sum, u, v, 1/(sum relu + 1e-5), d, -, -), and the backward writes d loss / d score from them
 * pixel-major, directly as the bf16 NHWC tensor [N,h,w,Cp] (padding channels zero) the network's backward pass consumes --
 * no fp32 NCHW score gradient, no layout-conversion launch.  Built for J = 41, Cp = 128. */
int sh_softargmax_fwd_aux(const void* score, int N, int Ns, int J, int C, int h, int w, float depth_scale_inv,
                          const void* target_uv, void* xyz, void* sse2, void* aux, void* stream);
int sh_softargmax_bwd_nhwc(const void* score, const void* gxyz, const void* aux, int N, int Ns, int J, int C, int h, int w,
                           float depth_scale_inv, const void* target_uv, float c_synt, float c_real, void* dscore, int Cp,
                           void* stream);

/* ------------------------------------------------------------------------------------------------ synthetic branch
 * sh_fk_fwd replaces HandTransformationMat.forward (mesh/kinematicsTransformation.py:169-177) and, when `scales`
 * [B,3] is given, RandScale.forward (mesh/pointTransformation.py:135-148).  params [B,26]; offset_mats,
 * inv_offset_mats [17,4,4]; mats [B,17,4,4]. */
int sh_fk_fwd(const void* params, const void* scales, const void* offset_mats, const void* inv_offset_mats, int B,
              void* mats, void* stream);
/* LinearBlendSkinning.forward (mesh/pointTransformation.py:39-46) over a CSR of the non-zero weights, optionally fused
 * with OthographicalProjection.forward (:84-99): mode 0 none, 1 per-sample focal jitter rand_f [B], 2 K-matrix.
 * row_ptr int32 [Nv+1], bone int32 [nnz], wv float4 [nnz]; out_points float4 [B,Nv]. */
int sh_lbs_fwd(const void* mats, const void* row_ptr, const void* bone, const void* wv, int B, int Nv, int right_hand,
               int mode, float cx, float cy, float fx, float fy, const void* rand_f, void* out_points, void* stream);
/* face_vertices[b,f,3,3] = points[b, faces[f,:], 0:3]   (mesh/render.py:308-309); faces int32 [F,3]. */
int sh_gather_faces(const void* points, const void* faces, int B, int Nv, int F, void* face_vertices, void* stream);
/* clamp(max=100) + bilinear resize + * depth_scale on the lattice z-buffer (mesh/render.py:286,311;
 * network/util_modules.py:112): z [B,S*noff,S*noff] -> dm [B,S,S]. */
int sh_lattice_to_depth(const void* z, int B, int S, int noff, float depth_scale, void* dm, void* stream);
/* DepthNoise.forward (network/util_modules.py:60-84) with the three N(0,1) draws supplied by the caller. */
int sh_depth_noise(const void* dm, const void* nx, const void* ny, const void* nz, int B, int H, int W, float sx,
                   float sy, float sz, void* out, void* stream);
/* PoseDenoiser.forward in eval mode (network/pose_denoiser.py:56-73): out = fea with the out_idx entries replaced by
 * MLP(fea[:, in_idx] * scale) / scale.  blob: sh_pose_denoiser_blob_floats(n_in, n_out) floats = W1^T [n_in,256], b1, gamma1,
 * beta1, W2^T [256,256], b2, gamma2, beta2, W3^T [256,n_out], b3.  fea, out fp32 [M,n_fea]; indices int32. */
size_t sh_pose_denoiser_blob_floats(int n_in, int n_out);
int sh_pose_denoiser_fwd(const void* fea, const void* in_idx, const void* out_idx, const void* blob, int M, int n_fea, int n_in,
                         int n_out, float scale, void* out, void* stream);
/* JointAngleDataset.__getitem__ (dataset/joint_angle.py:21-233) for n poses in one launch: pose i consumes its slice of a
 * pre-drawn uniform stream (u[offsets[i]] ..., at most 44 values) in the reference's order with the reference's fp32 operation
 * sequence, so the same uniforms give the reference's pose bit for bit.  u fp32, offsets int32 [n], out fp32 [n,26]. */
int sh_sample_poses(const void* u, const void* offsets, int n, void* out, void* stream);
/* HeatmapRender.forward + InverseOthographicalProjection (mesh/render.py:226-248, 274-279): uvd float4 [B,J] ->
 * uv_hms, d_hms [B,J,hm,hm], xyz float4 [B,J]. */
int sh_heatmap_render(const void* uvd, int B, int J, int hm, float sigma, float uv_scale, float depth_scale, float cx,
                      float cy, float fx, float fy, void* uv_hms, void* d_hms, void* xyz, void* stream);

/* ------------------------------------------------------------------------------------------------ hourglass layers
 * Activations are NHWC bf16.  `stats` buffers are fp32 [N,G,2] (sum, sum of squares per sample and GroupNorm group),
 * accumulated atomically: zero them before the producing call.  Replaces nn.Conv2d / nn.GroupNorm / ReLU /
 * F.max_pool2d / F.interpolate of network/hourglass.py:7-41, 44-85, 88-173.
 *
 * sh_conv_fwd: stride-1 'same' convolution as an implicit GEMM on tcgen05 tensor cores.
 *   x bf16 [N,H,W,Cin] (Cin % 64 == 0; H, W powers of two >= 4); w bf16 [taps, cout_pad, Cin] (taps = 1 or 9, tap-major,
 *   rows >= Cout zero); bias fp32 [Cout] or NULL; residual bf16 [N,H,W,Cout] or NULL;
 *   y bf16 [N,H,W,y_ld] or NULL; y_nchw fp32 [N,Cout,H,W] or NULL; stats of the (rounded) output or NULL.
 *   The data gradient is the same call on dY with the flipped/transposed weights produced by sh_pack_weights. */
int sh_conv_fwd(const void* x, const void* w, const void* bias, const void* residual, int N, int H, int W, int Cin,
                int Cout, int cout_pad, int taps, void* y, int y_ld, void* y_nchw, void* stats, int groups,
                void* stream);
/* The same 1x1 convolution on a = relu(groupnorm(x)) with x the RAW GroupNorm input (gn_stats [N,gn_groups,2] = its per-(sample,
 * group) sum / sum of squares, gn_gamma / gn_beta [Cin]): GroupNorm + ReLU (network/hourglass.py:26-36) are applied to every
 * operand tile in shared memory between the TMA arrival and the MMA; bit-identical to sh_gn_relu_fwd followed by sh_conv_fwd.
 * Images of >= 64 pixels.  sh_conv_wgrad_gn: the matching weight gradient (x_C == Cin). */
int sh_conv_fwd_gn(const void* x, const void* gn_stats, const void* gn_gamma, const void* gn_beta, int gn_groups, float gn_eps,
                   const void* w, const void* bias, const void* residual, int N, int H, int W, int Cin, int Cout, int cout_pad,
                   void* y, int y_ld, void* y_nchw, void* stats, int groups, void* stream);
int sh_conv_wgrad_gn(const void* dy, const void* x, const void* gn_stats, const void* gn_gamma, const void* gn_beta, int gn_groups,
                     float gn_eps, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, void* dw, void* stream);
/* dw fp32 [Cout,Cin,k,k] (reference layout) += dY^T X (accumulated atomically: zero first).  dy bf16 [N,H,W,dy_C],
 * x bf16 [N,H,W,x_C]; x_C, dy_C multiples of 64; Cin <= x_C and Cout <= dy_C are the real channel counts. */
int sh_conv_wgrad(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout, int taps,
                  void* dw, void* stream);
/* 3x3 weight gradient accumulated into a [9][Cout][Cin] fp32 scratch (contiguous in Cin: vector reductions; zero it once per
 * step; W >= 16: kernel-row kernel, narrower images: per-tap kernel), and the one-launch transposition of every layer's scratch into the reference layout:
 * table int32 [n,4] (device) rows = (scratch offset, grad offset, Cout, Cin) in floats; grad[co][ci][kh][kw] += scratch[tap][co][ci]. */
int sh_conv_wgrad3x3(const void* dy, const void* x, int N, int H, int W, int x_C, int Cin, int dy_C, int Cout,
                     void* scratch, void* stream);
int sh_unpack_wgrad_batch(const void* table, int n, const void* scratch, void* grad, void* stream);
/* GroupNorm(G) + ReLU: y = relu(gn(x)); optional statistics of y for a following GroupNorm(G_out). */
int sh_gn_relu_fwd(const void* x, const void* stats_in, const void* gamma, const void* beta, int N, int HW, int C,
                   int G, float eps, void* y, void* stats_out, int G_out, void* stream);
/* Backward of the above in one pass over HBM: dx = d/dx (+ addend), dgamma/dbeta accumulated, optional column sum of dx
 * (bias gradient of the convolution that produced x).  red: scratch of sh_gn_relu_bwd_scratch_words(N,G) 4-byte words
 * (per-(n,g) partial sums + one arrive counter per sample; zeroed by the call).  HW*C <= 2M elements per sample. */
size_t sh_gn_relu_bwd_scratch_words(int N, int G);
int sh_gn_relu_bwd(const void* da, const void* x, const void* stats_in, const void* gamma, const void* beta,
                   const void* addend, int N, int HW, int C, int G, float eps, void* red, void* dgamma, void* dbeta,
                   void* dx, void* colsum, void* stream);
/* Same as sh_gn_relu_bwd for a `red` scratch the caller has already zeroed (one arena cleared once per backward pass). */
int sh_gn_relu_bwd_prezeroed(const void* da, const void* x, const void* stats_in, const void* gamma, const void* beta,
                   const void* addend, int N, int HW, int C, int G, float eps, void* red, void* dgamma, void* dbeta,
                   void* dx, void* colsum, void* stream);
int sh_maxpool_fwd(const void* x, int N, int H, int W, int C, void* y, void* stats_out, int G_out, void* stream);
int sh_maxpool_bwd(const void* dy, const void* x, const void* addend, int N, int H, int W, int C, void* dx,
                   void* colsum, void* stream);
int sh_upsample_add_fwd(const void* up1, const void* low, int N, int h, int w, int C, void* y, void* stats_out,
                        int G_out, void* stream);
int sh_upsample_bwd(const void* dy, int N, int h, int w, int C, void* dlow, void* colsum, void* stream);
int sh_add(const void* a, const void* b, const void* c, int N, int HW, int C, void* y, void* stats_out, int G_out,
           void* colsum, void* stream);
int sh_colsum(const void* x, int N, int HW, int C, void* colsum, void* stream);
/* Stem: 5x5 stride-2 convolution 1 -> 64 channels (network/hourglass.py:95-96).  img fp32 [N,S,S]; w fp32 [64,1,5,5]. */
int sh_stem_conv_fwd(const void* img, const void* w, const void* b, int N, int S, void* y, void* stats_out, int G_out,
                     void* stream);
int sh_stem_conv_wgrad(const void* img, const void* dy, int N, int S, void* dw, void* db, void* stream);
/* Layout conversion between the reference's fp32 NCHW tensors and the internal bf16 NHWC (channels padded to Cp). */
int sh_nchw_to_nhwc(const void* x, int N, int C, int HW, int Cp, void* y, void* stream);
int sh_nhwc_to_nchw(const void* x, int N, int C, int HW, void* y, void* stream);
/* fp32 master weight [Cout,Cin,k,k] -> bf16 forward layout wf [taps,cout_pad,cin_pad] and (optional) data-gradient
 * layout wb [taps,b_rows,b_cols] (flipped taps, transposed). */
int sh_pack_weights(const void* w, int Cout, int Cin, int taps, int cout_pad, int cin_pad, int b_rows, int b_cols,
                    void* wf, void* wb, void* stream);
/* The same for every convolution of a network in one launch.  flat: the flat fp32 parameter buffer; table int32 [n,10]
 * (device) rows = (src offset in flat, Cout, Cin, taps, cout_pad, cin_pad, b_rows, b_cols, wf offset, wb offset or -1),
 * offsets in elements; arena: bf16 buffer the wf/wb offsets index. */
int sh_pack_weights_batch(const void* flat, const void* table, int n, void* arena, void* stream);
int sh_unpack_wgrad(const void* dw, int Cout, int Cin, int taps, int cout_ld, int cin_ld, void* grad, void* stream);
/* torch.optim.Adam step with L2 weight decay on flat fp32 buffers (network/engine.py:95-97); grad is read as g*grad_scale. */
int sh_adam_step(void* p, const void* g, void* m, void* v, long n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, float grad_scale, void* stream);
/* Same update with the learning rate (fp32) and the step counter (int32, incremented by the call) in DEVICE memory,
 * so that a captured CUDA graph of the whole train step can be replayed while the StepLR schedule
 * (network/engine.py:98-99) moves the learning rate. */
int sh_adam_step_dev(void* p, const void* g, void* m, void* v, long n, const void* lr_dev, void* step_dev, float beta1,
                     float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------ train-step glue
 * Replaces the weighting / summation of MultiTaskLoss.forward (network/create_network_and_criterion.py:171-181,
 * 183-263) and Engine.sum_loss_terms (network/engine.py:144-148) for one stack output.
 * Rows n < Ns of xyz [Ns+M,J,3] are synthetic (target_xyz4 float4 [Ns,J]: 0.1*MSE on z), rows Ns.. are the M = B*V real
 * views.  g_mvproj [M,J,3], g_pose3 [3,M,J,3], g_prior [M,J*3] and the loss vectors are the outputs of the entries above
 * (any may be NULL = head disabled); sse2 from sh_softargmax_fwd.  weights8 (HOST pointer) = (synt_hm, synt_pt,
 * mv_projection, mv_consistency, hm_mean, prior, collision, bone_length).  mean_scale multiplies the gradient of every
 * batch-MEAN term (1/world_size under data parallelism; the batch-SUM terms, collision and the KLD, are not scaled).
 * Out: gxyz [Ns+M,J,3] = d(total)/d(xyz) (overwritten); terms9 += (synt_uv, synt_d, mv_projection, mv_consistency,
 * uv_hm_mean, pose_prior, collision, bone_length, total)  (accumulated across stacks: zero it once per step). */
int sh_step_combine(const void* g_mvproj, const void* g_pose3, const void* g_prior, const void* xyz,
                    const void* target_xyz4, const void* loss_mv3, const void* loss_pose3, const void* loss_prior3,
                    const void* sse2, int Ns, int M, int J, int hw, const float* weights8, float mean_scale, void* gxyz,
                    void* terms9, void* stream);
/* ------------------------------------------------------------------------------------------------ stand-alone module entries
 * What the fused step folds into larger kernels, exported one by one for the drop-in nn.Modules.
 * sh_data_to_model_fwdbwd replaces DataToModelLoss.forward (mesh/render.py:123-142) and its backward: dms [N,H,W] mm
 * (background > 99), joints [N,J,3], radii [J] -> loss[1], grad_joints [N,J,3] (upstream 1).  N <= 65535, J <= 64. */
size_t sh_data_to_model_scratch_bytes(int N, int J);
int sh_data_to_model_fwdbwd(const void* dms, const void* joints, const void* radii, int N, int J, int H, int W,
                            void* loss, void* grad_joints, void* scratch, void* stream);
/* OthographicalProjection.forward (mesh/pointTransformation.py:84-99; mode 1 = per-sample focal jitter rand_f [B], w := 1;
 * mode 2 = K-matrix) and InverseOthographicalProjection.forward (:118-124; mode 3).  points, out float4 [B,Nv]. */
int sh_ortho_project(const void* points, int B, int Nv, int mode, float cx, float cy, float fx, float fy,
                     const void* rand_f, void* out, void* stream);
/* RandScale.forward's matrix product (mesh/pointTransformation.py:144-148): out = diag(scales[b],1) * mats[b,m]. */
int sh_rand_scale_apply(const void* mats, const void* scales, int B, int nmat, void* out, void* stream);
/* torch.clamp(x, max=max_value) of DepthRasterizationFunction.forward (mesh/render.py:286); NaN propagates. */
int sh_clamp_max(const void* x, long n, float max_value, void* y, void* stream);

/* ResizeCropImage.forward (network/util_modules.py:388-424), all images in one launch: nearest-neighbour resize of image n by
 * (v_scales[n], u_scales[n]) pasted into the centre of an all-ones canvas.  depth_maps, out [N,H,W]. */
int sh_resize_crop(const void* depth_maps, const void* u_scales, const void* v_scales, int N, int H, int W, void* out,
                   void* stream);

/* xyz[m][j][0] /= u_scales[m], xyz[m][j][1] /= v_scales[m] in place on fp32 [M,J,3]: undoes the scale augmentation on the real
 * views' joints (network/create_network_and_criterion.py:124-126); applied to dL/dxyz it is that division's backward. */
int sh_unscale_xy(void* xyz, const void* u_scales, const void* v_scales, int M, int J, void* stream);

/* y = x * s on n fp32 elements: real_dms * depth_scale (network/engine.py:337), xyz / 100 (create_network_and_criterion.py:240). */
int sh_scale(const void* x, float s, long n, void* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPHEREHAND_B200_H */
