/*
 * pn2_b200.h -- C ABI of libpn2_b200.so, the sm_100a implementation of the
 * PointNet++ encoding hot path of YunzeMan/Situation3D (lib/pointnet2).
 *
 * Boundary.  Each entry point replaces one raw-pointer "kernel wrapper" of the
 * reference extension -- the layer directly under its ATen functions -- and is
 * what a binding for `pointnet2._ext` links against.  Reference interface
 * (paths relative to /root/reference/lib/pointnet2/_ext_src/):
 *
 *   pn2_gather_points            <- gather_points_kernel_wrapper            src/sampling.cpp:4-6,   src/sampling_gpu.cu:22-30
 *   pn2_gather_points_grad       <- gather_points_grad_kernel_wrapper       src/sampling.cpp:7-9,   src/sampling_gpu.cu:49-57
 *   pn2_furthest_point_sampling  <- furthest_point_sampling_kernel_wrapper  src/sampling.cpp:11-13, src/sampling_gpu.cu:175-229
 *   pn2_three_nn                 <- three_nn_kernel_wrapper                 src/interpolate.cpp:4-5,   src/interpolate_gpu.cu:61-68
 *   pn2_three_interpolate        <- three_interpolate_kernel_wrapper        src/interpolate.cpp:6-8,   src/interpolate_gpu.cu:103-111
 *   pn2_three_interpolate_grad   <- three_interpolate_grad_kernel_wrapper   src/interpolate.cpp:9-12,  src/interpolate_gpu.cu:145-154
 *   pn2_ball_query               <- query_ball_point_kernel_wrapper         src/ball_query.cpp:4-6,    src/ball_query_gpu.cu:46-54
 *   pn2_group_points             <- group_points_kernel_wrapper             src/group_points.cpp:4-6,  src/group_points_gpu.cu:30-39
 *   pn2_group_points_grad        <- group_points_grad_kernel_wrapper        src/group_points.cpp:8-10, src/group_points_gpu.cu:66-75
 *
 * and the fused entry points replace whole Python call chains of the
 * reference (cited at each declaration).
 *
 * Conventions.
 *   - Plain pointers and sizes only.  All pointers are device pointers on the
 *     current CUDA device unless a parameter says "host".
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream).  Every call is asynchronous on that stream, like the reference
 *     (which uses at::cuda::getCurrentCUDAStream()).
 *   - Inputs are borrowed and never written.  Outputs are caller-allocated.
 *     Functions that need scratch take a workspace pointer whose size comes
 *     from the matching *_workspace_bytes() query; nothing is allocated
 *     inside the library.
 *   - Return value: PN2_OK (0) or a negative PN2_ERR_* code.  The library
 *     never calls exit() (the reference's CUDA_CHECK_ERRORS does,
 *     include/cuda_utils.h:30-39).
 *   - Tensor layouts, dtypes and index semantics are the reference's:
 *     float32 / int32, row-major contiguous, indices bit-exact with the
 *     reference kernels.  All offsets are 64-bit.
 */
#ifndef PN2_B200_H_
#define PN2_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define PN2_OK 0
#define PN2_ERR_INVALID_ARGUMENT (-1)   /* bad dims / null pointer / unsupported shape */
#define PN2_ERR_CUDA (-2)               /* a CUDA call or launch failed; see pn2_last_cuda_error() */
#define PN2_ERR_WORKSPACE (-3)          /* workspace missing or too small */
#define PN2_ERR_UNSUPPORTED_DEVICE (-4) /* not an sm_100 device */

typedef void *pn2_stream_t;

/* ---- library ----------------------------------------------------------- */
int pn2_version(void);                       /* major*10000 + minor*100 + patch */
const char *pn2_error_string(int code);
const char *pn2_last_cuda_error(void);       /* text of the last CUDA failure on this thread */
int pn2_device_check(void);
/* kernels launched by this library since it was loaded (every entry point counts its own launches) */
unsigned long long pn2_launch_count(void);                  /* PN2_OK iff the current device is compute capability 10.x */

/* ---- the nine reference operators ------------------------------------- */

/* points (b,c,n) f32, idx (b,m) i32 -> out (b,c,m) f32 */
int pn2_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                      float *out, pn2_stream_t stream);

/* grad_out (b,c,m), idx (b,m) -> grad_points (b,c,n) += scatter (caller zero-fills, as the
 * reference allocates with torch::zeros, sampling.cpp:51-53) */
int pn2_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                           float *grad_points, pn2_stream_t stream);

/* xyz (b,n,3) f32 -> idxs (b,m) i32.  workspace: pn2_furthest_point_sampling_workspace_bytes
 * bytes (may be 0, then workspace may be NULL); it replaces the reference's `temp` (b,n)
 * scratch (sampling.cpp:74-76) and needs no initialisation. */
size_t pn2_furthest_point_sampling_workspace_bytes(int b, int n, int m);
int pn2_furthest_point_sampling(int b, int n, int m, const float *xyz, void *workspace,
                                size_t workspace_bytes, int *idxs, pn2_stream_t stream);

/* unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) f32 (squared), idx (b,n,3) i32 */
int pn2_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, pn2_stream_t stream);

/* points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n) */
int pn2_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, pn2_stream_t stream);

/* grad_out (b,c,n), idx, weight -> grad_points (b,c,m) += scatter (caller zero-fills) */
int pn2_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                               const int *idx, const float *weight, float *grad_points,
                               pn2_stream_t stream);

/* new_xyz (b,m,3) centres, xyz (b,n,3) -> idx (b,m,nsample) i32; all slots are written
 * (zeros when a ball is empty, as the reference's zero-filled output). */
int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, pn2_stream_t stream);

/* Same result, bit for bit, through a uniform grid when the scene is large (n >= 4096): the points are
 * binned into cells of edge >= 1.001*radius and each centre tests only the 27 surrounding cells; the
 * ascending-index order of the reference is restored with a per-warp bitmap.  workspace:
 * pn2_ball_query_workspace_bytes() bytes (0 for small scenes, then the plain scan runs). */
size_t pn2_ball_query_workspace_bytes(int b, int n, int m, int nsample);
int pn2_ball_query_ws(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                      const float *xyz, int *idx, void *workspace, size_t workspace_bytes,
                      pn2_stream_t stream);

/* The two halves of pn2_ball_query_ws.  The grid depends on the points and the radius only, so a caller
 * whose centres come from a sampling kernel on another stream (the backbone's SA1: lib/pointnet2/
 * pointnet2_modules.py:181-205 runs FPS, then QueryAndGroup) builds it while that kernel runs.
 * xyz_pitch = floats between consecutive points (3 for (b,n,3), 3+C for point_clouds rows).
 * Both return PN2_ERR_WORKSPACE when the shape takes the plain scan (workspace_bytes() == 0) or the
 * workspace is missing/short; the same workspace, untouched in between, must be passed to both. */
int pn2_ball_query_grid_build(int b, int n, int m, float radius, int nsample, const float *xyz,
                              int xyz_pitch, void *workspace, size_t workspace_bytes, pn2_stream_t stream);
int pn2_ball_query_grid_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              int *idx, void *workspace, size_t workspace_bytes, pn2_stream_t stream);

/* points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample) */
int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, pn2_stream_t stream);

/* grad_out (b,c,npoints,nsample), idx -> grad_points (b,c,n) += scatter (caller zero-fills) */
int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, pn2_stream_t stream);

/* Same as pn2_furthest_point_sampling, and also writes the sampled coordinates
 * new_xyz (b,m,3) = xyz[idxs] (may be NULL).  Fuses the gather_operation + two transposes
 * of PointnetSAModuleVotes.forward (pointnet2_modules.py:233-240) into the sampling kernel:
 * every CTA already holds the winner's coordinates each round. */
int pn2_furthest_point_sampling_xyz(int b, int n, int m, const float *xyz, int *idxs,
                                    float *new_xyz, pn2_stream_t stream);
/* Same over rows of `pitch` >= 3 floats whose first three are x, y, z (the (b, n, 3+C) point_clouds tensor read in
 * place); xyz_copy (b,n,3) or NULL receives the contiguous coordinates (written once while the points are loaded). */
int pn2_furthest_point_sampling_rows(int b, int n, int m, const float *rows, int pitch, int *idxs,
                                     float *new_xyz, float *xyz_copy, pn2_stream_t stream);

/* Same two entry points with a workspace of pn2_furthest_point_sampling_workspace_bytes[_mode](b, n, m) bytes (16-byte
 * aligned device memory, no initialisation needed).  With a workspace the bucketed kernel runs (csrc/fps_bucket.cu: one
 * CTA per scene, points binned into spatial buckets, the distance update pruned exactly by bounding boxes); without
 * one (the size query returned 0) the register-resident kernels of csrc/fps.cu run.  Results are identical either
 * way; the two differ in what they optimise.  PN2_FPS_LATENCY (what the plain size query answers for): the fastest
 * single launch -- the cluster kernel up to 81 920 points.  PN2_FPS_THROUGHPUT: the least SM time per scene, for callers
 * that keep enough batches in flight to fill the GPU (>= ~150 scenes) -- the bucketed kernel from 32 768 points. */
#define PN2_FPS_LATENCY 0
#define PN2_FPS_THROUGHPUT 1
size_t pn2_furthest_point_sampling_workspace_bytes_mode(int b, int n, int m, int mode);
int pn2_furthest_point_sampling_xyz_ws(int b, int n, int m, const float *xyz, int *idxs, float *new_xyz,
                                       void *workspace, size_t workspace_bytes, pn2_stream_t stream);
int pn2_furthest_point_sampling_rows_ws(int b, int n, int m, const float *rows, int pitch, int *idxs,
                                        float *new_xyz, float *xyz_copy, void *workspace, size_t workspace_bytes,
                                        pn2_stream_t stream);
/* Diagnostic for the bucketed kernel: prof (device, 6 x int64) = SM cycles of thread 0 of CTA 0 spent in
 * {bucket tests, bucket updates, thread/warp argmax, barrier, table argmax, -}, summed over the rounds. */
int pn2_debug_fps_bucket_profile(int b, int n, int m, const float *xyz, int *idxs, long long *prof,
                                 void *workspace, size_t workspace_bytes, pn2_stream_t stream);

/* Diagnostic: same launch as pn2_furthest_point_sampling; prof (device, 5 x int64) receives the SM cycles
 * thread 0 of CTA 0 spent per phase of a round, summed over the m-1 rounds. */
int pn2_debug_fps_profile(int b, int n, int m, const float *xyz, int *idxs, long long *prof,
                          pn2_stream_t stream);

/* ---- fused SharedMLP layers, full precision ------------------------------
 * A SharedMLP (pytorch_utils.py:11-36: 1x1 Conv2d without bias -> BatchNorm2d -> ReLU per
 * layer) in eval mode is described by `nlayers` and host array dims[nlayers+1]
 * (dims[0] = input width) plus, per layer, the BatchNorm-folded weight (cout,cin) row-major
 * and bias (cout).  pn2_mlp_f32_pack turns them into one device image
 * (transposed, padded) of pn2_mlp_f32_image_bytes() bytes that the fused kernels consume.
 * w / bias are HOST arrays of nlayers DEVICE pointers; bias or bias[l] may be NULL (zeros).
 * At most 4 layers; pn2_mlp_f32_supported() says whether the widths fit shared memory.
 */
size_t pn2_mlp_f32_image_bytes(int nlayers, const int *dims);
int pn2_mlp_f32_supported(int nlayers, const int *dims);
int pn2_mlp_f32_pack(int nlayers, const int *dims, const float *const *w, const float *const *bias,
                     void *image, pn2_stream_t stream);

/* (b,c,n) channel-first -> (b,n,c) channel-last rows: the layout the fused kernels gather
 * from (one contiguous row per point instead of the reference's stride-N reads,
 * group_points_gpu.cu:23-26). */
int pn2_rows_from_channels(int b, int c, int n, const float *src, float *dst, pn2_stream_t stream);

/* Fused set-abstraction layer, fp32.  Replaces, for eval-mode
 * PointnetSAModuleVotes(pooling='max'), the chain
 *   QueryAndGroup.forward   pointnet2_utils.py:348-359 (group xyz, recentre, /radius, group feats, cat)
 *   SharedMLP               pytorch_utils.py:11-36,67-121
 *   F.max_pool2d            pointnet2_modules.py:259-262
 * given the ball-query indices idx (b,npoint,nsample).
 *   table: channel-last feature rows, row (s,p) at table + (s*n + p)*ld, first c floats used
 *          (for the backbone input point_clouds (b,n,3+c): table = pc + 3, ld = 3 + c);
 *   use_xyz: prepend (xyz[idx] - new_xyz) * inv_radius (inv_radius = 1/radius when
 *          normalize_xyz, else 1); dims[0] must equal c + 3*use_xyz;
 *   out (b,cout,npoint) f32; out_rows (b,npoint,cout) f32 or NULL (next layer's table).
 */
int pn2_sa_forward_f32(int b, int n, int npoint, int nsample, int c, const float *table, int ld,
                       int use_xyz, float inv_radius, const float *xyz, const float *new_xyz,
                       const int *idx, int nlayers, const int *dims, const void *image, float *out,
                       float *out_rows, pn2_stream_t stream);

/* ---- fused set-abstraction layer on tcgen05 tensor cores (bf16 operands, fp32 accumulate) ----
 * Same chain as pn2_sa_forward_f32 (QueryAndGroup gather -> 3-layer SharedMLP -> max-pool) with the
 * three 1x1-conv layers as tcgen05.mma tiles whose accumulators live in TMEM.
 *   table: bf16 channel-last rows (b, n, pn2_sa_tc_row_elems(c)), made by pn2_sa_tc_pack_rows (from
 *          channel-last f32, e.g. point_clouds + 3 with src_ld = 3 + c and src_batch_stride = n*src_ld)
 *          or pn2_sa_tc_pack_channels (from the reference's (b,c,n) layout), or produced as out_table
 *          by the previous layer;
 *   weight image: pn2_sa_tc_pack_weights from the BatchNorm-folded w1 (c1, 3+c) [xyz columns first, the
 *          reference's channel order], w2 (c2,c1), w3 (c3,c2) and biases (device pointers, f32);
 *   out (b,c3,npoint) f32; out_table (b,npoint,c3) bf16 or NULL.
 * Supported shapes (pn2_sa_tc_supported): use_xyz, exactly three layers, c1,c2 multiples of 16 up to
 * 256, c3 = 128 or 256, nsample in {16,32,64,128}, npoint*nsample a multiple of 128.  Other shapes
 * run through pn2_sa_forward_f32. */
int pn2_sa_tc_row_elems(int c);
int pn2_sa_tc_supported(int c, int c1, int c2, int c3, int npoint, int nsample);
size_t pn2_sa_tc_weight_image_bytes(int c, int c1, int c2, int c3);
int pn2_sa_tc_pack_weights(int c, int c1, int c2, int c3, const float *w1, const float *b1,
                           const float *w2, const float *b2, const float *w3, const float *b3,
                           void *image, pn2_stream_t stream);
int pn2_sa_tc_pack_rows(int b, int n, int c, const float *src, long long src_batch_stride, int src_ld,
                        void *table, pn2_stream_t stream);
int pn2_sa_tc_pack_channels(int b, int c, int n, const float *features, void *table, pn2_stream_t stream);
int pn2_sa_tc_forward(int b, int n, int npoint, int nsample, int c, int c1, int c2, int c3,
                      float inv_radius, const float *xyz, const float *new_xyz, const void *table,
                      const int *idx, const void *weight_image, float *out, void *out_table,
                      pn2_stream_t stream);
/* ---- per-point half of an SA layer's first 1x1 convolution (csrc/lin_tc.cu) ---------------------
 * Layer 1 of the SharedMLP is linear before its ReLU and only its three xyz input channels depend on the
 * centre (pointnet2_utils.py:348-359 concatenates [xyz - centre ; features], pytorch_utils.py:11-36
 * convolves): W1 [dxyz ; f] = W1[:, :3] dxyz + W1[:, 3:] f.  pn2_lin_tc_forward computes
 * P = X W^T (bf16 out, f32 accumulate) once per point; pn2_sa_tc_forward is then run over the c1-wide P rows
 * with a layer-1 weight [W1[:, :3] | I].  x: `rows` channel-last rows of pitch ld elements, f32 (16-byte
 * aligned base, ld % 4 == 0; the first `skip` elements of a row get zero weights, so point_clouds can be
 * read in place with skip = 3) or bf16 (ld % 8 == 0); kin = elements read per row.  c1 = 64 or 128. */
int pn2_lin_tc_supported(int kin, int c1, int x_is_bf16);
size_t pn2_lin_tc_weight_image_bytes(int kin, int c1);
int pn2_lin_tc_pack_weights(int kin, int skip, int c, int c1, const float *w, int w_ld, int w_col0,
                            void *image, pn2_stream_t stream);
int pn2_lin_tc_forward(long long rows, int kin, int c1, const void *x, int x_is_bf16, int ld,
                       const void *weight_image, void *out, pn2_stream_t stream);

/* ---- fused feature-propagation layer on tcgen05 tensor cores ------------------------------------
 * Same chain as pn2_fp_forward_f32 (3-NN weights -> three_interpolate -> cat(skip) -> 2-layer SharedMLP),
 * inputs and outputs as bf16 channel-last rows: known_rows (b,m,c_known), skip_rows (b,n,c_skip),
 * out (b,c2,n) f32, out_rows (b,n,c2) bf16 or NULL.  The weight image is built from the folded
 * w1 (c1, c_known+c_skip) [interpolated channels first, as the reference concatenates], w2 (c2,c1).
 * Supported: c_known, c_skip, c1 multiples of 64, c2 multiple of 16, c1,c2 <= 256. */
int pn2_fp_tc_supported(int c_known, int c_skip, int c1, int c2);
size_t pn2_fp_tc_weight_image_bytes(int c_known, int c_skip, int c1, int c2);
int pn2_fp_tc_pack_weights(int c_known, int c_skip, int c1, int c2, const float *w1, const float *b1,
                           const float *w2, const float *b2, void *image, pn2_stream_t stream);
int pn2_fp_tc_forward(int b, int n, int m, int c_known, int c_skip, int c1, int c2, const float *dist2,
                      const int *idx, const void *known_rows, const void *skip_rows,
                      const void *weight_image, float *out, void *out_rows, pn2_stream_t stream);

/* Second-generation fused FP layer (csrc/fp_tc2.cu): three_nn (interpolate_gpu.cu:9-59 + the sqrt of
 * pointnet2_utils.py:142) included, so it takes the coordinates instead of dist2 / idx: unknown (b,n,3), known (b,m,3)
 * f32.  One launch per layer; a 4-CTA cluster owns a 128-point tile and splits the output channels of both MLP layers,
 * each CTA keeping its weight quarter resident in shared memory.  Same weight image, same rows, same outputs as
 * pn2_fp_tc_forward (bit-identical indices and weights; the bf16 rounding points are the same).
 * Supported: c_known, c_skip multiples of 64 (c_skip may be 0), c1, c2 in {64, 128, 192, 256}, m * 12 bytes <= the
 * operand buffer (m <= 10922 for the backbone's widths). */
int pn2_fp_tc2_supported(int c_known, int c_skip, int c1, int c2, int m);
int pn2_fp_tc2_forward(int b, int n, int m, int c_known, int c_skip, int c1, int c2, const float *unknown,
                       const float *known, const void *known_rows, const void *skip_rows, const void *weight_image,
                       float *out, void *out_rows, pn2_stream_t stream);

/* Diagnostic: while prof (device, 16 x int64) is non-NULL, pn2_fp_tc2_forward records the SM clock of thread 0 of
 * CTA 0 at its phase boundaries (csrc/fp_tc2.cu). */
int pn2_debug_fp_tc2_profile(long long *prof);

/* ---- the linear + GELU that consumes the visual tokens (csrc/head_tc.cu) -------
 * SIG3D.scene_feat_linear = Sequential(Linear(256, 768), GELU()), situation3d/models/sqa_module.py:180-183,344:
 * out (rows, n) f32 = GELU_erf(x (rows, k) f32 . W^T + bias) as one tcgen05 kernel (bf16 operands, fp32 accumulate).
 * W (n, k) f32 -> image by pn2_linear_gelu_tc_pack_weights.  k multiple of 64, n multiple of 128. */
int pn2_linear_gelu_tc_supported(int k, int n);
size_t pn2_linear_gelu_tc_weight_image_bytes(int k, int n);
int pn2_linear_gelu_tc_pack_weights(int k, int n, const float *w, void *image, pn2_stream_t stream);
int pn2_linear_gelu_tc_forward(long long rows, int k, int n, const float *x, const void *weight_image,
                               const float *bias, float *out, pn2_stream_t stream);

/* ---- visual-token construction in front of the re-encoding (csrc/tokens.cu) ----
 * Replaces the per-scene loop of SIG3D.forward, situation3d/models/sqa_module.py:297-315.
 * pn2_column_pool: scene s owns voxels offsets[s] .. offsets[s+1] of coords (M,3) i32 / feats (M,c) f32 (all
 *   device pointers, offsets (b+1) i32).  Its unique (x,y) columns, in the ascending order of
 *   torch.unique(dim=0) (:298-299), go to rows offsets[s] .. offsets[s] + ncols[s] of out_coords (M,2) /
 *   out_feats (M,c), with out_feats = sum over the column's voxels / (count + 1) -- scatter_reduce_('mean') onto a
 *   zero tensor counts the zero (:300-301).  inverse (M) (may be NULL) = return_inverse.  At most
 *   pn2_column_pool_max_voxels() voxels per scene, |x|,|y| < 2^19 (else *status = 1).
 * pn2_token_gather: tokens[s,j,:] = pooled[offsets[s] + sampled[s,j], :], positions[s,j,:] =
 *   (pooled_coords + stride/2) * voxel_size in fp32 as torch evaluates :311; the caller draws `sampled` (:303-308). */
int pn2_column_pool_max_voxels(void);
int pn2_column_pool(int b, const int *offsets, const int *coords, const float *feats, int c, int *ncols,
                    int *out_coords, float *out_feats, int *inverse, int *status, pn2_stream_t stream);
int pn2_token_gather(int b, int t, int c, const int *offsets, const int *sampled, const float *pooled,
                     const int *pooled_coords, float half_x, float half_y, float voxel_size, float *tokens,
                     float *positions, pn2_stream_t stream);

/* ---- positional embedding looked up by integer voxel coordinate (csrc/voxel_pe.cu) ----
 * Replaces the per-sample CPU loop of 3D-LLM's Blip2T5.forward / predict_answers
 * (3DLLM_BLIP2-base/lavis/models/blip2_models/blip2_t5.py:104-118, :279-293) and Blip2OPT.forward
 * (blip2_opt.py:92-104): pe[s,j, a*seg + k] = table[coords[s,j,a], k] for the axes a = 0,1,2 (channels
 * 3*seg .. c-1 stay zero; the reference has seg = 469 = 1408 // 3, c = 1408, a (256, 469) table), then
 *   PN2_VOXEL_PE_ADD: out (b,p,c)  = feat + scale * pe, the product rounded first as torch does (blip2_t5.py:118);
 *   PN2_VOXEL_PE_CAT: out (b,2p,c) = cat([feat, pe], 1) (blip2_opt.py:104); feat may be NULL when rows [0,p) of
 *                     every sample already hold the features -- then only the pe rows are written.
 * coords: (b,p,coord_stride >= 3) device array of kind PN2_COORD_*; F32 is truncated toward zero like `.long()`
 * (:107).  Indexing follows torch: a negative index counts from the end of the table; anything outside
 * [-table_rows, table_rows) is an IndexError in the reference -- here *status (device int) becomes 1, the row is
 * clamped, and the caller decides when to read the flag.  table (table_rows, seg) f32, feat/out f32 contiguous. */
#define PN2_VOXEL_PE_ADD 0
#define PN2_VOXEL_PE_CAT 1
#define PN2_COORD_I32 0
#define PN2_COORD_I64 1
#define PN2_COORD_F32 2
int pn2_voxel_pe(int b, int p, int c, int seg, int table_rows, int coord_kind, int coord_stride, const void *coords,
                 const float *table, const float *feat, float *out, float scale, int mode, int *status,
                 pn2_stream_t stream);

/* ---- point <-> pixel correspondences and back-projection of image features (csrc/projection.cu)
 * Replaces ProjectionHelper.compute_projection / project of lib/projection.py (:191-254, :257-279) for all camera
 * views of a scene at once.  Host arrays: intrinsic4 = {intrinsic[0][0], [1][1], [0][2], [1][2]}, depth_range3 =
 * {depth_min, depth_max, accuracy}, corner_points (8,3) = ProjectionHelper._compute_corner_points (:29-46);
 * width / height = image_dims[0] / [1].  Device arrays: points (n,3) f32, depth (views, height, width) f32,
 * camera_to_world / world_to_camera (views,4,4) f32 row-major (the reference calls torch.inverse, :203; the caller
 * supplies the inverse here).
 * pn2_frustum_planes: corners (views,8,4) = compute_frustum_corners (:48-70), normals (views,6,3) =
 *   compute_frustum_normals (:72-119); either may be NULL.
 * pn2_points_in_frustum: points_in_frustum (:121-155) for the caller's corners (8,4) / normals (6,3) (device):
 *   mask (n) u8 (may be NULL) and *count (device int) = number of points inside.
 * pn2_compute_projection: per view v, indices_3d[v] / indices_2d[v] are the reference's (n+1) int64 arrays --
 *   element 0 the number of correspondences, then the point indices in ascending order / their pixel indices
 *   y * width + x, zero-padded (:246-252).  A view for which the reference returns None (:217,232,241) has count 0.
 *   counts (views) i32 may be NULL.  A point corresponds to a pixel iff it passes the six half-space tests
 *   round(dot * 100) / 100 < 0 (:143-146), projects to a pixel inside the image (:226-231) and the depth map agrees:
 *   depth_min <= d <= depth_max and |d - z_camera| <= accuracy (:239-240).
 * pn2_project: out (views, c, n) f32 = zeros; out[v, :, indices_3d[v][1+k]] = label[v, :, indices_2d[v][1+k]]
 *   (:271-277); label (views, c, hw).  An index outside [0,n) / [0,hw) is an error in the reference: *status = 1.
 * Workspaces are device memory owned by the caller (pn2_*_workspace_bytes). */
int pn2_frustum_planes(int views, const float *camera_to_world, const float *intrinsic4, const float *depth_range3,
                       int width, int height, const float *corner_points, float *corners, float *normals,
                       pn2_stream_t stream);
int pn2_points_in_frustum(int n, const float *points, const float *corners, const float *normals,
                          unsigned char *mask, int *count, pn2_stream_t stream);
size_t pn2_compute_projection_workspace_bytes(int views, int n);
int pn2_compute_projection(int views, int n, const float *points, const float *depth, const float *camera_to_world,
                           const float *world_to_camera, const float *intrinsic4, const float *depth_range3,
                           int width, int height, const float *corner_points, long long *indices_3d,
                           long long *indices_2d, int *counts, void *workspace, size_t workspace_bytes,
                           pn2_stream_t stream);
size_t pn2_project_workspace_bytes(int views, int n);
int pn2_project(int views, int c, int hw, int n, const float *label, const long long *indices_3d,
                const long long *indices_2d, float *out, int *status, void *workspace, size_t workspace_bytes,
                pn2_stream_t stream);

/* Across-view max-pooling of the back-projected features, the last step of the multiview-feature production: out[pt, ch]
 * = max(0, max over views v with a correspondence for pt of label[v, ch, pixel(v, pt)]) -- the per-frame project()
 * (lib/projection.py:257-279) folded with the running element-wise max into a zero-initialised per-point array that
 * yields "enet_feats_maxpool" (lib/config.py:36).  Same lists and label layout as pn2_project; out is (n, c) when
 * rows_layout != 0 (the channel-last rows the backbone reads), else (c, n).  Workspace: pn2_project_workspace_bytes. */
int pn2_project_maxpool(int views, int c, int hw, int n, const float *label, const long long *indices_3d,
                        const long long *indices_2d, float *out, int rows_layout, int *status, void *workspace,
                        size_t workspace_bytes, pn2_stream_t stream);

/* ---- SM partitions for callers that keep several batches in flight (csrc/sm_partition.cu) ----------
 * The sampling chain of a batch is a latency-bound kernel of half-SM CTAs that lives ~2 ms; the fused MLP
 * kernels are persistent whole-SM CTAs.  pn2_sm_partition_create splits the current device's SMs into a first
 * group of at least sms_first SMs and the rest (CUDA green contexts); streams created on a group confine their
 * kernels to it, and the persistent kernels of this library size their grids from the stream's group
 * (pn2_stream_sm_count).  Purely a scheduling aid: results do not depend on it. */
int pn2_sm_partition_create(int sms_first, void **handle);
int pn2_sm_partition_sms(void *handle, int which);
int pn2_sm_partition_stream_create(void *handle, int which, void **stream);
int pn2_stream_sm_count(pn2_stream_t stream);

/* Diagnostic: while prof (device, 32 x int64) is non-NULL, the pn2_sa_tc_forward kernels record the SM cycles
 * CTA 0 spends per phase.  Software-pipelined kernel: [0..5] epilogue warp 0 (wait D1, E1, wait D2, E2, wait D3,
 * E3), [7] tiles, [8..10] layer-1 issuer, [11..14] layer-2/3 issuer, [16..18] gather warps. */
int pn2_debug_sa_tc_profile(long long *prof);

/* Diagnostic: D (128 x n f32) = A (128 x k bf16) * B^T (n x k bf16) through the same shared-memory
 * layouts, descriptors and TMEM path as the fused kernel (one tile).  n, k multiples of 16. */
int pn2_selftest_umma(int n, int k, const void *a, const void *b, float *d, pn2_stream_t stream);

/* Fused feature-propagation layer, fp32.  Replaces PointnetFPModule.forward
 * (pointnet2_modules.py:399-421) in eval mode after three_nn: sqrt -> 1/(d+1e-8) ->
 * normalise -> three_interpolate -> cat(skip) -> SharedMLP.
 *   dist2, idx (b,n,3) from pn2_three_nn; known_rows (b,m,c_known), skip_rows (b,n,c_skip)
 *   channel-last (skip_rows may be NULL when c_skip == 0); dims[0] = c_known + c_skip;
 *   out (b,cout,n); out_rows (b,n,cout) or NULL.
 */
int pn2_fp_forward_f32(int b, int n, int m, int c_known, int c_skip, const float *dist2,
                       const int *idx, const float *known_rows, const float *skip_rows, int nlayers,
                       const int *dims, const void *image, float *out, float *out_rows,
                       pn2_stream_t stream);

/* ---- situation-conditioned re-encoding --------------------------------
 * tokens (b,t,d) f32, positions (b,t,3) f32, situation (b,7) f32 = (tx,ty,tz,qx,qy,qz,qw).
 * p' = R(q) p + t in the form of situation3d/utils/temp.py:42-80 (mode 0) or the inverse
 * R^T (p - t) (mode 1); pe = W2 gelu(W1 p'_xy + b1) + b2 (sqa_module.py:274-278,319-321);
 * out = tokens + pe; prior (b,t) = normalised exp(-|p_xy - t_xy|^2 / (2 sigma^2))
 * (sqa_module.py:328-336) of the untransformed positions; new_pos (b,t,3) = p'.
 * w1 (h,2), b1 (h), w2 (d,h), b2 (d) f32 (nn.Linear layout).  prior/new_pos may be NULL.
 */
int pn2_reencode_forward(int b, int t, int d, int h, int mode, float sigma, const float *tokens,
                         const float *positions, const float *situation, const float *w1,
                         const float *b1, const float *w2, const float *b2, float *out,
                         float *new_pos, float *prior, pn2_stream_t stream);

/* quats (b,4) xyzw -> (b,3,3): quaternions_to_rotation_matrices, sqa_module.py:12-30 */
int pn2_quaternions_to_rotation_matrices(int b, const float *quats, float *out, pn2_stream_t stream);
/* rotvecs (b,3) -> (b,3,3): batch_rotation_vector_to_matrix, sqa_module.py:33-64 */
int pn2_rotation_vectors_to_matrices(int b, const float *rotvecs, float *out, pn2_stream_t stream);
/* situation (b,7) -> (b,4,4): batch_matrix_function, situation3d/utils/temp.py:42-80 */
int pn2_situation_matrices(int b, const float *situation, float *out, pn2_stream_t stream);

#if defined(__GNUC__)
#endif
/* ---- training-mode BatchNorm + ReLU (+ max-pool over nsample) on channel-last rows -------------------------------
 * Replaces, for the training step (BASELINE.json config 4), the element-wise chain the reference runs per SharedMLP
 * layer -- BatchNorm2d (batch statistics, biased variance, eps inside the root) -> ReLU, lib/pointnet2/pytorch_utils.py:
 * 11-36,67-121 -- and the F.max_pool2d over nsample that follows the last layer (lib/pointnet2/pointnet2_modules.py:
 * 259-262), forward and backward.  x, y, dy, dx: (rows, c) fp32 row-major, c % 4 == 0 and 256 % (c/4) == 0
 * (pn2_rows_bn_supported); a, b, k1, k2, k3: (c,) fp32.
 *   stats:            partials[cta] = [sum x | sum x^2] over the CTA's rows (no atomics: deterministic); *nparts slots
 *   finalize:         mean / biased variance -> a = gamma/sqrt(var+eps), b = beta - mean*a, stat = [mean | 1/sqrt(var+eps)]
 *                     (fp64, for the backward pass); running_mean / running_var (may be NULL) updated as nn.BatchNorm2d
 *                     does (momentum, unbiased variance)
 *   apply:            y = max(x*a + b, 0)
 *   pool:             pooled[g] = max_j y[g*nsample + j], arg[g] = first j attaining it (u8); y itself is not written
 *   bwd_reduce:       partials[cta] = [sum g | sum g*x],  g = dy where x*a + b > 0 else 0
 *   bwd_finalize:     dgamma, dbeta and k1..k3 of dx = k1*g + k2 + k3*x (BatchNorm's input gradient)
 *   bwd_apply:        dx = k1*g + k2 + k3*x
 *   pool_bwd_*:       the same with g non-zero only at row arg[g] of each group, read from dpooled (groups, c)
 * partials: pn2_rows_bn_partials_bytes(rows, c) bytes of device memory.                                              */
int pn2_rows_bn_supported(long long rows, int c, int nsample);
size_t pn2_rows_bn_partials_bytes(long long rows, int c);
int pn2_rows_bn_stats(long long rows, int c, const float *x, double *partials, int *nparts, pn2_stream_t stream);
int pn2_rows_bn_finalize(int c, int nparts, const double *partials, long long rows, double eps, const float *weight,
                         const float *bias, float momentum, float *running_mean, float *running_var, float *a, float *b,
                         double *stat, pn2_stream_t stream);
int pn2_rows_bn_bwd_finalize(int c, int nparts, const double *partials, long long rows, const double *stat,
                             const float *weight, float *k1, float *k2, float *k3, float *dgamma, float *dbeta,
                             pn2_stream_t stream);
int pn2_rows_bn_relu_apply(long long rows, int c, const float *x, const float *a, const float *b, float *y,
                           pn2_stream_t stream);
int pn2_rows_bn_relu_pool(long long groups, int nsample, int c, const float *x, const float *a, const float *b,
                          float *pooled, unsigned char *arg, pn2_stream_t stream);
int pn2_rows_bn_relu_bwd_reduce(long long rows, int c, const float *dy, const float *x, const float *a, const float *b,
                                double *partials, int *nparts, pn2_stream_t stream);
int pn2_rows_bn_relu_bwd_apply(long long rows, int c, const float *dy, const float *x, const float *a, const float *b,
                               const float *k1, const float *k2, const float *k3, float *dx, pn2_stream_t stream);
int pn2_rows_bn_relu_pool_bwd_reduce(long long groups, int nsample, int c, const float *dpooled, const float *x,
                                     const unsigned char *arg, const float *a, const float *b, double *partials,
                                     int *nparts, pn2_stream_t stream);
int pn2_rows_bn_relu_pool_bwd_apply(long long groups, int nsample, int c, const float *dpooled, const float *x,
                                    const unsigned char *arg, const float *a, const float *b, const float *k1,
                                    const float *k2, const float *k3, float *dx, pn2_stream_t stream);

/* The grouped input matrix of an SA module in the row layout, one pass: out (b*npoint*nsample, c+3) =
 * [ (xyz[idx] - centre) (/ radius if normalize_xyz) | rows[idx] ], xyz channels first -- QueryAndGroup's group(xyz), -=, /=,
 * group(features), cat (lib/pointnet2/pointnet2_utils.py:348-359).  idx (b, npoint, nsample) scene-local; rows (b, n, ld). */
int pn2_group_rows(int b, int n, int npoint, int nsample, int c, int ld, const int *idx, const float *xyz,
                   const float *new_xyz, const float *rows, float radius, int normalize_xyz, float *out,
                   pn2_stream_t stream);

#pragma GCC visibility pop

#ifdef __cplusplus
}
#endif
#endif /* PN2_B200_H_ */
