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
int pn2_device_check(void);                  /* PN2_OK iff the current device is compute capability 10.x */

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

/* points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample) */
int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, pn2_stream_t stream);

/* grad_out (b,c,npoints,nsample), idx -> grad_points (b,c,n) += scatter (caller zero-fills) */
int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, pn2_stream_t stream);

/* ---- fused set-abstraction layer --------------------------------------
 * Replaces, for eval-mode PointnetSAModuleVotes(use_xyz=True, pooling='max'), the chain
 *   QueryAndGroup.forward           pointnet2_utils.py:348-359 (group xyz, recentre, /radius, group feats, cat)
 *   SharedMLP (3x Conv2d 1x1 + BN + ReLU)   pytorch_utils.py:11-36,67-121
 *   F.max_pool2d over nsample       pointnet2_modules.py:259-262
 * given the ball-query indices.  BatchNorm is folded into the weights by the caller
 * (SURVEY.md A.5); the MLP runs on tcgen05 tensor cores with bf16 operands and fp32
 * accumulation in TMEM.
 */

/* Pack (b,c,n) f32 channel-first features into the channel-last bf16 row table the fused
 * kernel gathers from: table (b, n, row_elems) bf16, row = [feat_0..feat_{c-1}, 0 ...].
 * row_elems = pn2_sa_row_elems(c).  features may be NULL when c == 0. */
int pn2_sa_row_elems(int c);
int pn2_sa_pack_features(int b, int c, int n, const float *features, void *table,
                         pn2_stream_t stream);
/* Same, from a channel-last f32 source with row stride `src_stride` floats whose first
 * `skip` floats are not features (the backbone's point_clouds (b,n,3+c) input). */
int pn2_sa_pack_features_cl(int b, int c, int n, const float *src, int src_stride, int skip,
                            void *table, pn2_stream_t stream);

/* Folded weights -> device image consumed by pn2_sa_forward.
 *   w1 (c1, 3+c) f32 row-major with the reference channel order [dx,dy,dz,feat...],
 *   w2 (c2, c1), w3 (c3, c2); b1,b2,b3 folded biases.  image: pn2_sa_weight_image_bytes(). */
size_t pn2_sa_weight_image_bytes(int c, int c1, int c2, int c3);
int pn2_sa_pack_weights(int c, int c1, int c2, int c3, const float *w1, const float *b1,
                        const float *w2, const float *b2, const float *w3, const float *b3,
                        void *image, pn2_stream_t stream);

/* xyz (b,n,3) f32, new_xyz (b,npoint,3) f32, table from pn2_sa_pack_features, idx
 * (b,npoint,nsample) i32 -> out (b,c3,npoint) f32 and, if out_table != NULL, the same values
 * as the next layer's bf16 row table (b,npoint,pn2_sa_row_elems(c3)).
 * inv_radius = 1/radius when normalize_xyz, else 1.  Returns PN2_ERR_INVALID_ARGUMENT for
 * shapes outside pn2_sa_supported(). */
int pn2_sa_supported(int c, int c1, int c2, int c3, int nsample);
int pn2_sa_forward(int b, int n, int npoint, int nsample, int c, int c1, int c2, int c3,
                   float inv_radius, const float *xyz, const float *new_xyz, const void *table,
                   const int *idx, const void *weight_image, float *out, void *out_table,
                   pn2_stream_t stream);

/* ---- fused feature-propagation layer ----------------------------------
 * Replaces PointnetFPModule.forward (pointnet2_modules.py:399-421) for eval mode:
 * three_nn -> 1/(d+1e-8) weights -> three_interpolate -> cat(skip) -> SharedMLP (2 layers).
 */
size_t pn2_fp_weight_image_bytes(int c_in, int c1, int c2);
int pn2_fp_pack_weights(int c_in, int c1, int c2, const float *w1, const float *b1,
                        const float *w2, const float *b2, void *image, pn2_stream_t stream);
int pn2_fp_supported(int c_known, int c_skip, int c1, int c2);
/* unknown (b,n,3), known (b,m,3), known_feats (b,c_known,m) f32, skip_feats (b,c_skip,n) f32
 * -> out (b,c2,n) f32.  dist2/idx (b,n,3) from pn2_three_nn. */
int pn2_fp_forward(int b, int n, int m, int c_known, int c_skip, int c1, int c2,
                   const float *dist2, const int *idx, const float *known_feats,
                   const float *skip_feats, const void *weight_image, float *out,
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

#ifdef __cplusplus
}
#endif
#endif /* PN2_B200_H_ */
