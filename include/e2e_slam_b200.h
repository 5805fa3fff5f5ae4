/*
 * e2e_slam_b200 -- C ABI of the B200 (sm_100a) differentiable-geometry hot path.
 *
 * This is the drop-in boundary.  The reference
 * (ivanalberico/End-To-End-Self-Supervised-SLAM) is 100 % Python and has no FFI of its own; the
 * functions below are what a Python binding for each reference call site needs (ctypes stub in
 * INTEGRATION.md; the shipped host-side mirror is end-to-end-self-supervised-slam_b200/e2e_slam_b200).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - fp32 everywhere (the reference never uses reduced precision), indices are int64 / int32 as stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *     and never synchronise, except e2e_prepare_divisor / e2e_fusion_count which say so;
 *   - image tensors are addressed through element strides {batch, channel, row, col}, so the
 *     reference's NCHW *views* of channels-last memory (train_depth.py:451-453) are consumed in place;
 *   - the caller owns all memory; `workspace` buffers are caller-provided (size from the *_workspace_bytes
 *     functions) and need no initialisation;
 *   - return value 0 = success, otherwise a cudaError_t (or E2E_ERR_*) value; e2e_last_error() gives text.
 *   - padding_mode: 0 = zeros, 1 = border (MODEL.padding_mode, configs/config.yaml:35).
 */
#ifndef E2E_SLAM_B200_H
#define E2E_SLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2E_ERR_BAD_ARG 100001
#define E2E_ERR_UNSUPPORTED 100002

int e2e_abi_version(void);
const char *e2e_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
unsigned long long e2e_launch_count(void);

/* The reference divides by (W-1), (H-1), 9 and 3 with IEEE division.  The kernels replace x/d by a
 * 3-instruction sequence that is correctly rounded for every x iff a property of d holds; this call
 * checks that property exhaustively on the device (all 2^23 mantissas, ~20 us), caches the verdict and
 * makes the kernels fall back to IEEE division for a divisor that fails.  Synchronises `stream` the
 * first time a divisor is seen.  Returns 1 (fast path exact), 0 (IEEE fallback) or <0 on error. */
int e2e_prepare_divisor(float d, void *stream);

/* 8-bit frames -> unit-range fp32 frames, out[i] = RN(in[i] / 255): the reference's host-side `colors /= 255.0`
 * (train_depth.py:255, online_adaption.py:215) moved to the device (bit-identical: one correctly rounded division), so a
 * frame crosses PCIe as 3 bytes per pixel instead of 12.  `in` 4-byte aligned, `out` 16-byte aligned. */
int e2e_u8_to_unit(const unsigned char *in, long long n, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused inverse warp + photometric loss.  Replaces, in one forward and one backward kernel:
 *   BackprojectDepth.forward   depth_estimation/view_synthesis.py:34-40
 *   Project3D.forward          depth_estimation/view_synthesis.py:54-78 (geometric=False)
 *   F.grid_sample              train_depth.py:587-590 / online_adaption.py:450-453 (align_corners=False)
 *   mask multiply              train_depth.py:713-718
 *   SSIM.forward               loss/losses.py:23-37
 *   photometric_loss           loss/losses.py:97-117
 *   .mean() over pixels        train_depth.py:657
 * depth [B,1,H,W] contiguous; inv_K, K, T [B,4,4] row-major; src/tgt 3-channel images via strides.
 * Optional outputs (NULL to skip): syn [B,3,H,W], valid [B,1,H,W], pix [B,H,W,2], loss_map [B,1,H,W],
 * loss_mean [1] (mean of loss_map over B*H*W; needs workspace).
 * --------------------------------------------------------------------------------------------- */
size_t e2e_warp_photo_workspace_bytes(int B, int H, int W);

int e2e_warp_photo_fwd(const float *depth, const float *inv_K, const float *K, const float *T,
                       const float *src, const int64_t src_strides[4],
                       const float *tgt, const int64_t tgt_strides[4],
                       int B, int H, int W, int padding_mode, int use_mask, float eps,
                       float *syn, float *valid, float *pix, float *loss_map, float *loss_mean,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Upstream gradient: grad_loss_map [B,1,H,W] if non-NULL, else the scalar (*grad_scalar) * scalar_scale
 * for every pixel (grad_scalar is a device pointer; NULL means 1.0).  Outputs: grad_depth [B,1,H,W]
 * (written), grad_src (ACCUMULATED with atomics into a caller-zeroed buffer addressed by
 * grad_src_strides; NULL to skip), grad_P [B,3,4] = dL/d((K@T)[:3]) (written, deterministic two-pass
 * reduction; NULL to skip).  The host maps grad_P to grad_T = K[:3]^T grad_P, grad_K[:3] = grad_P T^T. */
int e2e_warp_photo_bwd(const float *depth, const float *inv_K, const float *K, const float *T,
                       const float *src, const int64_t src_strides[4],
                       const float *tgt, const int64_t tgt_strides[4],
                       int B, int H, int W, int padding_mode, int use_mask, float eps,
                       const float *grad_loss_map, const float *grad_scalar, float scalar_scale,
                       float *grad_depth, float *grad_src, const int64_t grad_src_strides[4], float *grad_P,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Conditional form of e2e_warp_photo_bwd for callers that already hold gradients computed for a UNIFORM upstream gradient
 * (e2e_warp_photo_vg_map below): when the device-resident float *skip_if_nonzero is non-zero the kernels return at once.
 * grad_loss_map is required. */
int e2e_warp_photo_bwd_cond(const float *depth, const float *inv_K, const float *K, const float *T,
                            const float *src, const int64_t src_strides[4],
                            const float *tgt, const int64_t tgt_strides[4],
                            int B, int H, int W, int padding_mode, int use_mask, float eps,
                            const float *grad_loss_map, const float *grad_scalar, float scalar_scale, const float *skip_if_nonzero,
                            float *grad_depth, float *grad_src, const int64_t grad_src_strides[4], float *grad_P,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Single-pass value + gradient of the SCALAR loss  mean_{B,H,W} photometric_loss(...)  (the reference's use:
 * `losses += photometric.mean()` train_depth.py:657 followed by `loss.backward()` :307): one sweep over the
 * inputs produces loss_mean [1] and d loss_mean / d {depth, source image, P = (K@T)[:3]} for an upstream
 * gradient of 1.  Same argument meaning as e2e_warp_photo_fwd / _bwd: grad_depth [B,1,H,W] written,
 * grad_src ACCUMULATED into a caller-zeroed buffer (NULL to skip), grad_P [B,3,4] written (NULL to skip).
 * Workspace: e2e_warp_photo_vg_workspace_bytes().  e2e_scale_by_scalar multiplies up to three gradient
 * buffers in place by a device-resident upstream scalar and exits without touching memory when it is 1. */
size_t e2e_warp_photo_vg_workspace_bytes(int B, int H, int W);

int e2e_warp_photo_vg(const float *depth, const float *inv_K, const float *K, const float *T,
                      const float *src, const int64_t src_strides[4],
                      const float *tgt, const int64_t tgt_strides[4],
                      int B, int H, int W, int padding_mode, int use_mask, float eps,
                      float *loss_mean, float *grad_depth, float *grad_src, const int64_t grad_src_strides[4],
                      float *grad_P, void *workspace, size_t workspace_bytes, void *stream);

/* S source frames per target in ONE launch (`int S`; train_depth.py:545-613 loops over DATA.frames[1:]): the grid's z
 * extent is (pair, source).  depth [B,1,H,W], inv_K / K [B,4,4] and tgt (B pairs) are indexed by the pair; T [B*S,4,4], src (a
 * [B*S,3,H,W] view), grad_src [B*S,3,H,W] and grad_P [B*S,3,4] by (pair, source).  loss_mean = mean over the B*S*H*W per-frame loss
 * values (= `.mean(1, keepdim=True).mean()`), grad_depth [B,1,H,W] = the sum over the pair's source frames (a second, tiny launch). */
size_t e2e_warp_photo_vg_multi_workspace_bytes(int B, int S, int H, int W);
int e2e_warp_photo_vg_multi(const float *depth, const float *inv_K, const float *K, const float *T,
                            const float *src, const int64_t src_strides[4], const float *tgt, const int64_t tgt_strides[4],
                            int B, int S, int H, int W, int padding_mode, int use_mask, float eps,
                            float *loss_mean, float *grad_depth, float *grad_src, const int64_t grad_src_strides[4],
                            float *grad_P, void *workspace, size_t workspace_bytes, void *stream);

/* The same sweep fed with the depth network's DISPARITY (online_adaption.py:282, 295-298): depth =
 * (1 / disp) * ratio is formed at the kernel's depth load (ratio = device scalar of the median scaling, NULL = none; reciprocal and
 * scaling are two roundings as in the reference, so the loss is bit-identical to e2e_disp_to_depth_fwd + e2e_warp_photo_vg) and
 * grad_disp = d loss / d disp comes back directly. */
int e2e_warp_photo_vg_disp(const float *disp, const float *ratio, const float *inv_K, const float *K, const float *T,
                           const float *src, const int64_t src_strides[4], const float *tgt, const int64_t tgt_strides[4],
                           int B, int H, int W, int padding_mode, int use_mask, float eps,
                           float *loss_mean, float *grad_disp, float *grad_src, const int64_t grad_src_strides[4],
                           float *grad_P, void *workspace, size_t workspace_bytes, void *stream);

/* The same sweep with the forward tensors the reference's scripts keep (train_depth.py:581-590, 726) as outputs: loss_map
 * [B,1,H,W] (bit-exact, like e2e_warp_photo_fwd), syn [B,3,H,W], valid [B,1,H,W], pix [B,H,W,2]; each nullable, at least one
 * required; loss_mean nullable.  The gradients are those of mean(loss_map), i.e. of an upstream gradient 1/(B*H*W) at every
 * pixel -- what `photometric.mean(1, keepdim=True).mean()` (train_depth.py:629, 657) sends back.  A caller that later
 * receives an arbitrary upstream map g checks it with e2e_upstream_uniform (scale2[0] = factor for the stored gradients, 0
 * if g is not uniform; scale2[1] = 1 if uniform), applies e2e_scale_or_zero and runs e2e_warp_photo_bwd_cond with
 * skip_if_nonzero = scale2 + 1: no host synchronisation, and the backward kernel only does work when g is not uniform. */
int e2e_warp_photo_vg_map(const float *depth, const float *inv_K, const float *K, const float *T,
                          const float *src, const int64_t src_strides[4],
                          const float *tgt, const int64_t tgt_strides[4],
                          int B, int H, int W, int padding_mode, int use_mask, float eps,
                          float *loss_map, float *syn, float *valid, float *pix, float *loss_mean,
                          float *grad_depth, float *grad_src, const int64_t grad_src_strides[4],
                          float *grad_P, void *workspace, size_t workspace_bytes, void *stream);
int e2e_upstream_uniform(const float *g, long long n, double n_total, float *scale2, void *stream);
int e2e_scale_or_zero(float *a, long long na, float *b, long long nb, float *c, long long nc, const float *scale, void *stream);

int e2e_scale_by_scalar(float *a, long long na, float *b, long long nb, float *c, long long nc,
                        const float *scalar, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Stand-alone SSIM / photometric loss on given images (tier (i) drop-in for loss/losses.py:6-37 and
 * :97-117, also used for the auto-masking variant train_depth.py:729-750).  x, y are [B,C,H,W] via
 * strides.  ssim_map [B,C,H,W] and/or loss_map [B,1,H,W] (loss_map requires C == 3) may be NULL.
 * Backward: upstream grad_ssim [B,C,H,W] and/or grad_loss_map [B,1,H,W]; outputs grad_x, grad_y
 * [B,C,H,W] contiguous (either may be NULL).
 * --------------------------------------------------------------------------------------------- */
int e2e_ssim_fwd(const float *x, const int64_t x_strides[4], const float *y, const int64_t y_strides[4],
                 int B, int C, int H, int W, float *ssim_map, float *loss_map, void *stream);

int e2e_ssim_bwd(const float *x, const int64_t x_strides[4], const float *y, const int64_t y_strides[4],
                 int B, int C, int H, int W, const float *grad_ssim, const float *grad_loss_map,
                 float *grad_x, float *grad_y, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Granular view-synthesis ops (tier (i): one kernel per reference call, same tensors in and out).
 *   e2e_backproject_*  BackprojectDepth.forward, view_synthesis.py:34-40 -> cam_points [B,4,H*W]
 *   e2e_project3d_*    Project3D.forward, view_synthesis.py:54-78 -> pix [B,H,W,2], valid [B,1,H,W],
 *                      warped_depth [B,1,H,W] (geometric=True, :73-76; NULL to skip)
 *   e2e_grid_sample_*  F.grid_sample bilinear, padding zeros|border, align_corners 0|1
 *                      (train_depth.py:568-590); input [B,C,H,W] via strides, grid [B,Ho,Wo,2] contiguous
 * --------------------------------------------------------------------------------------------- */
int e2e_backproject_fwd(const float *depth, const float *inv_K, int B, int H, int W, float *cam_points, void *stream);
int e2e_backproject_bwd(const float *grad_cam, const float *inv_K, int B, int H, int W, float *grad_depth, void *stream);

int e2e_project3d_fwd(const float *points, const float *K, const float *T, int B, int H, int W, float eps,
                      float *pix, float *valid, float *warped_depth, void *stream);
/* grad_points [B,4,H*W] written; grad_P [B,3,4] written (NULL to skip; needs workspace). */
int e2e_project3d_bwd(const float *points, const float *K, const float *T, int B, int H, int W, float eps,
                      const float *grad_pix, const float *grad_warped_depth,
                      float *grad_points, float *grad_P, void *workspace, size_t workspace_bytes, void *stream);

int e2e_grid_sample_fwd(const float *input, const int64_t in_strides[4], const float *grid,
                        int B, int C, int H, int W, int Ho, int Wo, int padding_mode, int align_corners,
                        float *output, void *stream);
/* grad_input is ACCUMULATED (caller zeroes it) through grad_in_strides; grad_grid [B,Ho,Wo,2] written. */
int e2e_grid_sample_bwd(const float *grad_output, const float *input, const int64_t in_strides[4],
                        const float *grid, int B, int C, int H, int W, int Ho, int Wo,
                        int padding_mode, int align_corners,
                        float *grad_input, const int64_t grad_in_strides[4], float *grad_grid, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Small fused losses.
 *   e2e_smooth_*     compute_smoothness_loss (train_depth.py:763-773) + disparity_smoothness_loss
 *                    (loss/losses.py:119-132): per-image mean normalisation, edge-aware |dx|,|dy|, two means.
 *                    disp [B,1,H,W] contiguous, img 3-channel via strides.  loss [1].
 *   e2e_sparse_l1_*  depth_gt_loss (loss/losses.py:151-160): mean over all n elements of |pred*mask - gt|.
 *   e2e_depth_reg_*  depth_reguralizer (loss/losses.py:134-148): kind 1 = L1, 2 = L2 (MSE).
 *   e2e_geometric_fwd geometric_consistency_loss (loss/losses.py:84-95), forward + backward; the >10000
 *                    mask-count test is evaluated on the device (no host sync).
 * Backward entry points take the upstream scalar gradient as a device pointer (NULL = 1.0).
 * --------------------------------------------------------------------------------------------- */
size_t e2e_reduce_workspace_bytes(long long n_elements);

int e2e_smooth_fwd(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                   float *loss, void *workspace, size_t workspace_bytes, void *stream);
int e2e_smooth_bwd(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                   const float *grad_loss, float *grad_disp, void *workspace, size_t workspace_bytes, void *stream);

/* Smoothness value AND gradient in one sweep (every pixel loaded once, every edge weight computed once): loss [1],
 * gn [B,1,H,W] = d loss / d (normalised disparity) for an upstream gradient of 1, stats [B,2] = {1/(mean+1e-7),
 * sum(gn*disp)/((mean+1e-7)^2 HW)}.  e2e_smooth_apply then gives grad_disp = up * (gn * stats[b][0] - stats[b][1]) for a
 * device-resident upstream scalar (NULL = 1).  6.8x faster than e2e_smooth_fwd + e2e_smooth_bwd at 256 x 480 x 640. */
size_t e2e_smooth_vg_workspace_bytes(int B, int H, int W);
int e2e_smooth_vg(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                  float *loss, float *gn, float *stats, void *workspace, size_t workspace_bytes, void *stream);
int e2e_smooth_apply(const float *gn, const float *stats, const float *grad_loss, int B, int H, int W, float *grad_disp, void *stream);
/* disparity_smoothness_loss itself (loss/losses.py:119-132) for callers that normalise the disparity on their own, as the
 * unmodified compute_smoothness_loss does (train_depth.py:763-773): the same sweep on the disparity AS GIVEN; stats = {1, 0}, so
 * e2e_smooth_apply returns grad_disp = up * gn. */
int e2e_smooth_vg_raw(const float *disp, const float *img, const int64_t img_strides[4], int B, int H, int W,
                      float *loss, float *gn, float *stats, void *workspace, size_t workspace_bytes, void *stream);

int e2e_sparse_l1_fwd(const float *pred, const float *mask, const float *gt, long long n, float *loss,
                      void *workspace, size_t workspace_bytes, void *stream);
int e2e_sparse_l1_bwd(const float *pred, const float *mask, const float *gt, long long n,
                      const float *grad_loss, float *grad_pred, void *stream);

int e2e_depth_reg_fwd(const float *initial, const float *refined, long long n, int kind, float *loss,
                      void *workspace, size_t workspace_bytes, void *stream);
int e2e_depth_reg_bwd(const float *initial, const float *refined, long long n, int kind,
                      const float *grad_loss, float *grad_refined, void *stream);

int e2e_geometric_fwd(const float *warped_depth, const float *interp_depth, const float *valid, long long n,
                      float *loss, void *workspace, size_t workspace_bytes, void *stream);
/* gradients to warped / interpolated depth (either may be NULL); mask_sum = device float holding sum(valid) (the > 10000 test
 * of loss/losses.py:90 is evaluated on the device), grad_loss = device-resident upstream scalar (NULL = 1) */
int e2e_geometric_bwd(const float *warped_depth, const float *interp_depth, const float *valid, long long n, const float *mask_sum,
                      const float *grad_loss, float *grad_warped, float *grad_interp, void *stream);

/* Min-reprojection / auto-masking objective (train_depth.py:642-658): loss = mean over pixels of the per-pixel MINIMUM over
 * n_candidates (<= 8) loss maps of n = B*H*W values each (`candidates` / `grad_candidates` are HOST arrays of device pointers; the
 * reference's torch.cat is never materialised).  torch.min semantics: first minimal candidate wins, NaN propagates.  `index`
 * (uint8 [n]) receives the winner; the backward pass writes grad_loss / n into the winner's map and 0 into the others
 * (NULL entries of grad_candidates are skipped). */
int e2e_min_composite_fwd(const float *const *candidates, int n_candidates, long long n, unsigned char *index, float *loss,
                          void *workspace, size_t workspace_bytes, void *stream);
int e2e_min_composite_bwd(const unsigned char *index, int n_candidates, long long n, const float *grad_loss,
                          float *const *grad_candidates, void *stream);

/* ---------------------------------------------------------------------------------------------
 * PointFusion (gradslam semantics; gradslam itself is not vendored by the
 * reference -- call sites slam/custom_slam.py:33, online_adaption.py:354-363, 466-469, train_depth.py:266).
 * Map layout: structure of arrays, points/normals/colors [cap,3] fp32, ccount [cap] fp32.
 *
 *   e2e_rgbd_maps         RGBDImages.vertex_map / normal_map / global_* / valid mask + PointFusion alpha for
 *                         one frame: depth [H,W], rgb [H,W,3], intrinsics K [4,4], pose [4,4] (camera->world).
 *                         Outputs vertex_g, normal_g [H,W,3], alpha [H,W], valid [H,W] (uint8).
 *   e2e_fusion_associate  find_active_map_points + find_similar_map_points + find_best_unique_correspondences:
 *                         writes index_map [H,W] int64 (map point matched to each live pixel, -1 = none).
 *                         `keys` is an [H,W] uint64 scratch image, `candidates` an int32 [n_upper] scratch array
 *                         (pixel each map point competes for, -1 = none).  The map size is read from the DEVICE
 *                         int64 `n_map` (no host sync per step); `n_upper` >= *n_map only sizes the grid.
 *   e2e_fusion_merge_append fuse_with_map: confidence-weighted merge of matched map points in place, then
 *                         stream-compacted append (row-major pixel order) of valid unmatched live pixels at
 *                         map[n_map ...].  `n_out` (device int64[1]) receives the new point count;
 *                         `append_slot` [H,W] int64 receives, per pixel, the slot it was appended to (-1 = not
 *                         appended) so the backward can route gradients.  Capacity must be >= n_map + H*W.
 * --------------------------------------------------------------------------------------------- */
int e2e_rgbd_maps(const float *depth, const float *rgb, const float *K, const float *pose, int H, int W, float sigma,
                  float *vertex_g, float *normal_g, float *alpha, unsigned char *valid, void *stream);

int e2e_rgbd_maps_bwd(const float *depth, const float *K, const float *pose, int H, int W, float sigma,
                      const float *grad_vertex_g, const float *grad_normal_g, const float *grad_alpha,
                      float *grad_depth, void *stream);

int e2e_fusion_associate(const float *map_points, const float *map_normals, const float *map_ccount,
                         const long long *n_map, long long n_upper,
                         const float *K, const float *pose, const float *vertex_g, const float *normal_g,
                         int H, int W, float dist_th, float dot_th,
                         unsigned long long *keys, int *candidates, long long *index_map, void *stream);

/* find_active_map_points (gradslam.slam.fusionutils, imported by the reference at online_adaption.py:35):
 * rows (batch_index, n, h, w) int64 of the map points in front of the live camera that project into the frame
 * (-1e-3 < u < W - 0.999, -1e-3 < v < H - 0.999; pixel = round-half-even, clamped), ordered by n.  `rows` holds up to n rows;
 * `n_active` (device int64[1]) receives the row count.  Three launches, no host synchronisation. */
size_t e2e_fusion_active_points_workspace_bytes(long long n);
int e2e_fusion_active_points(const float *map_points, long long n, const float *K, const float *pose, int H, int W,
                             long long batch_index, long long *rows, long long *n_active,
                             void *workspace, size_t workspace_bytes, void *stream);

size_t e2e_fusion_workspace_bytes(int H, int W);

int e2e_fusion_merge_append(float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                            const long long *n_map, long long capacity,
                            const float *vertex_g, const float *normal_g, const float *rgb, const float *alpha,
                            const unsigned char *valid, const long long *index_map, int H, int W,
                            long long *append_slot, long long *n_out,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Backward of merge + append: gradients w.r.t. the NEW map (points, colors, ccount; any may be NULL) are
 * routed to the live frame (grad_vertex_g, grad_rgb [H,W,3], grad_alpha [H,W], written) and, for matched
 * points, to the map that entered the step (grad_old_*; the caller pre-fills them with the pass-through
 * gradient of the unmatched points; NULL to skip).  Normals are not differentiated. */
/* Whole-sequence fusion with known poses and no autograd (slam/custom_slam.py:26-34 with odom="gt"): the frame loop of
 * rgbd_maps -> associate -> merge_append runs inside the library, one call per sequence.  depth [L,H,W], rgb [L,H,W,3],
 * K [4,4], poses [L,4,4] (camera -> world); the map arrays need capacity >= n_upper + L*H*W; n_map is a device int64[2]
 * (slot 0 = point count in / out, slot 1 = scratch). */
size_t e2e_fusion_sequence_workspace_bytes(int H, int W, long long capacity);
int e2e_fusion_sequence(const float *depth, const float *rgb, const float *K, const float *poses, int L, int H, int W,
                        float sigma, float dist_th, float dot_th,
                        float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                        long long *n_map, long long n_upper, long long capacity,
                        void *workspace, size_t workspace_bytes, void *stream);

/* B independent sequences of equal shape in ONE cooperative launch (gradslam's batch dimension: train_depth.py:263-267): grid =
 * (CTAs per sequence, B), every sequence with its own barrier and working buffers, so the sequences fill each other's dependent-
 * round-trip and barrier bubbles.  Layouts: depth [B,L,H,W], rgb [B,L,H,W,3], K [B,4,4], poses [B,L,4,4]; map arrays [B,capacity,3]
 * / [B,capacity], maps start empty; n_map int64 [B,2] (zero-filled by the caller; n_map[b][0] receives the point count).  If B
 * exceeds what one launch can co-schedule the batch runs in waves. */
size_t e2e_fusion_sequence_batch_workspace_bytes(int B, int H, int W, long long capacity);
int e2e_fusion_sequence_batch(const float *depth, const float *rgb, const float *K, const float *poses, int B, int L, int H, int W,
                              float sigma, float dist_th, float dot_th,
                              float *map_points, float *map_normals, float *map_colors, float *map_ccount,
                              long long *n_map, long long capacity, void *workspace, size_t workspace_bytes, void *stream);

int e2e_fusion_merge_append_bwd(const float *grad_points, const float *grad_colors, const float *grad_ccount,
                                const float *old_points, const float *old_colors, const float *old_ccount,
                                const float *vertex_g, const float *rgb, const float *alpha,
                                const long long *index_map, const long long *append_slot, int H, int W,
                                float *grad_vertex_g, float *grad_rgb, float *grad_alpha,
                                float *grad_old_points, float *grad_old_colors, float *grad_old_ccount, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K = 1 nearest neighbour (chamferdist.chamfer.knn_points as used by knn_points_loss,
 * loss/losses.py:39-63, and compute_3d_loss, online_adaption.py:638-645).
 * query [P1,3] (optionally transformed on load by a row-major 4x4 `transform`, which is
 * gradslam.geometry.geometryutils.transform_pointcloud fused in), ref [P2,3].
 * Outputs dist2 [P1] (squared L2) and idx [P1] int64 (lowest index among exact ties).
 * Backward: grad_query[i] = 2*(q_i - r_idx[i])*g_i written (in the un-transformed frame when a
 * transform is given); grad_ref accumulated with atomics (NULL to skip).
 * --------------------------------------------------------------------------------------------- */
int e2e_knn1_fwd(const float *query, const float *transform, const float *ref, long long P1, long long P2,
                 float *dist2, long long *idx, void *stream);
int e2e_knn1_bwd(const float *query, const float *transform, const float *ref, long long P1, long long P2,
                 const long long *idx, const float *grad_dist2, float *grad_query, float *grad_ref, void *stream);

/* gradslam.geometry.geometryutils.transform_pointcloud (online_adaption.py:642): out = R p + t for (n,3) points and a 4x4 rigid
 * transform (left-to-right accumulation, identical to the query load fused into e2e_knn1_*); backward: grad_points = R^T grad_out. */
int e2e_transform_points_fwd(const float *points, const float *transform, long long n, float *out, void *stream);
int e2e_transform_points_bwd(const float *transform, const float *grad_out, long long n, float *grad_points, void *stream);

/* color_points_loss (loss/losses.py:65-82): loss = mean | noisy_colors[i] - gt_colors[idx[i]] | over P1 x 3 values (idx from
 * e2e_knn1_*).  Backward: grad_noisy [P1,3] (written), grad_gt [P2,3] (ACCUMULATED with atomics into a zero-filled buffer);
 * grad_loss = device-resident upstream scalar (NULL = 1). */
size_t e2e_color_points_workspace_bytes(long long P1);
int e2e_color_points_fwd(const float *gt_colors, const float *noisy_colors, const long long *idx, long long P1, float *loss,
                         void *workspace, size_t workspace_bytes, void *stream);
int e2e_color_points_bwd(const float *gt_colors, const float *noisy_colors, const long long *idx, long long P1, const float *grad_loss,
                         float *grad_noisy, float *grad_gt, void *stream);

/* The same answer (bit for bit: distances in the same operation order, lowest index among exact ties) through a uniform grid
 * over the reference cloud, built on the device by the call itself; cost grows with P1 + P2 instead of P1 * P2.  The host
 * side uses it when P1 * P2 is large (the online loop's point supervision: 307 200 queries against a map of millions). */
size_t e2e_knn1_grid_workspace_bytes(long long P2);
/* build once, query many times against the same reference cloud (the workspace IS the grid) */
int e2e_knn1_grid_build(const float *ref, long long P2, void *workspace, size_t workspace_bytes, void *stream);
int e2e_knn1_grid_query(const float *query, const float *transform, long long P1, long long P2,
                        float *dist2, long long *idx, const void *workspace, void *stream);
int e2e_knn1_grid_fwd(const float *query, const float *transform, const float *ref, long long P1, long long P2,
                      float *dist2, long long *idx, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Point-to-plane ICP / GradICP odometry (gradslam PointFusion with odom = "icp" / "gradicp", the reference's shipped
 * default: configs/config.yaml:30-34; constructed at train_depth.py:111-116, stepped at online_adaption.py:362-363,
 * train_depth.py:378-381).  Replaces gradslam.odometry.icputils.point_to_plane_ICP / point_to_plane_gradICP (kNN through
 * chamferdist, Jacobian / normal equations / solve as ~40 torch launches and two host syncs per iteration).
 * src [N,3], tgt / tgt_normals [M,3], T_init / T_out device 4x4 row-major.  The iteration loop runs without host
 * synchronisation (state on the device).  dist_thresh < 0: keep every pair; otherwise pairs with SQUARED distance below
 * it.  grad_icp = 0: Gauss-Newton with fixed damping `damp`; 1: logistic-gated Levenberg-Marquardt (lambda_max, B, B2,
 * nu; conventions in oracle/icp_oracle.py).  The target cloud is gridded once per call (e2e_knn1_grid_build) and queried
 * every iteration.  idx_out (nullable) [N] int64: correspondences of the last linearisation;
 * errs (nullable) [numiters] fp32: |b|^2 before every step.
 * --------------------------------------------------------------------------------------------- */
size_t e2e_icp_workspace_bytes(long long N, long long M);
int e2e_icp_point_to_plane(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                           const float *T_init, int numiters, float damp, float dist_thresh,
                           int grad_icp, float lambda_max, float B, float B2, float nu,
                           float *T_out, long long *idx_out, float *errs, void *workspace, size_t workspace_bytes, void *stream);

/* The same alignment as a DIFFERENTIABLE operation (GradICP's purpose: the pose is differentiable w.r.t. the live depth;
 * gradslam gets it from autograd through its torch ops, online_adaption.py:362-363 with odom = "gradicp").
 * e2e_icp_point_to_plane_saved runs the identical loop and leaves in `history` (e2e_icp_history_bytes(N, numiters)) what the
 * reverse sweep needs: per iteration the step, damping, gate, normal equations and transforms, the source cloud before the
 * step, both sets of correspondences (constants of the derivative, as in gradslam).  e2e_icp_backward then propagates
 * grad_T_out [4x4, the last row is ignored] to grad_src [N,3] (written), grad_tgt / grad_normals [M,3] (ACCUMULATED with atomics:
 * pass zeroed buffers, or NULL to skip) and grad_T_init [4x4] (nullable): four launches per iteration, no host synchronisation;
 * the derivative of the se3 exponential is taken with float64 dual numbers on the device. */
size_t e2e_icp_history_bytes(long long N, int numiters);
int e2e_icp_point_to_plane_saved(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                                 const float *T_init, int numiters, float damp, float dist_thresh,
                                 int grad_icp, float lambda_max, float B, float B2, float nu,
                                 float *T_out, long long *idx_out, void *workspace, size_t workspace_bytes,
                                 void *history, size_t history_bytes, void *stream);
size_t e2e_icp_backward_workspace_bytes(long long N);
int e2e_icp_backward(const float *src, long long N, const float *tgt, const float *tgt_normals, long long M,
                     const float *T_init, int numiters, float damp, float dist_thresh,
                     int grad_icp, float lambda_max, float B, float B2, float nu,
                     const void *history, const float *grad_T_out,
                     float *grad_src, float *grad_tgt, float *grad_normals, float *grad_T_init,
                     void *workspace, size_t workspace_bytes, void *stream);

/* k-th smallest (0-based) of n fp32 values by radix select: four rounds of an 8-bit histogram + pick, nothing is sorted and the host
 * is never synchronised.  torch.median(x) (online_adaption.py:295) is k = (n-1)/2; a NaN anywhere gives NaN, as torch does.
 * median(1 / disp) of a positive disparity map is 1 / (the n/2-th smallest disparity) exactly (x -> RN(1/x) is monotone), so the
 * median-scaling ratio needs no materialised depth. */
size_t e2e_select_workspace_bytes(void);
int e2e_select_kth(const float *x, long long n, long long k, float *out, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The elementwise passes next to the depth network.
 *   e2e_disp_to_depth_*    depth = (1 / disp) * ratio: `1 / inputs[("disp", ...)]` and the median rescaling `*= ratio`
 *                          (online_adaption.py:282, 295-298; train_depth.py:323-340) in one pass; ratio = device scalar, NULL = 1.
 *   e2e_dual_disparity_*   process_disparity (train_depth.py:224-237): left [H,W] = disparity of the frame, right [H,W] = disparity
 *                          of its mirror image, row_mask [H] = 1 - clip(20 (linspace(0,1,H) - 0.05), 0, 1).
 * --------------------------------------------------------------------------------------------- */
int e2e_disp_to_depth_fwd(const float *disp, const float *ratio, long long n, float *depth, void *stream);
int e2e_disp_to_depth_bwd(const float *disp, const float *ratio, const float *grad_depth, long long n, float *grad_disp, void *stream);
int e2e_dual_disparity_fwd(const float *left, const float *right, const float *row_mask, int H, int W, float *out, void *stream);
int e2e_dual_disparity_bwd(const float *grad_out, const float *row_mask, int H, int W, float *grad_left, float *grad_right, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: the one collective of the data-parallel path -- all-reduce(mean) of the depth network's adaptation
 * gradients (~57 MB fp32) -- as an NVLS kernel: `multicast_ptr` is the multicast address of a symmetric-memory bucket of `numel`
 * floats (multiple of 4, 16-byte aligned); rank r reduces the r-th slice in the switch (multimem.ld_reduce) and broadcasts the mean
 * (multimem.st).  The caller brackets the call with cross-rank barriers on the same stream.  `ctas` <= 0: default (8).
 * --------------------------------------------------------------------------------------------- */
int e2e_multimem_allreduce_avg(void *multicast_ptr, long long numel, int rank, int world, int ctas, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* E2E_SLAM_B200_H */
