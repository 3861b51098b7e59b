/* sqlx.h -- C ABI of libsqlx.so: the B200 (sm_100a) hot path of SQLdepth self-supervised training.
 *
 * The reference (hisfog/SfMNeXt-Impl, pure Python/PyTorch) has no FFI; the drop-in boundary is the
 * set of nn.Module / Trainer signatures.  This header is what a maintainer binds from
 * those Python call sites (ctypes).  Each entry point cites the reference lines
 * (relative to /root/reference) whose device work it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (NCHW for images) unless marked "host";
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the device or allocates
 *     persistent device memory: outputs and workspaces are caller-owned (sizes via *_workspace_bytes);
 *   - return value 0 = success, negative SQLX_E* otherwise; sqlx_last_error() gives a message
 *     (thread-local).  The library is re-entrant (forward thread + autograd thread).
 */
#ifndef SQLX_H_
#define SQLX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQLX_OK 0
#define SQLX_EINVAL (-1)      /* bad argument / unsupported shape */
#define SQLX_ECUDA (-2)       /* CUDA runtime error (launch or sticky) */
#define SQLX_EWORKSPACE (-3)  /* workspace too small */

/* flags of sqlx_photo_desc.flags  (trainer.py option it mirrors) */
#define SQLX_AUTOMASK 1u      /* NOT --disable_automasking   trainer.py:478-493,514-519 */
#define SQLX_AVG_REPROJ 2u    /* --avg_reprojection          trainer.py:489,509 */
#define SQLX_NO_SSIM 4u       /* --no_ssim                   trainer.py:446-447 */

#define SQLX_MAX_SOURCES 4

typedef struct sqlx_photo_desc {
  int32_t B, H, W;        /* batch (per GPU), full-resolution height/width (opt.height/opt.width) */
  int32_t h, w;           /* resolution of the network output at this scale, outputs[("disp",s)] */
  int32_t S;              /* number of source frames, len(frame_ids)-1, 1..SQLX_MAX_SOURCES */
  int32_t ssim_radius;    /* 3 = layers.py:19 (7x7, the live SSIM); 1 = calc_layers.py:223 (3x3) */
  uint32_t flags;         /* SQLX_AUTOMASK | SQLX_AVG_REPROJ | SQLX_NO_SSIM */
  float w_ssim, w_l1;     /* 0.85 / 0.15, trainer.py:451 */
  float noise_scale;      /* 1e-5, trainer.py:517 */
  float eps;              /* Project3D eps 1e-7, layers.py:239 */
} sqlx_photo_desc;

const char* sqlx_last_error(void);
int sqlx_version(void);
/* 1 if the visible device is compute capability 10.x (the only target this library is built for) */
int sqlx_device_ok(int device);
/* kernels launched by this library in this process so far (bench.py reports the per-step difference) */
unsigned long long sqlx_launch_count(void);

/* Per-kernel device timing for bench.py's roofline figure: while enabled, every main kernel launch is
 * bracketed by CUDA events on its stream (skipped during CUDA-graph capture).  sqlx_profile_report waits for
 * the recorded events and writes "<kernel> <launches> <total_ms>\n" lines into buf; returns bytes written. */
int sqlx_profile_enable(int on);
int sqlx_profile_report(char* buf, size_t buf_bytes);

/* ---------------------------------------------------------------------------------------------
 * Photometric block
 * ------------------------------------------------------------------------------------------- */

/* Per-sample statistics of the bilinearly upsampled (align_corners=False) depth:
 *   stats[b][0] = mean_{HxW} d_up,  stats[b][1] = mean_{HxW} 1/d_up
 * replaces F.interpolate + (1/depth).mean  (trainer.py:395-396, 417-418) and disp.mean (trainer.py:535).
 * stats is overwritten (no zeroing needed).  workspace: sqlx_depth_stats_workspace_bytes. */
size_t sqlx_depth_stats_workspace_bytes(int B, int H, int W);
int sqlx_depth_stats_fwd(const float* depth_lr, int B, int h, int w, int H, int W,
                         float* stats /*[B,2]*/, void* workspace, size_t workspace_bytes, void* stream);
/* d_depth_lr[b,i,j] += g_stats[b][0]*d(mean d)/d(lr) + g_stats[b][1]*d(mean 1/d)/d(lr)   (accumulates) */
int sqlx_depth_stats_bwd(const float* depth_lr, int B, int h, int w, int H, int W,
                         const float* g_stats /*[B,2]*/, float* d_depth_lr /*[B,h,w]*/, void* stream);

/* Reprojection loss of two image batches: out[b,0,v,u] = w_ssim*mean_c SSIM(pred,target) + w_l1*mean_c|target-pred|
 * replaces Trainer.compute_reprojection_loss (trainer.py:441-453) and SSIM.forward (layers.py:31-46).
 * Used once per step for the identity (auto-mask) losses of every source (trainer.py:480-493). */
int sqlx_reprojection_loss_fwd(const float* pred, const float* target, int B, int H, int W, int ssim_radius,
                               float w_ssim, float w_l1, int no_ssim, float* out /*[B,H,W]*/, void* stream);

/* Full SSIM map, module-level drop-in for layers.SSIM.forward (layers.py:31-46): out [B,3,H,W] */
int sqlx_ssim_fwd(const float* x, const float* y, int B, int C, int H, int W, int ssim_radius,
                  float* out, void* stream);
/* gradient wrt x and (optionally, may be NULL) y given g_out [B,C,H,W] */
int sqlx_ssim_bwd(const float* x, const float* y, const float* g_out, int B, int C, int H, int W,
                  int ssim_radius, float* gx, float* gy, void* stream);

/* Fused: upsample -> backproject -> project -> border-clamped bilinear gather of S sources ->
 * SSIM+L1 vs target -> min with identity (+noise) -> sum.
 * replaces layers.py:210-215,247-258,31-46 + trainer.py:395-396,423-435,444-451,474-532.
 *   depth_lr  [B,1,h,w]       network output at this scale (this IS depth: trainer.py:399-402)
 *   target    [B,3,H,W]       inputs[("color",0,0)]
 *   sources_rgba  host array of S device pointers, each a [B,H,W,4] pixel-interleaved copy of inputs[("color",f,0)]
 *             made by sqlx_pack_rgba (16-byte aligned): one 128-bit load per bilinear tap.  The frames are packed
 *             once per step and shared by every loss scale, forward and backward.
 *   K, inv_K  [B,4,4]
 *   T         [B,S,4,4]       camera transforms (already rescaled by mean inverse depth when posecnn)
 *   identity  [B,S,H,W] or NULL (no automask): sqlx_reprojection_loss_fwd(source_f, target), WITHOUT noise
 *   noise     [B,Sn,H,W] or NULL: standard-normal tie-break noise, Sn = 1 if AVG_REPROJ else S (trainer.py:516)
 * outputs
 *   loss_sum  [1]  float: sum over b,v,u of the per-pixel minimum (caller divides by B*H*W)
 *   argmin    [B,H,W] uint8: index into cat(identity, reprojection) exactly as torch.min(combined,dim=1)
 *   ssim_coef [B,S,3,3,H,W] (sqlx_photo_coef_bytes) or NULL for a forward-only call: d SSIM/d(mean_x, E[x^2], E[xy])
 *             per source / channel / pixel.  sqlx_photo_bwd is an arg-min masked adjoint box filter of these planes
 *             (no recomputation of the warp and the box sums on a doubled halo).
 */
size_t sqlx_photo_workspace_bytes(const sqlx_photo_desc* desc);
size_t sqlx_photo_coef_bytes(const sqlx_photo_desc* desc);
/* [B,3,H,W] planar frame -> [B,H,W,4] pixel-interleaved (r,g,b,0) frame; `out` 16-byte aligned. */
int sqlx_pack_rgba(const float* image, int B, int H, int W, float* out, void* stream);
int sqlx_photo_fwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                   const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                   const float* identity, const float* noise,
                   float* loss_sum, uint8_t* argmin, float* ssim_coef, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Backward of sqlx_photo_fwd.  g_loss is a DEVICE scalar (upstream gradient of the per-scale loss);
 * `scale` is a host multiplier (1/(B*H*W)).
 *   d_depth_lr [B,h,w]  accumulated (+=) -- caller zeroes
 *   d_T        [B,S,4,4] overwritten     (gradient wrt T; rows 3 are zero)
 */
int sqlx_photo_bwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                   const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                   const uint8_t* argmin, const float* ssim_coef, const float* g_loss, float scale,
                   float* d_depth_lr, float* d_T, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Indoor variant: generate_images_pred + the photometric part of compute_losses_with_occ,
 * trainer_indoor.py:512-599 and 615-699 (--use_improved_mini_reproj_loss), for one loss scale.  On top of
 * sqlx_photo_fwd, per source f and pixel:
 *   pd     = grid_sample(depth_ref_f, pix_coords, border, align_corners=True)         trainer_indoor.py:583-587
 *   valid  = mean_c |warped_f| > 1e-3                                                 :636
 *   diff   = |d - pd| / (d + pd),  d = the upsampled depth                            :642-643
 *   weight = 1 - sqrt(1 - (diff - 1)^2)   (detached)                                  :647-648
 *   reprojection_f *= weight * valid  before the mean / minimum over candidates       :650-651, 690
 *   ref_depths  host array of S device pointers, each [B,1,H,W]: outputs[("depth_ref",f,0)]
 *   sums [2]    sums[0] = sum over pixels of the per-pixel minimum, sums[1] = sum over sources and pixels of
 *               diff * valid  (the caller forms mean(min) + reg_wt * sums[1] / (S*B*H*W), :698-699)
 * The exported SSIM coefficients already carry the per-pixel weight.  Backward: g_sums [2] device (upstream
 * gradients of sums[0], sums[1]); d_depth_lr accumulated (caller zeroes), d_T overwritten, d_ref_depths[f] [B,1,H,W]
 * accumulated with atomics (caller zeroes): the source depths are network outputs too (:374-377). */
size_t sqlx_photo_occ_workspace_bytes(const sqlx_photo_desc* desc);
int sqlx_photo_occ_fwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                       const float* const* sources_rgba, const float* const* ref_depths, const float* K,
                       const float* inv_K, const float* T, const float* identity, const float* noise,
                       float* sums, uint8_t* argmin, float* ssim_coef, void* workspace, size_t workspace_bytes,
                       void* stream);
int sqlx_photo_occ_bwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                       const float* const* sources_rgba, const float* const* ref_depths, const float* K,
                       const float* inv_K, const float* T, const uint8_t* argmin, const float* ssim_coef,
                       const float* g_sums, float scale, float* d_depth_lr, float* d_T, float* const* d_ref_depths,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Warp only (materialises what Trainer.log reads): sample [B,H,W,2] normalised grid (layers.py:255-257),
 * color [B,3,H,W] warped source (trainer.py:431-435), depth_up [B,1,H,W] (trainer.py:402). Any output may be NULL. */
int sqlx_warp_fwd(const float* depth_lr, const float* source, const float* K, const float* inv_K,
                  const float* T /*[B,4,4] with batch stride T_stride floats*/, int T_stride,
                  int B, int h, int w, int H, int W, float eps,
                  float* depth_up, float* sample, float* color, void* stream);

/* Identity (auto-mask) losses of all S sources, written straight into identity [B,S,H,W]
 * (trainer.py:480-493: compute_reprojection_loss(inputs[("color",f,0)], target) for every source). */
int sqlx_identity_losses_fwd(const float* target, const float* const* sources, int S, int B, int H, int W,
                             int ssim_radius, float w_ssim, float w_l1, int no_ssim, float* identity, void* stream);

/* ---- One whole loss scale (the body of the `for scale in self.opt.scales` loops, trainer.py:390-439 and 461-545)
 * as one forward and one backward call; the library chains its own kernels on `stream`:
 *   depth statistics -> pose matrices (posecnn translation rescale, trainer.py:412-421) -> fused photometric
 *   kernel -> smoothness (trainer.py:533-542) -> loss_s = mean(min) + smooth_weight * smooth
 * smooth_weight = disparity_smoothness / 2^s (trainer.py:542). */
typedef struct sqlx_scale_desc {
  sqlx_photo_desc photo;
  int32_t Hc, Wc;               /* resolution of inputs[("color",0,s)] used by the smoothness term */
  float smooth_weight;
  int32_t rescale_translation;  /* 1: translation *= mean inverse depth (posecnn and not use_stereo) */
} sqlx_scale_desc;

typedef struct sqlx_pose_inputs {
  const float* axisangle[SQLX_MAX_SOURCES];    /* [B,3] per source, or NULL when the source uses fixed_T */
  const float* translation[SQLX_MAX_SOURCES];  /* [B,3] */
  const float* fixed_T[SQLX_MAX_SOURCES];      /* [B,4,4] (inputs["stereo_T"]) when axisangle is NULL */
  uint32_t invert_mask;                        /* bit s set: invert=True for source s (frame_id < 0, trainer.py:336) */
} sqlx_pose_inputs;

/* `sources_rgba`: as for sqlx_photo_fwd (sqlx_pack_rgba copies).
 * `saved` (sqlx_scale_saved_bytes) carries T, depth statistics and smoothness sums from forward to backward;
 * `workspace` (sqlx_scale_workspace_bytes) is scratch.  loss: device scalar [1].  argmin [B,H,W] u8. */
size_t sqlx_scale_saved_bytes(const sqlx_scale_desc* desc);
size_t sqlx_scale_workspace_bytes(const sqlx_scale_desc* desc);
int sqlx_scale_loss_fwd(const sqlx_scale_desc* desc, const float* depth_lr, const float* target,
                        const float* const* sources_rgba, const float* color_s, const float* K, const float* inv_K,
                        const sqlx_pose_inputs* poses, const float* identity, const float* noise, float* loss,
                        uint8_t* argmin, void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                        void* stream);
/* g_loss: device scalar, upstream gradient of loss.  d_depth_lr [B,h,w] overwritten; d_axisangle[s], d_translation[s]
 * [B,3] overwritten for pose-net sources (entries may be NULL). */
int sqlx_scale_loss_bwd(const sqlx_scale_desc* desc, const float* depth_lr, const float* target,
                        const float* const* sources_rgba, const float* color_s, const float* K, const float* inv_K,
                        const sqlx_pose_inputs* poses, const uint8_t* argmin, const float* g_loss, const void* saved,
                        float* d_depth_lr, float* const* d_axisangle, float* const* d_translation, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- ALL loss scales as one forward and one backward call (the whole of generate_images_pred + compute_losses,
 * trainer.py:386-439 and 455-549).  Every kernel is batched across scales (the fused photometric forward / backward
 * kernels take all scales in one launch): 3 + 4 kernel launches per step; every reduction has a fixed order and the upsample adjoint is a gather, so the
 * gradients are bit-reproducible run to run.
 *   depth_lr[s] [B,1,h[s],w[s]]   outputs[("disp", s)] (these ARE depth, trainer.py:399-402)
 *   color[s]    [B,3,Hc[s],Wc[s]] inputs[("color",0,s)]: either the depth map's own shape (used as is) or HxW
 *                                 (the map is upsampled first, trainer.py:533-534)
 *   noise[s]    [B,Sn,H,W]        tie-break noise of scale s
 *   loss        [1 + num_scales]  device: {losses["loss"], losses["loss/0"], losses["loss/1"], ...}
 *   argmin[s]   [B,H,W] u8 */
#define SQLX_MAX_SCALES 8
typedef struct sqlx_ms_desc {
  sqlx_photo_desc photo;                 /* B, H, W, S, flags, weights; photo.h / photo.w are ignored */
  int32_t num_scales;
  int32_t h[SQLX_MAX_SCALES], w[SQLX_MAX_SCALES];
  int32_t Hc[SQLX_MAX_SCALES], Wc[SQLX_MAX_SCALES];
  float smooth_weight[SQLX_MAX_SCALES];  /* disparity_smoothness / 2^s (trainer.py:542) */
  int32_t rescale_translation;           /* 1: translation *= mean inverse depth (posecnn and not use_stereo) */
} sqlx_ms_desc;
size_t sqlx_ms_saved_bytes(const sqlx_ms_desc* desc);
size_t sqlx_ms_workspace_bytes(const sqlx_ms_desc* desc);
int sqlx_ms_loss_fwd(const sqlx_ms_desc* desc, const float* const* depth_lr, const float* target,
                     const float* const* sources_rgba, const float* const* color, const float* K, const float* inv_K,
                     const sqlx_pose_inputs* poses, const float* identity, const float* const* noise, float* loss,
                     uint8_t* const* argmin, void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                     void* stream);
/* g_loss: device scalar, upstream gradient of loss[0].  d_depth_lr[s] [B,h[s],w[s]] overwritten; d_axisangle[f],
 * d_translation[f] [B,3] overwritten with the sum over scales (entries may be NULL). */
int sqlx_ms_loss_bwd(const sqlx_ms_desc* desc, const float* const* depth_lr, const float* target,
                     const float* const* sources_rgba, const float* const* color, const float* K, const float* inv_K,
                     const sqlx_pose_inputs* poses, const uint8_t* const* argmin, const float* g_loss, const void* saved,
                     float* const* d_depth_lr, float* const* d_axisangle, float* const* d_translation, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- Bins head: the nn.Linear layers of bins_regressor and the bin-centre arithmetic
 * (networks/depth_decoder_QTR.py:48-66) as weight-streaming kernels for batch <= 16 per GPU.
 *   linear_fwd : y [B,N] = act(x [B,K] W^T [N,K] + bias), act = LeakyReLU(0.01) when leaky
 *   linear_bwd : dy is the gradient wrt y; dz [B,N] scratch; dW [N,K], db [N] and (unless NULL) dx [B,K] overwritten
 *   centers    : raw [B,D] -> centers [B,D] for norm == 'linear' (relu + 0.1, normalise, widths, cumsum, mid-points) */
int sqlx_head_linear_fwd(const float* W, const float* bias, const float* x, int B, int N, int K, int leaky, float* y,
                         void* stream);
int sqlx_head_linear_bwd(const float* W, const float* x, const float* y, const float* dy, int B, int N, int K, int leaky,
                         float* dz, float* dW, float* db, float* dx, void* stream);
int sqlx_head_centers_fwd(const float* raw, int B, int D, float min_val, float max_val, float* centers, void* stream);
int sqlx_head_centers_bwd(const float* raw, const float* centers, const float* g_centers, int B, int D, float min_val,
                          float max_val, float* d_raw, void* stream);

/* ---- Supervised fine-tuning loss: finetune/loss.py:29-42 SILogLoss.forward with the
 * align_corners=True bilinear resize (train_ft_SQLdepth.py:235) and the boolean-mask gather fused into one pass.
 *   pred [B,1,h,w]; gt [B,1,H,W]; mask [B,1,H,W] u8 (torch.bool storage) or NULL; loss [1]; saved [4] floats
 *   (mean, count, Dg, loss) carried to the backward; d_pred [B,1,h,w] overwritten. */
/* inverse_rotation_warp of the indoor trainer's rectification step (layers.py:460-479; N4):
 *   out = grid_sample(img, pix, padding_mode="zeros", align_corners=True), pix = (P w).xy / ((P w).z + 1e-7),
 *   w = depth_to_3d(ones, K)(u,v) = ((u - cx)/fx, (v - cy)/fy, 1), P [B,3,3] = K . euler2mat(rot) built by the caller.
 * Backward: d_P [B,3,3] (autograd carries it to rot); workspace zero-initialised ONCE by the caller (left zero). */
size_t sqlx_rotation_warp_workspace_bytes(int B);
int sqlx_rotation_warp_fwd(const float* img, const float* P, const float* K3, int B, int H, int W, float* out, void* stream);
int sqlx_rotation_warp_bwd(const float* img, const float* P, const float* K3, const float* g_out, int B, int H, int W,
                           float* d_P, void* workspace, size_t workspace_bytes, void* stream);

/* Per-sample median scaling of the fine-tuning loop (finetune/train_ft_SQLdepth.py:236-266), on the device:
 *   ratio[i] = median(depth_i[valid]) / median(pred_i[valid]) for i < count (1 when either median is NaN), 1 for i >= count
 *   valid = min_depth_eval < depth < max_depth_eval inside the crop rows [r0,r1) x columns [c0,c1) (garg / eigen crop)
 * pred, depth [B,H,W] fp32 (pred already resized to the ground truth, :235).  Exact radix select of the two middle order
 * statistics (numpy.median): four 8-bit digit passes + one "next key" pass, five launches of (slices x 2 arrays x count)
 * CTAs whose shared-memory histograms are merged into global ones; replaces one device->host->device round trip per
 * sample.  workspace: scratch of sqlx_median_ratio_workspace_bytes(B) bytes, cleared by the call itself. */
size_t sqlx_median_ratio_workspace_bytes(int B);
int sqlx_median_ratio(const float* pred, const float* depth, int B, int H, int W, int count, float min_depth_eval,
                      float max_depth_eval, int r0, int r1, int c0, int c1, float* ratio, void* workspace,
                      size_t workspace_bytes, void* stream);
/* The same with pred [B,h,w] at its own resolution, read through the bilinear align_corners=True resize to H x W of
 * finetune/train_ft_SQLdepth.py:235 (the resized map is never materialised). */
int sqlx_median_ratio_resized(const float* pred, int h, int w, const float* depth, int B, int H, int W, int count,
                              float min_depth_eval, float max_depth_eval, int r0, int r1, int c0, int c1, float* ratio,
                              void* workspace, size_t workspace_bytes, void* stream);

size_t sqlx_silog_workspace_bytes(void);
int sqlx_silog_fwd(const float* pred, const float* gt, const uint8_t* mask, int B, int h, int w, int H, int W,
                   float variance_focus, float* loss, float* saved, void* workspace, size_t workspace_bytes,
                   void* stream);
int sqlx_silog_bwd(const float* pred, const float* gt, const uint8_t* mask, int B, int h, int w, int H, int W,
                   float variance_focus, const float* saved, const float* g_loss, float* d_pred, void* stream);

/* ---- Evaluation: flip test-time-augmentation blend, evaluate_depth_config.py:51-59
 * (batch_post_process_disparity) fused with the un-flip of the second pass (:157).  l_disp, r_disp, out [N,h,w];
 * r_is_flipped = 1 when r_disp still is in the mirrored frame (as the network returned it). */
int sqlx_postprocess_disparity(const float* l_disp, const float* r_disp, int N, int h, int w, int r_is_flipped,
                               float* out, void* stream);

/* Module-level geometry drop-ins (the fused path above never materialises these tensors).
 * BackprojectDepth.forward (layers.py:210-215): depth [B,1,H,W], inv_K [B,4,4] -> points [B,4,H*W] (row 3 = 1). */
int sqlx_backproject_fwd(const float* depth, const float* inv_K, int B, int H, int W, float* points, void* stream);
/* d_depth [B,1,H,W] (overwritten) given g_points [B,4,H*W] */
int sqlx_backproject_bwd(const float* g_points, const float* inv_K, int B, int H, int W, float* d_depth, void* stream);
/* Project3D.forward (layers.py:247-258): points [B,4,N], K, T [B,4,4] -> normalised grid [B,H,W,2] */
int sqlx_project_fwd(const float* points, const float* K, const float* T, int B, int H, int W, float eps,
                     float* grid, void* stream);
/* d_points [B,4,N] (overwritten, may be NULL) and d_T [B,4,4] (overwritten) given g_grid [B,H,W,2];
 * workspace: 48*B bytes. */
int sqlx_project_bwd(const float* points, const float* K, const float* T, const float* g_grid, int B, int H, int W,
                     float eps, float* d_points, float* d_T, void* workspace, size_t workspace_bytes, void* stream);

/* Edge-aware smoothness (layers.py:267-280 + trainer.py:533-542) of the mean-normalised disparity.
 * disp_lr [B,1,h,w] is upsampled to the colour resolution [Hc,Wc] when the shapes differ.
 *   sums [B,3] = { sum_x |dx d| e^{-|dx I|}, sum_y |dy d| e^{-|dy I|}, sum d }   (over the upsampled map)
 *   loss = sum_b sums[b][0]/(mean_b+1e-7) / (B*Hc*(Wc-1)) + sum_b sums[b][1]/(mean_b+1e-7) / (B*(Hc-1)*Wc)
 * The tiny [B,3] -> scalar step is done by the host wrapper in torch (differentiable). */
size_t sqlx_smooth_workspace_bytes(int B, int Hc, int Wc);
int sqlx_smooth_fwd(const float* disp_lr, const float* color, int B, int h, int w, int Hc, int Wc,
                    float* sums /*[B,3]*/, void* workspace, size_t workspace_bytes, void* stream);
/* d_disp_lr += adjoint given g_sums [B,3] (device) */
int sqlx_smooth_bwd(const float* disp_lr, const float* color, int B, int h, int w, int Hc, int Wc,
                    const float* g_sums, float* d_disp_lr, void* stream);

/* Pose matrix: T[b] = transformation_from_parameters(axisangle[b], translation[b]*scale[b], invert)
 * replaces layers.py:75-150 (~60 tiny kernels per call in the reference) and the translation rescale of
 * trainer.py:417-421.  axisangle, translation [B,3]; scale [B] or NULL (= 1); T [B,4,4]. */
int sqlx_pose_fwd(const float* axisangle, const float* translation, const float* scale, int B, int invert,
                  float* T, void* stream);
/* gradients wrt the three inputs given dT [B,4,4]; d_scale NULL iff scale NULL */
int sqlx_pose_bwd(const float* axisangle, const float* translation, const float* scale, int B, int invert,
                  const float* dT, float* d_axisangle, float* d_translation, float* d_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SQL block (Self Query Layer tail of the depth decoder)
 * ------------------------------------------------------------------------------------------- */

/* Pixel-softmax summaries: summary[b,q,:] = sum_p softmax_p(x^T K)[p,q] * x[:,p]
 * replaces FullQueryLayer.forward's summary path (networks/layers.py:17-19) without materialising [B,n,Q].
 *   x [B,E,n] (n = h*w, NCHW view), queries [B,Q,E]
 *   summary [B,Q,E]; row_max [B,Q], row_sum [B,Q] (softmax statistics, saved for backward)
 *   energy: optional [B,Q,n] (the module-level drop-in returns it; NULL inside the fused decoder) */
size_t sqlx_sql_workspace_bytes(int B, int E, int Q, int D, int n);
int sqlx_sql_summary_fwd(const float* x, const float* queries, int B, int E, int Q, int n,
                         float* summary, float* row_max, float* row_sum, float* energy,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Tensor-core (tcgen05 / TMEM / TMA) variants.  sqlx_sql_tc_supported returns 1 when the shape is taken by the
 * tensor-core kernels (E = 32, Q <= 128, D <= 128, n % 4 == 0); the fp32 entry points above/below dispatch to
 * them automatically, so these are exported mainly for tests and profiling.
 * sqlx_sql_energy_tc: energy[b,q,p] = sum_e x[b,e,p] queries[b,q,e] as 3xTF32 (networks/layers.py:17,20). */
int sqlx_sql_tc_supported(int E, int Q, int D, int n);
/* on = 0 forces the exact-fp32 CUDA-core kernels for every shape (A/B tests); returns the previous setting */
int sqlx_sql_set_tensor_cores(int on);
int sqlx_sql_get_tensor_cores(void);   /* current setting (read-only) */
/* SMs the one-CTA-per-SM SQL kernels launched from now on may occupy (default all 148; clamped to [8, 148]); returns the
 * previous budget.  For callers that run a communication kernel beside the summary-path backward. */
int sqlx_sql_set_sm_budget(int sms);
int sqlx_sql_energy_tc(const float* x, const float* queries, int B, int E, int Q, int n, float* energy, void* stream);

/* Mixed-weight decomposition (tensor cores only; sqlx_sql_tc_supported must hold):
 *   logits = Wp (K x) + b = (Wp K) x + b = M x + b with M [B,D,E] = Wp . queries (sqlx_sql_mix_weights).
 * The regression and its backward then contract over E = 32 instead of Q and need no Wp tiles on chip:
 *   sqlx_sql_pred_mix_fwd   pred [B,n]                                  (depth_decoder_QTR.py:61,70)
 *   sqlx_sql_bwd_pred_mix   d_M [B,D,E], d_bp [D], d_centers [B,D], d_x [B,E,n] (regression path; all overwritten)
 *   sqlx_sql_bwd_summary    d_x (+)= summary path, d_queries [B,Q,E] = summary-path part of d_K (overwritten)
 * The caller finishes with d_Wp = sum_b d_M queries^T and d_queries += Wp^T d_M (sqlx_sql_mix_weights_bwd).
 *   sqlx_sql_mix_weights      Mx [B,D,E] = Wp [D,Q] . queries [B,Q,E]   (the 1x1 conv weight of depth_decoder_QTR.py:28
 *                             folded into the queries of networks/layers.py:17)
 *   sqlx_sql_mix_weights_bwd  d_Wp [D,Q] = sum_b d_Mx[b] queries[b]^T (overwritten);
 *                             d_queries [B,Q,E] (+)= Wp^T d_Mx[b]  (accumulate_d_queries: add to what
 *                             sqlx_sql_bwd_summary wrote); either output may be NULL (skipped), so the weight
 *                             gradient can be produced before the gradient exchange starts and the query part later
 * Fixed summation order; no workspace. */
size_t sqlx_sql_mix_workspace_bytes(int B, int Q, int D, int n);
int sqlx_sql_mix_weights(const float* Wp, const float* queries, int B, int Q, int D, int E, float* Mx, void* stream);
int sqlx_sql_mix_weights_bwd(const float* d_Mx, const float* queries, const float* Wp, int B, int Q, int D, int E,
                             int accumulate_d_queries, float* d_Wp, float* d_queries, void* stream);
/* stats [2][B][n] (written by the forward, read by the backward): the per-pixel softmax statistics of the base-2 logits
 * t = (M x + b) log2(e) -- plane 0: max_d t, plane 1: 1 / sum_d 2^(t - max).  With them the backward needs no max / sum pass. */
int sqlx_sql_pred_mix_fwd(const float* x, const float* Mx, const float* bp, const float* centers, int B, int E, int D,
                          int n, float* pred, float* stats, void* stream);
int sqlx_sql_bwd_pred_mix(const float* x, const float* Mx, const float* bp, const float* centers, const float* g_pred,
                          const float* pred, const float* stats, int B, int E, int D, int n, float* d_M, float* d_bp,
                          float* d_centers, float* d_x, void* workspace, size_t workspace_bytes, void* stream);
/* round-1 generation of the two entry points above (one warpgroup, serial phases; no statistics): kept for one round as the
 * A/B and cross-check of the warp-specialised kernels */
int sqlx_sql_pred_mix_fwd_v1(const float* x, const float* Mx, const float* bp, const float* centers, int B, int E, int D,
                             int n, float* pred, void* stream);
int sqlx_sql_summary_fwd_v1(const float* x, const float* queries, int B, int E, int Q, int n, float* summary,
                            float* row_max, float* row_sum, void* workspace, size_t workspace_bytes, void* stream);
int sqlx_sql_bwd_summary_v1(const float* x, const float* queries, const float* summary, const float* row_max,
                            const float* row_sum, const float* d_summary, int B, int E, int Q, int n, int accumulate,
                            float* d_x, float* d_queries, void* workspace, size_t workspace_bytes, void* stream);
int sqlx_sql_bwd_pred_mix_v1(const float* x, const float* Mx, const float* bp, const float* centers, const float* g_pred,
                             int B, int E, int D, int n, float* d_M, float* d_bp, float* d_centers, float* d_x,
                             void* workspace, size_t workspace_bytes, void* stream);
int sqlx_sql_bwd_summary(const float* x, const float* queries, const float* summary, const float* row_max,
                         const float* row_sum, const float* d_summary, int B, int E, int Q, int n, int accumulate,
                         float* d_x, float* d_queries, void* workspace, size_t workspace_bytes, void* stream);

/* Depth regression: pred[b,p] = sum_d softmax_d(Wp (x^T K)[p,:] + bp)[d] * centers[b,d]
 * replaces networks/layers.py:17,20 + depth_decoder_QTR.py:61,70 (1x1 conv, Softmax(dim=1), sum). */
int sqlx_sql_pred_fwd(const float* x, const float* queries, const float* Wp /*[D,Q]*/, const float* bp /*[D]*/,
                      const float* centers /*[B,D]*/, int B, int E, int Q, int D, int n,
                      float* pred /*[B,n]*/, void* stream);

/* Backward pass 1 (pixel reductions):  d_centers [B,D], d_Wp [D,Q], d_bp [D]  (all overwritten) */
int sqlx_sql_bwd_reduce(const float* x, const float* queries, const float* Wp, const float* bp,
                        const float* centers, const float* pred, const float* g_pred /*[B,n]*/,
                        int B, int E, int Q, int D, int n,
                        float* d_centers, float* d_Wp, float* d_bp,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Backward pass 2: d_x [B,E,n] (overwritten), d_queries [B,Q,E] (overwritten).
 *   d_summary [B,Q,E] comes from autograd through the bins MLP (kept in PyTorch);
 *   g_energy optional [B,Q,n]: extra upstream gradient on the energy maps (module-level FullQueryLayer);
 *   g_pred may be NULL (then Wp/bp/centers/pred are ignored: pure FullQueryLayer backward). */
int sqlx_sql_bwd_dx(const float* x, const float* queries, const float* Wp, const float* bp,
                    const float* centers, const float* pred, const float* g_pred,
                    const float* summary, const float* row_max, const float* row_sum,
                    const float* d_summary, const float* g_energy,
                    int B, int E, int Q, int D, int n,
                    float* d_x, float* d_queries, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SQLX_H_ */
