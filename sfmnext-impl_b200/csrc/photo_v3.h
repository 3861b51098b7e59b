// Internal (non-ABI) entry points of photo_v3.cu used by multiscale.cu.
#pragma once
#include "common.cuh"

namespace sqlx {
// launches photo_fwd3_kernel; per-CTA partial sums of the per-pixel minimum land in partial[0 .. *ctas)
// depth_up: optional [B,H,W] already-upsampled depth (read instead of upsampling depth_lr per pixel)
int photo_fwd3_launch(const sqlx_photo_desc* desc, const float* depth_lr, const float* depth_up, const float* target,
                      const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                      const float* identity, const float* noise, float* partial, int* ctas, uint8_t* argmin,
                      float* ssim_coef, cudaStream_t st,
                      // indoor variant: depth of every source frame [B,H,W] + per-CTA sums of the regularisation term
                      const float* const* ref_depths = nullptr, float* partial_reg = nullptr);
// All loss scales in ONE launch: every CTA stages its target tile, the tile's box statistics and the identity losses
// once and runs the scales back to back.  depth_up[i] [B,H,W] (upsampled depth of scale i), noise[i], argmin[i];
// T + i * T_stride, partial + i * partial_stride, ssim_coef + i * coef_stride (strides in floats).
int photo_fwd3_ms_launch(const sqlx_photo_desc* desc, int ns, const float* const* depth_up, const float* target,
                         const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                         size_t T_stride, const float* identity, const float* const* noise, float* partial,
                         size_t partial_stride, int* ctas, uint8_t* const* argmin, float* ssim_coef, size_t coef_stride,
                         cudaStream_t st);
// launches photo_bwd3_kernel; exactly one of d_depth_lr (atomic upsample adjoint) / g_up (per-pixel plane) is given
int photo_bwd3_launch(const sqlx_photo_desc* desc, const float* depth_lr, const float* depth_up, const float* target,
                      const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                      const uint8_t* argmin, const float* ssim_coef, const float* g_loss, float scale,
                      float* d_depth_lr, float* g_up, float* q_up /*optional 1/d_up^2 plane*/, int g_up_accumulate, float* dP,
                      cudaStream_t st,
                      // indoor variant: source depths, their gradients (atomically accumulated), d loss / d reg-sum
                      const float* const* ref_depths = nullptr, float* const* d_ref_depths = nullptr,
                      const float* g_reg = nullptr);
// All loss scales in ONE backward launch (blockIdx.z = scale * B + sample): g_up[i] / q_up[i] [B,H,W] per scale (q_up
// entries may be NULL), bit i of accumulate_mask: g_up[i] += instead of =; T, ssim_coef and dP advance by their strides
// (in floats) per scale.
int photo_bwd3_ms_launch(const sqlx_photo_desc* desc, int ns, const float* const* depth_up, const float* target,
                         const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                         size_t T_stride, const uint8_t* const* argmin, const float* ssim_coef, size_t coef_stride,
                         const float* g_loss, float scale, float* const* g_up, float* const* q_up, unsigned accumulate_mask,
                         float* dP, size_t dP_stride, cudaStream_t st);
size_t photo_max_ctas(const sqlx_photo_desc* d);   // upper bound of *ctas for any tile configuration
}  // namespace sqlx
