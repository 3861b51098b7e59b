// Pose-matrix assembly: T = transformation_from_parameters(axisangle, translation*scale, invert)
// Reference: layers.py:75-150 (Rodrigues rotation, translation matrix, M = T*R or R^T*T(-t)) and the
// mean-inverse-depth rescale of the translation, trainer.py:417-421.  The reference spends ~60 tiny
// kernels per call on this; here it is one launch forward and one backward (forward-mode duals).
#include "pose.cuh"

namespace sqlx {

__global__ void pose_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr,
                                const float* __restrict__ scale, int B, int invert, float* __restrict__ T) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float a[3] = {aa[b * 3], aa[b * 3 + 1], aa[b * 3 + 2]};
  const float t[3] = {tr[b * 3], tr[b * 3 + 1], tr[b * 3 + 2]};
  float M[16];
  pose_eval<float>(a, t, scale ? scale[b] : 1.f, invert != 0, M);
#pragma unroll
  for (int i = 0; i < 16; ++i) T[b * 16 + i] = M[i];
}

// one thread per (sample, input j): j = 0..2 axisangle, 3..5 translation, 6 scale
__global__ void pose_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr,
                                const float* __restrict__ scale, int B, int invert, const float* __restrict__ dT,
                                float* __restrict__ d_aa, float* __restrict__ d_tr, float* __restrict__ d_scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 7) return;
  const int b = idx / 7, j = idx - b * 7;
  if (j == 6 && !d_scale) return;
  Dual a[3], t[3], sc;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a[i] = {aa[b * 3 + i], j == i ? 1.f : 0.f};
    t[i] = {tr[b * 3 + i], j == 3 + i ? 1.f : 0.f};
  }
  sc = {scale ? scale[b] : 1.f, j == 6 ? 1.f : 0.f};
  Dual M[16];
  pose_eval<Dual>(a, t, sc, invert != 0, M);
  float g = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) g += dT[b * 16 + i] * M[i].d;
  if (j < 3) d_aa[b * 3 + j] = g;
  else if (j < 6) d_tr[b * 3 + (j - 3)] = g;
  else d_scale[b] = g;
}

}  // namespace sqlx

using namespace sqlx;

extern "C" int sqlx_pose_fwd(const float* axisangle, const float* translation, const float* scale, int B, int invert,
                             float* T, void* stream) {
  SQLX_REQUIRE(axisangle && translation && T, "NULL pointer argument");
  SQLX_REQUIRE(B > 0, "non-positive batch");
  pose_fwd_kernel<<<ceil_div(B, 64), 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(axisangle, translation, scale, B,
                                                                                     invert, T);
  return check_launch("pose_fwd_kernel");
}

extern "C" int sqlx_pose_bwd(const float* axisangle, const float* translation, const float* scale, int B, int invert,
                             const float* dT, float* d_axisangle, float* d_translation, float* d_scale, void* stream) {
  SQLX_REQUIRE(axisangle && translation && dT && d_axisangle && d_translation, "NULL pointer argument");
  SQLX_REQUIRE(B > 0, "non-positive batch");
  SQLX_REQUIRE(!scale == !d_scale, "scale and d_scale must both be given or both be NULL");
  pose_bwd_kernel<<<ceil_div(B * 7, 64), 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      axisangle, translation, scale, B, invert, dT, d_axisangle, d_translation, d_scale);
  return check_launch("pose_bwd_kernel");
}
