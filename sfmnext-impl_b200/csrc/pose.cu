// Pose-matrix assembly: T = transformation_from_parameters(axisangle, translation*scale, invert)
// Reference: layers.py:75-150 (Rodrigues rotation, translation matrix, M = T*R or R^T*T(-t)) and the
// mean-inverse-depth rescale of the translation, trainer.py:417-421.  The reference spends ~60 tiny
// kernels per call on this; here it is one launch forward and one backward (forward-mode duals).
#include "common.cuh"

namespace sqlx {

template <typename S>
struct PoseOps;

struct Dual {
  float v, d;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }

template <>
struct PoseOps<float> {
  static __device__ __forceinline__ float c(float x) { return x; }
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
  static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
};
template <>
struct PoseOps<Dual> {
  static __device__ __forceinline__ Dual c(float x) { return {x, 0.f}; }
  static __device__ __forceinline__ Dual sqrt_(Dual x) {
    const float s = sqrtf(x.v);
    return {s, s > 0.f ? 0.5f * x.d / s : 0.f};
  }
  static __device__ __forceinline__ Dual sin_(Dual x) { return {sinf(x.v), cosf(x.v) * x.d}; }
  static __device__ __forceinline__ Dual cos_(Dual x) { return {cosf(x.v), -sinf(x.v) * x.d}; }
};

// M[16] row-major
template <typename S>
__device__ __forceinline__ void pose_eval(const S aa[3], const S tr[3], S scale, bool invert, S M[16]) {
  using O = PoseOps<S>;
  const S angle = O::sqrt_(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  const S den = angle + O::c(1e-7f);
  const S x = aa[0] / den, y = aa[1] / den, z = aa[2] / den;
  const S ca = O::cos_(angle), sa = O::sin_(angle);
  const S C = O::c(1.f) - ca;
  const S xs = x * sa, ys = y * sa, zs = z * sa;
  const S xC = x * C, yC = y * C, zC = z * C;
  const S xyC = x * yC, yzC = y * zC, zxC = z * xC;
  S R[9];
  R[0] = x * xC + ca; R[1] = xyC - zs;    R[2] = zxC + ys;
  R[3] = xyC + zs;    R[4] = y * yC + ca; R[5] = yzC - xs;
  R[6] = zxC - ys;    R[7] = yzC + xs;    R[8] = z * zC + ca;
  S t[3] = {tr[0] * scale, tr[1] * scale, tr[2] * scale};
  const S zero = O::c(0.f), one = O::c(1.f);
  if (!invert) {
    // M = T * R = [R | t]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      M[i * 4 + 0] = R[i * 3 + 0]; M[i * 4 + 1] = R[i * 3 + 1]; M[i * 4 + 2] = R[i * 3 + 2]; M[i * 4 + 3] = t[i];
    }
  } else {
    // M = R^T * T(-t) = [R^T | -R^T t]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      M[i * 4 + 0] = R[0 * 3 + i]; M[i * 4 + 1] = R[1 * 3 + i]; M[i * 4 + 2] = R[2 * 3 + i];
      M[i * 4 + 3] = R[0 * 3 + i] * (-t[0]) + R[1 * 3 + i] * (-t[1]) + R[2 * 3 + i] * (-t[2]);
    }
  }
  M[12] = zero; M[13] = zero; M[14] = zero; M[15] = one;
}

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
