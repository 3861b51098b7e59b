// Pose-matrix evaluation shared by pose.cu (module-level op) and scale_loss.cu (fused per-scale loss):
// Rodrigues rotation + translation, forward value and forward-mode derivative (duals).  layers.py:75-150.
#pragma once
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

}  // namespace sqlx
