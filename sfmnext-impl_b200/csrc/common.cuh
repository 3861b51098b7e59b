// Shared device/host helpers for libsqlx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/sqlx.h"

namespace sqlx {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SQLX_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      sqlx::set_error(__VA_ARGS__);        \
      return SQLX_EINVAL;                  \
    }                                      \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function attribute: set it once per (kernel, current
// device), thread-safe (core.cu).  The library is entered from the training thread and from autograd's per-device
// backward threads, and one process may drive several GPUs (nn.DataParallel).  Returns SQLX_OK or SQLX_ECUDA.
int ensure_dyn_smem(const void* kernel, size_t bytes);
template <class K>
inline int ensure_dyn_smem(K* kernel, size_t bytes) { return ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes); }

// Times the enclosed launches with CUDA events on `st` while sqlx_profile_enable(1) is in effect (core.cu).
class ProfScope {
 public:
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;

 private:
  int idx_;
  cudaStream_t st_;
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// index reflection of nn.ReflectionPad2d: -1 -> 1, n -> n-2  (layers.py:26)
__device__ __forceinline__ int reflect_index(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

// source index rule of F.interpolate(mode="bilinear", align_corners=False)
// (ATen UpSample.h area_pixel_compute_source_index): src = scale*(dst+0.5)-0.5, clamped at 0.
struct UpTap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ UpTap up_tap(int dst, float scale, int in_size) {
  float src = scale * (dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  int i0 = (int)src;
  i0 = i0 > in_size - 1 ? in_size - 1 : i0;
  UpTap t;
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  t.l1 = src - (float)i0;
  t.l0 = 1.f - t.l1;
  return t;
}

__device__ __forceinline__ float upsample_at(const float* __restrict__ lr, int h, int w, int v, int u,
                                             float sy, float sx) {
  UpTap ty = up_tap(v, sy, h), tx = up_tap(u, sx, w);
  const float* r0 = lr + (size_t)ty.i0 * w;
  const float* r1 = lr + (size_t)ty.i1 * w;
  return ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1)) +
         ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum, result valid in thread 0.  `red` needs >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// Camera model for one (sample, source): P = (K*T)[:3,:] and the 3x3 of inv_K, laid out in registers/smem.
struct Camera {
  float P[12];   // row-major 3x4
  float iK[9];   // row-major 3x3
};

__device__ __forceinline__ void load_camera(const float* __restrict__ K, const float* __restrict__ invK,
                                            const float* __restrict__ T, Camera& cam) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(K[i * 4 + k], T[k * 4 + j], acc);
      cam.P[i * 4 + j] = acc;
    }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) cam.iK[i * 3 + j] = invK[i * 4 + j];
}

// Projected sampling position of pixel (u,v) with depth d.  Mirrors layers.py:210-215 (backproject),
// :247-257 (project + normalise) and the un-normalise + border clip of F.grid_sample(align_corners=True,
// padding_mode="border") (ATen GridSampler.h).
struct Sample {
  float ix, iy;        // clipped source coordinates
  float gx, gy;        // normalised grid (what outputs[("sample",f,s)] holds)
  float X[3];          // camera point
  float z;             // cam z + eps
  float pu, pv;        // unclipped projected pixel coordinates (u', v')
  bool in_x, in_y;     // gradient passes (not clipped)
};
__device__ __forceinline__ Sample project_pixel(const Camera& cam, float u, float v, float d, int H, int W,
                                                float eps) {
  Sample s;
  const float r0 = cam.iK[0] * u + cam.iK[1] * v + cam.iK[2];
  const float r1 = cam.iK[3] * u + cam.iK[4] * v + cam.iK[5];
  const float r2 = cam.iK[6] * u + cam.iK[7] * v + cam.iK[8];
  s.X[0] = d * r0; s.X[1] = d * r1; s.X[2] = d * r2;
  const float c0 = cam.P[0] * s.X[0] + cam.P[1] * s.X[1] + cam.P[2] * s.X[2] + cam.P[3];
  const float c1 = cam.P[4] * s.X[0] + cam.P[5] * s.X[1] + cam.P[6] * s.X[2] + cam.P[7];
  const float c2 = cam.P[8] * s.X[0] + cam.P[9] * s.X[1] + cam.P[10] * s.X[2] + cam.P[11];
  s.z = c2 + eps;
  s.pu = c0 / s.z;
  s.pv = c1 / s.z;
  s.gx = (s.pu / (float)(W - 1) - 0.5f) * 2.f;
  s.gy = (s.pv / (float)(H - 1) - 0.5f) * 2.f;
  float ix = ((s.gx + 1.f) / 2.f) * (float)(W - 1);
  float iy = ((s.gy + 1.f) / 2.f) * (float)(H - 1);
  s.in_x = ix > 0.f && ix < (float)(W - 1);
  s.in_y = iy > 0.f && iy < (float)(H - 1);
  s.ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  s.iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  return s;
}

// Border-clamped bilinear taps of one plane (same weight formulas as ATen grid_sampler_2d).
struct Taps {
  int o00, o01, o10, o11;   // offsets into a plane
  float w00, w01, w10, w11; // nw, ne, sw, se
  float fx, fy;
};
__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W) {
  Taps t;
  const float x0f = floorf(ix), y0f = floorf(iy);
  int x0 = (int)x0f, y0 = (int)y0f;
  t.fx = ix - x0f;
  t.fy = iy - y0f;
  const int x1 = x0 + 1 < W ? x0 + 1 : x0;  // when ix == W-1 the ne/se weights are 0 (ATen skips them)
  const int y1 = y0 + 1 < H ? y0 + 1 : y0;
  t.w00 = (1.f - t.fx) * (1.f - t.fy);
  t.w01 = t.fx * (1.f - t.fy);
  t.w10 = (1.f - t.fx) * t.fy;
  t.w11 = t.fx * t.fy;
  t.o00 = y0 * W + x0; t.o01 = y0 * W + x1; t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
  return t;
}

}  // namespace sqlx
