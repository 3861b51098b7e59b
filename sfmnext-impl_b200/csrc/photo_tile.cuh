// Tile-level building blocks of the photometric kernels: region staging, warping into shared memory,
// separable (2R+1)^2 box sums, SSIM value and SSIM adjoint coefficients.
#pragma once
#include "common.cuh"

namespace sqlx {

__device__ __forceinline__ int clamp_reflect(int i, int n) {
  i = reflect_index(i, n);
  return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}

constexpr float kC1 = 0.01f * 0.01f;   // layers.py:28
constexpr float kC2 = 0.03f * 0.03f;   // layers.py:29

// Stage a (RH x RW) region of one image plane whose top-left corner sits at image position (v0,u0)
// (possibly negative); out-of-frame positions take the reflected pixel (ReflectionPad2d semantics).
template <int RH, int RW, int LD>
__device__ __forceinline__ void stage_plane(const float* __restrict__ plane, int H, int W, int v0, int u0,
                                            float* __restrict__ dst) {
  for (int idx = threadIdx.x; idx < RH * RW; idx += blockDim.x) {
    const int lr = idx / RW, lc = idx - lr * RW;
    const int v = clamp_reflect(v0 + lr, H), u = clamp_reflect(u0 + lc, W);
    dst[lr * LD + lc] = __ldg(plane + (size_t)v * W + u);
  }
}

// Upsampled depth (align_corners=False bilinear from the [h,w] map) on the same kind of region.
template <int RH, int RW, int LD>
__device__ __forceinline__ void stage_depth(const float* __restrict__ lr_map, int h, int w, int H, int W,
                                            int v0, int u0, float* __restrict__ dst) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (int idx = threadIdx.x; idx < RH * RW; idx += blockDim.x) {
    const int lr = idx / RW, lc = idx - lr * RW;
    const int v = clamp_reflect(v0 + lr, H), u = clamp_reflect(u0 + lc, W);
    dst[lr * LD + lc] = upsample_at(lr_map, h, w, v, u, sy, sx);
  }
}

// Warp one source frame into three shared-memory planes over the region.
template <int RH, int RW, int LD>
__device__ __forceinline__ void stage_warped(const float* __restrict__ src /*[3,H,W]*/, const Camera& cam,
                                             const float* __restrict__ dpl, int H, int W, int v0, int u0,
                                             float eps, float* __restrict__ w0, float* __restrict__ w1,
                                             float* __restrict__ w2) {
  const size_t plane = (size_t)H * W;
  for (int idx = threadIdx.x; idx < RH * RW; idx += blockDim.x) {
    const int lr = idx / RW, lc = idx - lr * RW;
    const int v = clamp_reflect(v0 + lr, H), u = clamp_reflect(u0 + lc, W);
    const int o = lr * LD + lc;
    const Sample sp = project_pixel(cam, (float)u, (float)v, dpl[o], H, W, eps);
    const Taps t = make_taps(sp.ix, sp.iy, H, W);
    const float* p = src;
    w0[o] = __ldg(p + t.o00) * t.w00 + __ldg(p + t.o01) * t.w01 + __ldg(p + t.o10) * t.w10 + __ldg(p + t.o11) * t.w11;
    p += plane;
    w1[o] = __ldg(p + t.o00) * t.w00 + __ldg(p + t.o01) * t.w01 + __ldg(p + t.o10) * t.w10 + __ldg(p + t.o11) * t.w11;
    p += plane;
    w2[o] = __ldg(p + t.o00) * t.w00 + __ldg(p + t.o01) * t.w01 + __ldg(p + t.o10) * t.w10 + __ldg(p + t.o11) * t.w11;
  }
}

// Horizontal (2R+1)-tap sums of {x, y, x^2, y^2, xy} for ROWS rows and OUTW output columns.
//   in planes have leading dimension LD and (OUTW + 2R) valid columns; out planes have leading dimension OLD.
template <int R, int ROWS, int OUTW, int LD, int OLD, bool WITH_Y>
__device__ __forceinline__ void hpass5(const float* __restrict__ X, const float* __restrict__ Y,
                                       float* __restrict__ hx, float* __restrict__ hxx, float* __restrict__ hxy,
                                       float* __restrict__ hy, float* __restrict__ hyy) {
  for (int idx = threadIdx.x; idx < ROWS * OUTW; idx += blockDim.x) {
    const int r = idx / OUTW, c = idx - r * OUTW;
    const float* xr = X + r * LD + c;
    const float* yr = Y + r * LD + c;
    float sx = 0.f, sxx = 0.f, sxy = 0.f, sy = 0.f, syy = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) {
      const float a = xr[k], b = yr[k];
      sx += a;
      sxx = fmaf(a, a, sxx);
      sxy = fmaf(a, b, sxy);
      if (WITH_Y) {
        sy += b;
        syy = fmaf(b, b, syy);
      }
    }
    const int o = r * OLD + c;
    hx[o] = sx; hxx[o] = sxx; hxy[o] = sxy;
    if (WITH_Y) { hy[o] = sy; hyy[o] = syy; }
  }
}

template <int R, int OLD>
__device__ __forceinline__ float vsum(const float* __restrict__ hb, int row, int col) {
  const float* p = hb + row * OLD + col;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k <= 2 * R; ++k) s += p[k * OLD];
  return s;
}

// ---- register-blocked variants (4 horizontal outputs per work item, PPT vertically adjacent outputs per thread):
// 40 % fewer shared-memory wavefronts than the one-output-per-thread versions above.
// Horizontal: item = (row, 4 consecutive output columns); 10 inputs fetched as five 8-byte loads per plane.
// LD must be even (8-byte aligned rows); LD = 42 keeps the 8-byte accesses of a half-warp on distinct banks.
template <int R, int ROWS, int OUTW, int LD, int OLD, bool WITH_X>
__device__ __forceinline__ void hpass_blocked(const float* __restrict__ X, const float* __restrict__ Y,
                                              float* __restrict__ h0, float* __restrict__ h1, float* __restrict__ h2) {
  static_assert(R == 3 || R == 1, "window radius");
  static_assert(OUTW % 4 == 0 && LD % 2 == 0 && OLD % 4 == 0, "blocked horizontal pass alignment");
  constexpr int NIN = 4 + 2 * R;          // inputs per item
  constexpr int ITEMS = ROWS * (OUTW / 4);
  for (int idx = threadIdx.x; idx < ITEMS; idx += blockDim.x) {
    const int r = idx / (OUTW / 4), c = (idx - r * (OUTW / 4)) * 4;
    float y[NIN], x[NIN];
    const float2* yp = reinterpret_cast<const float2*>(Y + r * LD + c);
#pragma unroll
    for (int k = 0; k < NIN / 2; ++k) { const float2 v = yp[k]; y[2 * k] = v.x; y[2 * k + 1] = v.y; }
    if (WITH_X) {
      const float2* xp = reinterpret_cast<const float2*>(X + r * LD + c);
#pragma unroll
      for (int k = 0; k < NIN / 2; ++k) { const float2 v = xp[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
    }
    float o0[4], o1[4], o2[4];
    // WITH_X: (sum x, sum x^2, sum x y); otherwise (sum y, sum y^2)
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) {
      const float u = WITH_X ? x[k] : y[k];
      a += u;
      b = fmaf(u, u, b);
      if (WITH_X) d = fmaf(u, y[k], d);
    }
    o0[0] = a; o1[0] = b; o2[0] = d;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      const float un = WITH_X ? x[j + 2 * R] : y[j + 2 * R], uo = WITH_X ? x[j - 1] : y[j - 1];
      a += un - uo;
      b += un * un - uo * uo;
      if (WITH_X) d += un * y[j + 2 * R] - uo * y[j - 1];
      o0[j] = a; o1[j] = b; o2[j] = d;
    }
    const int o = r * OLD + c;
    *reinterpret_cast<float4*>(h0 + o) = make_float4(o0[0], o0[1], o0[2], o0[3]);
    *reinterpret_cast<float4*>(h1 + o) = make_float4(o1[0], o1[1], o1[2], o1[3]);
    if (WITH_X) *reinterpret_cast<float4*>(h2 + o) = make_float4(o2[0], o2[1], o2[2], o2[3]);
  }
}

// Vertical (2R+1)-sums for N vertically adjacent outputs starting at row `row`: N + 2R loads instead of N (2R+1)
template <int R, int OLD, int N>
__device__ __forceinline__ void vsum_multi(const float* __restrict__ hb, int row, int col, float (&out)[N]) {
  const float* p = hb + row * OLD + col;
  float v[N + 2 * R];
#pragma unroll
  for (int k = 0; k < N + 2 * R; ++k) v[k] = p[k * OLD];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k <= 2 * R; ++k) s += v[k];
  out[0] = s;
#pragma unroll
  for (int j = 1; j < N; ++j) {
    s += v[j + 2 * R] - v[j - 1];
    out[j] = s;
  }
}

struct SsimStats {
  float mx, my, sxx, syy, sxy;  // means and (co)variances
};
template <int R>
__device__ __forceinline__ SsimStats make_stats(float Sx, float Sy, float Sxx, float Syy, float Sxy) {
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  SsimStats s;
  s.mx = Sx * ia; s.my = Sy * ia;
  s.sxx = Sxx * ia - s.mx * s.mx;
  s.syy = Syy * ia - s.my * s.my;
  s.sxy = Sxy * ia - s.mx * s.my;
  return s;
}
// layers.py:42-46
__device__ __forceinline__ float ssim_value(const SsimStats& s) {
  const float n = (2.f * s.mx * s.my + kC1) * (2.f * s.sxy + kC2);
  const float d = (s.mx * s.mx + s.my * s.my + kC1) * (s.sxx + s.syy + kC2);
  const float v = (1.f - n / d) * 0.5f;
  return fminf(fmaxf(v, 0.f), 1.f);
}
// d(ssim)/d(mean_x), d/d(E[x^2]), d/d(E[xy]) (total derivatives through the (co)variances), and
// optionally the same for y.  Zero where the clamp is active (torch.clamp backward).
struct SsimGrad {
  float dmx, dexx, dexy, dmy, deyy;
};
__device__ __forceinline__ SsimGrad ssim_grad(const SsimStats& s) {
  const float n1 = 2.f * s.mx * s.my + kC1, n2 = 2.f * s.sxy + kC2;
  const float d1 = s.mx * s.mx + s.my * s.my + kC1, d2 = s.sxx + s.syy + kC2;
  const float n = n1 * n2, d = d1 * d2;
  const float v = (1.f - n / d) * 0.5f;
  SsimGrad g;
  if (!(v >= 0.f && v <= 1.f)) {
    g.dmx = g.dexx = g.dexy = g.dmy = g.deyy = 0.f;
    return g;
  }
  const float inv_d = 1.f / d;
  const float dS_dn = -0.5f * inv_d;            // dS/dn
  const float dS_dd = 0.5f * n * inv_d * inv_d; // dS/dd
  g.dmx = dS_dn * (2.f * s.my * (n2 - n1)) + dS_dd * (2.f * s.mx * (d2 - d1));
  g.dmy = dS_dn * (2.f * s.mx * (n2 - n1)) + dS_dd * (2.f * s.my * (d2 - d1));
  g.dexx = dS_dd * d1;
  g.deyy = dS_dd * d1;
  g.dexy = dS_dn * 2.f * n1;
  return g;
}

}  // namespace sqlx
