// Photometric backward: gradient of the per-scale photometric loss wrt the low-resolution depth map and
// the camera transforms T.  Recomputes the forward (warp + box sums) on a tile with a 2R halo, forms the
// SSIM adjoint coefficients on the R halo, box-filters them back (with reflect-pad multiplicities) and
// chains through bilinear sampling -> projection -> back-projection -> bilinear upsampling.
// Reference: autograd of layers.py:31-46,210-215,247-258 + trainer.py:395-396,431-435,444-451,526-532.
#include "photo_tile.cuh"

#include <stdlib.h>

namespace sqlx {

struct PhotoBwdParams {
  sqlx_photo_desc d;
  const float* depth_lr;
  const float* target;
  const float* src[SQLX_MAX_SOURCES];
  const float* K;
  const float* invK;
  const float* T;
  const uint8_t* argmin;
  const float* g_loss;
  float scale;
  float* d_depth_lr;
  float* dP;  // [B,S,12] accumulators (zeroed by the host wrapper)
};

template <int R, int TH, int TW, int NT>
struct BwdCfg {
  static constexpr int PH2 = TH + 4 * R, PW2 = TW + 4 * R;
  static constexpr int LD2 = (PW2 + 3) & ~3;
  static constexpr int PLANE2 = PH2 * LD2;
  static constexpr int PH1 = TH + 2 * R, PW1 = TW + 2 * R;
  static constexpr int HB = PH2 * PW1;   // forward H-pass planes
  static constexpr int CF = PH1 * PW1;   // coefficient planes
  static constexpr int PPT = (TH * TW) / NT;
  static constexpr int LRH = TH + 2, LRW = TW + 2;
  static constexpr int HBTOT = (5 * HB > LRH * LRW) ? 5 * HB : LRH * LRW;
  static constexpr size_t smem_bytes =
      sizeof(float) * (7 * PLANE2 + HBTOT + 3 * CF + 32 + 16) + sizeof(Camera) * SQLX_MAX_SOURCES + CF /*argmin u8*/ + 16;
};

// multiplicity with which the window of output q (coordinate qi) covers pixel i under reflection padding
template <int R>
__device__ __forceinline__ float reflect_mult(int i, int qi, int n) {
  float m = 1.f;  // |qi - i| <= R is guaranteed by the caller's loop bounds
  if (i >= 1 && i <= R && qi + i <= R) m += 1.f;
  if (i <= n - 2 && i >= n - 1 - R && 2 * (n - 1) - i - qi <= R) m += 1.f;
  return m;
}

template <int R, int TH, int TW, int NT>
__global__ void __launch_bounds__(NT) photo_bwd_kernel(const PhotoBwdParams p) {
  using C = BwdCfg<R, TH, TW, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* dpl = reinterpret_cast<float*>(smem_raw);
  float* tg = dpl + C::PLANE2;
  float* wp = tg + 3 * C::PLANE2;
  float* hb = wp + 3 * C::PLANE2;
  float* cf = hb + C::HBTOT;
  float* red = cf + 3 * C::CF;
  float* dPs = red + 32;  // 12 (+pad)
  Camera* cams = reinterpret_cast<Camera*>(dPs + 16);
  uint8_t* amin = reinterpret_cast<uint8_t*>(cams + SQLX_MAX_SOURCES);

  const int H = p.d.H, W = p.d.W, S = p.d.S;
  const int b = blockIdx.z;
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  const bool automask = p.d.flags & SQLX_AUTOMASK;
  const bool avg = p.d.flags & SQLX_AVG_REPROJ;
  const int n_ident = automask ? (avg ? 1 : S) : 0;
  const float gscale = __ldg(p.g_loss) * p.scale;
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));

  if (threadIdx.x < S)
    load_camera(p.K + b * 16, p.invK + b * 16, p.T + ((size_t)b * S + threadIdx.x) * 16, cams[threadIdx.x]);
  stage_depth<C::PH2, C::PW2, C::LD2>(p.depth_lr + (size_t)b * p.d.h * p.d.w, p.d.h, p.d.w, H, W, v0 - 2 * R,
                                      u0 - 2 * R, dpl);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    stage_plane<C::PH2, C::PW2, C::LD2>(p.target + ((size_t)b * 3 + c) * plane, H, W, v0 - 2 * R, u0 - 2 * R,
                                        tg + c * C::PLANE2);
  // argmin on the R halo (255 = out of frame)
  for (int idx = threadIdx.x; idx < C::CF; idx += NT) {
    const int lr = idx / C::PW1, lc = idx - lr * C::PW1;
    const int v = v0 - R + lr, u = u0 - R + lc;
    amin[idx] = (v >= 0 && v < H && u >= 0 && u < W) ? p.argmin[(size_t)b * plane + (size_t)v * W + u] : (uint8_t)255;
  }
  __syncthreads();

  int prow[C::PPT], pcol[C::PPT];
  bool pin[C::PPT];
  float gd[C::PPT];
#pragma unroll
  for (int k = 0; k < C::PPT; ++k) {
    const int pix = threadIdx.x + k * NT;
    prow[k] = pix / TW;
    pcol[k] = pix - prow[k] * TW;
    pin[k] = (v0 + prow[k] < H) && (u0 + pcol[k] < W);
    gd[k] = 0.f;
  }

  for (int s = 0; s < S; ++s) {
    const float sel_w = avg ? 1.f / (float)S : 1.f;
    const int sel_idx = avg ? n_ident : n_ident + s;
    stage_warped<C::PH2, C::PW2, C::LD2>(p.src[s] + (size_t)b * 3 * plane, cams[s], dpl, H, W, v0 - 2 * R, u0 - 2 * R,
                                         p.d.eps, wp, wp + C::PLANE2, wp + 2 * C::PLANE2);
    __syncthreads();
    float gx[3][C::PPT];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* X = wp + c * C::PLANE2;
      const float* Y = tg + c * C::PLANE2;
      // L1 term (local)
#pragma unroll
      for (int k = 0; k < C::PPT; ++k) {
        const int o = (prow[k] + 2 * R) * C::LD2 + pcol[k] + 2 * R;
        const float diff = Y[o] - X[o];
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const bool sel = amin[(prow[k] + R) * C::PW1 + pcol[k] + R] == sel_idx;
        const float wl1 = (R > 0) ? p.d.w_l1 : 1.f;
        gx[c][k] = sel ? -(wl1 / 3.f) * sel_w * gscale * sgn : 0.f;
      }
      if (R > 0) {
        hpass5<R, C::PH2, C::PW1, C::LD2, C::PW1, true>(X, Y, hb, hb + C::HB, hb + 2 * C::HB, hb + 3 * C::HB,
                                                        hb + 4 * C::HB);
        __syncthreads();
        // SSIM adjoint coefficients on the R halo
        const float alpha = (p.d.w_ssim / 3.f) * sel_w * gscale;
        for (int idx = threadIdx.x; idx < C::CF; idx += NT) {
          const int lr = idx / C::PW1, lc = idx - lr * C::PW1;
          float a = 0.f, bb = 0.f, cc = 0.f;
          if (amin[idx] == sel_idx) {
            const float Sx = vsum<R, C::PW1>(hb, lr, lc);
            const float Sxx = vsum<R, C::PW1>(hb + C::HB, lr, lc);
            const float Sxy = vsum<R, C::PW1>(hb + 2 * C::HB, lr, lc);
            const float Sy = vsum<R, C::PW1>(hb + 3 * C::HB, lr, lc);
            const float Syy = vsum<R, C::PW1>(hb + 4 * C::HB, lr, lc);
            const SsimGrad g = ssim_grad(make_stats<R>(Sx, Sy, Sxx, Syy, Sxy));
            a = alpha * g.dmx; bb = alpha * g.dexx; cc = alpha * g.dexy;
          }
          cf[idx] = a; cf[C::CF + idx] = bb; cf[2 * C::CF + idx] = cc;
        }
        __syncthreads();
        // adjoint horizontal pass: PH1 rows x TW tile columns, reflect multiplicities on frame borders
        float* h2 = hb;  // 3 planes of PH1*TW
        for (int idx = threadIdx.x; idx < C::PH1 * TW; idx += NT) {
          const int lr = idx / TW, pc = idx - lr * TW;
          const int u = u0 + pc;
          const float* ca = cf + lr * C::PW1 + pc;  // window starts at column pc (= pc+R-R)
          float sa = 0.f, sb = 0.f, sc = 0.f;
          if (u > R && u < W - 1 - R) {
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) { sa += ca[k]; sb += ca[C::CF + k]; sc += ca[2 * C::CF + k]; }
          } else if (u < W) {
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
              const int qu = u - R + k;
              if (qu < 0 || qu >= W) continue;
              const float m = reflect_mult<R>(u, qu, W);
              sa += m * ca[k]; sb += m * ca[C::CF + k]; sc += m * ca[2 * C::CF + k];
            }
          }
          h2[idx] = sa; h2[C::PH1 * TW + idx] = sb; h2[2 * C::PH1 * TW + idx] = sc;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < C::PPT; ++k) {
          const int v = v0 + prow[k];
          const float* ha = h2 + prow[k] * TW + pcol[k];  // window starts at row prow (= prow+R-R)
          float sa = 0.f, sb = 0.f, sc = 0.f;
          if (v > R && v < H - 1 - R) {
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) {
              sa += ha[j * TW]; sb += ha[C::PH1 * TW + j * TW]; sc += ha[2 * C::PH1 * TW + j * TW];
            }
          } else if (v < H) {
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) {
              const int qv = v - R + j;
              if (qv < 0 || qv >= H) continue;
              const float m = reflect_mult<R>(v, qv, H);
              sa += m * ha[j * TW]; sb += m * ha[C::PH1 * TW + j * TW]; sc += m * ha[2 * C::PH1 * TW + j * TW];
            }
          }
          const int o = (prow[k] + 2 * R) * C::LD2 + pcol[k] + 2 * R;
          gx[c][k] += ia * (sa + 2.f * X[o] * sb + Y[o] * sc);
        }
        __syncthreads();
      }
    }
    // chain rule through bilinear sampling and projection, for the owned pixels
    float dPacc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) dPacc[i] = 0.f;
    const Camera& cam = cams[s];
    const float* srcb = p.src[s] + (size_t)b * 3 * plane;
#pragma unroll
    for (int k = 0; k < C::PPT; ++k) {
      if (!pin[k]) continue;
      if (gx[0][k] == 0.f && gx[1][k] == 0.f && gx[2][k] == 0.f) continue;
      const int v = v0 + prow[k], u = u0 + pcol[k];
      const float d = dpl[(prow[k] + 2 * R) * C::LD2 + pcol[k] + 2 * R];
      const Sample sp = project_pixel(cam, (float)u, (float)v, d, H, W, p.d.eps);
      const Taps t = make_taps(sp.ix, sp.iy, H, W);
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pl = srcb + c * plane;
        const float v00 = __ldg(pl + t.o00), v01 = __ldg(pl + t.o01), v10 = __ldg(pl + t.o10), v11 = __ldg(pl + t.o11);
        gix += gx[c][k] * ((v01 - v00) * (1.f - t.fy) + (v11 - v10) * t.fy);
        giy += gx[c][k] * ((v10 - v00) * (1.f - t.fx) + (v11 - v01) * t.fx);
      }
      if (!sp.in_x) gix = 0.f;
      if (!sp.in_y) giy = 0.f;
      const float iz = 1.f / sp.z;
      const float g0 = gix * iz, g1 = giy * iz, g2 = -(gix * sp.pu + giy * sp.pv) * iz;
      // d cam_i / d depth = P[i,0:3] . r   with r = X / d  (recomputed from inv_K to avoid the division)
      const float r0 = cam.iK[0] * u + cam.iK[1] * v + cam.iK[2];
      const float r1 = cam.iK[3] * u + cam.iK[4] * v + cam.iK[5];
      const float r2 = cam.iK[6] * u + cam.iK[7] * v + cam.iK[8];
      gd[k] += g0 * (cam.P[0] * r0 + cam.P[1] * r1 + cam.P[2] * r2) +
               g1 * (cam.P[4] * r0 + cam.P[5] * r1 + cam.P[6] * r2) +
               g2 * (cam.P[8] * r0 + cam.P[9] * r1 + cam.P[10] * r2);
      dPacc[0] += g0 * sp.X[0]; dPacc[1] += g0 * sp.X[1]; dPacc[2] += g0 * sp.X[2]; dPacc[3] += g0;
      dPacc[4] += g1 * sp.X[0]; dPacc[5] += g1 * sp.X[1]; dPacc[6] += g1 * sp.X[2]; dPacc[7] += g1;
      dPacc[8] += g2 * sp.X[0]; dPacc[9] += g2 * sp.X[1]; dPacc[10] += g2 * sp.X[2]; dPacc[11] += g2;
    }
    if (threadIdx.x < 12) dPs[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float t = warp_sum(dPacc[i]);
      if ((threadIdx.x & 31) == 0 && t != 0.f) atomicAdd(&dPs[i], t);
    }
    __syncthreads();
    if (threadIdx.x < 12 && dPs[threadIdx.x] != 0.f)
      atomicAdd(p.dP + ((size_t)b * S + s) * 12 + threadIdx.x, dPs[threadIdx.x]);
    // (the next iteration's stage_warped is followed by a __syncthreads before anyone touches dPs again)
  }

  // adjoint of the bilinear upsampling: accumulate the tile's contribution per low-res cell in shared
  // memory, then one global atomic per touched cell.
  const int h = p.d.h, w = p.d.w;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int vend = min(v0 + TH, H) - 1, uend = min(u0 + TW, W) - 1;
  const int i_lo = up_tap(v0, sy, h).i0, i_hi = up_tap(vend, sy, h).i1;
  const int j_lo = up_tap(u0, sx, w).i0, j_hi = up_tap(uend, sx, w).i1;
  const int nh = i_hi - i_lo + 1, nw = j_hi - j_lo + 1;  // <= TH+2, TW+2 because h<=H, w<=W
  float* acc = hb;
  __syncthreads();
  for (int idx = threadIdx.x; idx < nh * nw; idx += NT) acc[idx] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < C::PPT; ++k) {
    if (!pin[k] || gd[k] == 0.f) continue;
    const UpTap ty = up_tap(v0 + prow[k], sy, h), tx = up_tap(u0 + pcol[k], sx, w);
    const int a0 = (ty.i0 - i_lo) * nw, a1 = (ty.i1 - i_lo) * nw, c0 = tx.i0 - j_lo, c1 = tx.i1 - j_lo;
    atomicAdd(acc + a0 + c0, gd[k] * ty.l0 * tx.l0);
    atomicAdd(acc + a0 + c1, gd[k] * ty.l0 * tx.l1);
    atomicAdd(acc + a1 + c0, gd[k] * ty.l1 * tx.l0);
    atomicAdd(acc + a1 + c1, gd[k] * ty.l1 * tx.l1);
  }
  __syncthreads();
  float* out = p.d_depth_lr + (size_t)b * h * w;
  for (int idx = threadIdx.x; idx < nh * nw; idx += NT) {
    const float g = acc[idx];
    if (g != 0.f) {
      const int i = idx / nw, j = idx - i * nw;
      atomicAdd(out + (size_t)(i_lo + i) * w + (j_lo + j), g);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward from SAVED SSIM coefficients (the forward kernel exports d SSIM / d(mean_x, E[x^2], E[xy]) per pixel,
// channel and source): no warp / box-sum recompute on a 2R halo.  Per tile: stage the three coefficient planes of
// (source, channel) on the R halo masked by the arg-min selection, adjoint box filter, then chain through bilinear
// sampling -> projection -> back-projection -> upsampling for the tile's own pixels only.
// ------------------------------------------------------------------------------------------------
template <int R, int TH, int TW, int NT>
struct Bwd2Cfg {
  static constexpr int PH1 = TH + 2 * R, PW1 = TW + 2 * R;
  static constexpr int CF = PH1 * PW1;
  static constexpr int PPT = (TH * TW) / NT;
  static constexpr int LRH = TH + 2, LRW = TW + 2;
  static constexpr int H2 = 3 * PH1 * TW;
  static constexpr int SCR = (H2 > LRH * LRW) ? H2 : LRH * LRW;
  static constexpr size_t smem_bytes = sizeof(float) * (3 * CF + SCR + 32 + 16) + sizeof(Camera) * SQLX_MAX_SOURCES + CF + 16;
};

template <int R, int TH, int TW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) photo_bwd2_kernel(const PhotoBwdParams p, const float* __restrict__ coef) {
  using C = Bwd2Cfg<R, TH, TW, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* cf = reinterpret_cast<float*>(smem_raw);   // 3 planes on the R halo
  float* h2 = cf + 3 * C::CF;                         // 3 planes PH1 x TW ; later the low-res accumulation scratch
  float* red = h2 + C::SCR;
  float* dPs = red + 32;
  Camera* cams = reinterpret_cast<Camera*>(dPs + 16);
  uint8_t* amin = reinterpret_cast<uint8_t*>(cams + SQLX_MAX_SOURCES);

  const int H = p.d.H, W = p.d.W, S = p.d.S;
  const int b = blockIdx.z;
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  const bool automask = p.d.flags & SQLX_AUTOMASK;
  const bool avg = p.d.flags & SQLX_AVG_REPROJ;
  const int n_ident = automask ? (avg ? 1 : S) : 0;
  const float gscale = __ldg(p.g_loss) * p.scale;
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  const int h = p.d.h, w = p.d.w;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;

  if (threadIdx.x < S)
    load_camera(p.K + b * 16, p.invK + b * 16, p.T + ((size_t)b * S + threadIdx.x) * 16, cams[threadIdx.x]);
  for (int idx = threadIdx.x; idx < C::CF; idx += NT) {
    const int lr = idx / C::PW1, lc = idx - lr * C::PW1;
    const int v = v0 - R + lr, u = u0 - R + lc;
    amin[idx] = (v >= 0 && v < H && u >= 0 && u < W) ? p.argmin[(size_t)b * plane + (size_t)v * W + u] : (uint8_t)255;
  }
  int prow[C::PPT], pcol[C::PPT];
  bool pin[C::PPT];
  float gd[C::PPT], dep[C::PPT], Yv[3][C::PPT];
#pragma unroll
  for (int k = 0; k < C::PPT; ++k) {
    const int pix = threadIdx.x + k * NT;
    prow[k] = pix / TW;
    pcol[k] = pix - prow[k] * TW;
    pin[k] = (v0 + prow[k] < H) && (u0 + pcol[k] < W);
    gd[k] = 0.f;
    dep[k] = 1.f;
    if (pin[k]) {
      dep[k] = upsample_at(p.depth_lr + (size_t)b * h * w, h, w, v0 + prow[k], u0 + pcol[k], sy, sx);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        Yv[c][k] = __ldg(p.target + ((size_t)b * 3 + c) * plane + (size_t)(v0 + prow[k]) * W + (u0 + pcol[k]));
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) Yv[c][k] = 0.f;
    }
  }
  __syncthreads();

  for (int s = 0; s < S; ++s) {
    const float sel_w = avg ? 1.f / (float)S : 1.f;
    const int sel_idx = avg ? n_ident : n_ident + s;
    {   // nothing on this tile's halo selects source s (auto-masked or won by another source): no gradient at all
      int any = 0;
      for (int idx = threadIdx.x; idx < C::CF; idx += NT) any |= (amin[idx] == sel_idx);
      if (!__syncthreads_or(any)) continue;
    }
    const Camera& cam = cams[s];
    const float* srcb = p.src[s] + (size_t)b * 3 * plane;
    // own pixels: projection, taps, warped value and its spatial derivatives per channel
    float Xv[3][C::PPT], dXx[3][C::PPT], dXy[3][C::PPT];
    Sample sp[C::PPT];
#pragma unroll
    for (int k = 0; k < C::PPT; ++k) {
      sp[k] = project_pixel(cam, (float)(u0 + pcol[k]), (float)(v0 + prow[k]), dep[k], H, W, p.d.eps);
      const Taps t = make_taps(sp[k].ix, sp[k].iy, H, W);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pl = srcb + c * plane;
        float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
        if (pin[k]) { v00 = __ldg(pl + t.o00); v01 = __ldg(pl + t.o01); v10 = __ldg(pl + t.o10); v11 = __ldg(pl + t.o11); }
        Xv[c][k] = v00 * t.w00 + v01 * t.w01 + v10 * t.w10 + v11 * t.w11;
        dXx[c][k] = (v01 - v00) * (1.f - t.fy) + (v11 - v10) * t.fy;
        dXy[c][k] = (v10 - v00) * (1.f - t.fx) + (v11 - v01) * t.fx;
      }
    }
    float gix[C::PPT], giy[C::PPT];
#pragma unroll
    for (int k = 0; k < C::PPT; ++k) { gix[k] = 0.f; giy[k] = 0.f; }
    const float alpha = (p.d.w_ssim / 3.f) * sel_w * gscale;
    const float wl1 = ((R > 0) ? p.d.w_l1 : 1.f) / 3.f * sel_w * gscale;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gx[C::PPT];
#pragma unroll
      for (int k = 0; k < C::PPT; ++k) {
        const float diff = Yv[c][k] - Xv[c][k];
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const bool sel = amin[(prow[k] + R) * C::PW1 + pcol[k] + R] == sel_idx;
        gx[k] = sel ? -wl1 * sgn : 0.f;
      }
      if (R > 0) {
        const float* cbase = coef + ((((size_t)b * S + s) * 3 + c) * 3) * plane;
        for (int idx = threadIdx.x; idx < C::CF; idx += NT) {
          float a = 0.f, bb = 0.f, cc = 0.f;
          if (amin[idx] == sel_idx) {
            const int lr = idx / C::PW1, lc = idx - lr * C::PW1;
            const size_t o = (size_t)(v0 - R + lr) * W + (u0 - R + lc);
            a = __ldg(cbase + o); bb = __ldg(cbase + plane + o); cc = __ldg(cbase + 2 * plane + o);
          }
          cf[idx] = a; cf[C::CF + idx] = bb; cf[2 * C::CF + idx] = cc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < C::PH1 * TW; idx += NT) {
          const int lr = idx / TW, pc = idx - lr * TW;
          const int u = u0 + pc;
          const float* ca = cf + lr * C::PW1 + pc;
          float sa = 0.f, sb = 0.f, sc = 0.f;
          if (u > R && u < W - 1 - R) {
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) { sa += ca[k]; sb += ca[C::CF + k]; sc += ca[2 * C::CF + k]; }
          } else if (u < W) {
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
              const int qu = u - R + k;
              if (qu < 0 || qu >= W) continue;
              const float m = reflect_mult<R>(u, qu, W);
              sa += m * ca[k]; sb += m * ca[C::CF + k]; sc += m * ca[2 * C::CF + k];
            }
          }
          h2[idx] = sa; h2[C::PH1 * TW + idx] = sb; h2[2 * C::PH1 * TW + idx] = sc;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < C::PPT; ++k) {
          const int v = v0 + prow[k];
          const float* ha = h2 + prow[k] * TW + pcol[k];
          float sa = 0.f, sb = 0.f, sc = 0.f;
          if (v > R && v < H - 1 - R) {
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) {
              sa += ha[j * TW]; sb += ha[C::PH1 * TW + j * TW]; sc += ha[2 * C::PH1 * TW + j * TW];
            }
          } else if (v < H) {
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) {
              const int qv = v - R + j;
              if (qv < 0 || qv >= H) continue;
              const float m = reflect_mult<R>(v, qv, H);
              sa += m * ha[j * TW]; sb += m * ha[C::PH1 * TW + j * TW]; sc += m * ha[2 * C::PH1 * TW + j * TW];
            }
          }
          gx[k] += alpha * ia * (sa + 2.f * Xv[c][k] * sb + Yv[c][k] * sc);
        }
        __syncthreads();   // cf / h2 are rewritten by the next channel
      }
#pragma unroll
      for (int k = 0; k < C::PPT; ++k) {
        gix[k] = fmaf(gx[k], dXx[c][k], gix[k]);
        giy[k] = fmaf(gx[k], dXy[c][k], giy[k]);
      }
    }
    float dPacc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) dPacc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < C::PPT; ++k) {
      if (!pin[k]) continue;
      float gx_ = sp[k].in_x ? gix[k] : 0.f, gy_ = sp[k].in_y ? giy[k] : 0.f;
      if (gx_ == 0.f && gy_ == 0.f) continue;
      const int v = v0 + prow[k], u = u0 + pcol[k];
      const float iz = 1.f / sp[k].z;
      const float g0 = gx_ * iz, g1 = gy_ * iz, g2 = -(gx_ * sp[k].pu + gy_ * sp[k].pv) * iz;
      const float r0 = cam.iK[0] * u + cam.iK[1] * v + cam.iK[2];
      const float r1 = cam.iK[3] * u + cam.iK[4] * v + cam.iK[5];
      const float r2 = cam.iK[6] * u + cam.iK[7] * v + cam.iK[8];
      gd[k] += g0 * (cam.P[0] * r0 + cam.P[1] * r1 + cam.P[2] * r2) + g1 * (cam.P[4] * r0 + cam.P[5] * r1 + cam.P[6] * r2) +
               g2 * (cam.P[8] * r0 + cam.P[9] * r1 + cam.P[10] * r2);
      dPacc[0] += g0 * sp[k].X[0]; dPacc[1] += g0 * sp[k].X[1]; dPacc[2] += g0 * sp[k].X[2]; dPacc[3] += g0;
      dPacc[4] += g1 * sp[k].X[0]; dPacc[5] += g1 * sp[k].X[1]; dPacc[6] += g1 * sp[k].X[2]; dPacc[7] += g1;
      dPacc[8] += g2 * sp[k].X[0]; dPacc[9] += g2 * sp[k].X[1]; dPacc[10] += g2 * sp[k].X[2]; dPacc[11] += g2;
    }
    if (threadIdx.x < 12) dPs[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float t = warp_sum(dPacc[i]);
      if ((threadIdx.x & 31) == 0 && t != 0.f) atomicAdd(&dPs[i], t);
    }
    __syncthreads();
    if (threadIdx.x < 12 && dPs[threadIdx.x] != 0.f)
      atomicAdd(p.dP + ((size_t)b * S + s) * 12 + threadIdx.x, dPs[threadIdx.x]);
  }

  // adjoint of the bilinear upsampling (same scheme as photo_bwd_kernel)
  const int vend = min(v0 + TH, H) - 1, uend = min(u0 + TW, W) - 1;
  const int i_lo = up_tap(v0, sy, h).i0, i_hi = up_tap(vend, sy, h).i1;
  const int j_lo = up_tap(u0, sx, w).i0, j_hi = up_tap(uend, sx, w).i1;
  const int nh = i_hi - i_lo + 1, nw = j_hi - j_lo + 1;
  float* acc = h2;
  __syncthreads();
  for (int idx = threadIdx.x; idx < nh * nw; idx += NT) acc[idx] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < C::PPT; ++k) {
    if (!pin[k] || gd[k] == 0.f) continue;
    const UpTap ty = up_tap(v0 + prow[k], sy, h), tx = up_tap(u0 + pcol[k], sx, w);
    const int a0 = (ty.i0 - i_lo) * nw, a1 = (ty.i1 - i_lo) * nw, c0 = tx.i0 - j_lo, c1 = tx.i1 - j_lo;
    atomicAdd(acc + a0 + c0, gd[k] * ty.l0 * tx.l0);
    atomicAdd(acc + a0 + c1, gd[k] * ty.l0 * tx.l1);
    atomicAdd(acc + a1 + c0, gd[k] * ty.l1 * tx.l0);
    atomicAdd(acc + a1 + c1, gd[k] * ty.l1 * tx.l1);
  }
  __syncthreads();
  float* out = p.d_depth_lr + (size_t)b * h * w;
  for (int idx = threadIdx.x; idx < nh * nw; idx += NT) {
    const float g = acc[idx];
    if (g != 0.f) {
      const int i = idx / nw, j = idx - i * nw;
      atomicAdd(out + (size_t)(i_lo + i) * w + (j_lo + j), g);
    }
  }
}

// dT[b,s] = K[b][:3,:]^T * dP[b,s]   (P = (K T)[:3,:]  =>  dL/dT = K[:3,:]^T dL/dP)
__global__ void dT_from_dP_kernel(const float* __restrict__ K, const float* __restrict__ dP, int B, int S,
                                  float* __restrict__ dT) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S * 16) return;
  const int e = idx & 15, bs = idx >> 4, b = bs / S;
  const int i = e >> 2, j = e & 3;  // dT[i][j] = sum_k K[k][i] dP[k][j], k<3
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) acc += K[b * 16 + k * 4 + i] * dP[(size_t)bs * 12 + k * 4 + j];
  dT[idx] = acc;
}

// ------------------------------------------------------------------------------------------------
// smoothness (layers.py:267-280 with the mean normalisation of trainer.py:535-536 factored out)
// ------------------------------------------------------------------------------------------------
constexpr int kSmoothBlocksPerSample = 64;

__device__ __forceinline__ float color_absdiff_mean(const float* __restrict__ col, size_t plane, int o0, int o1) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s += fabsf(__ldg(col + c * plane + o0) - __ldg(col + c * plane + o1));
  return s / 3.f;
}

__global__ void smooth_fwd_kernel(const float* __restrict__ disp_lr, const float* __restrict__ color, int h, int w,
                                  int Hc, int Wc, float* __restrict__ partial /*[B][blocks][3]*/) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float sy = (float)h / (float)Hc, sx = (float)w / (float)Wc;
  const float* lr = disp_lr + (size_t)b * h * w;
  const size_t plane = (size_t)Hc * Wc;
  const float* col = color + (size_t)b * 3 * plane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < Hc * Wc; idx += gridDim.x * blockDim.x) {
    const int v = idx / Wc, u = idx - v * Wc;
    const float d = upsample_at(lr, h, w, v, u, sy, sx);
    s2 += d;
    if (u + 1 < Wc) {
      const float dr = upsample_at(lr, h, w, v, u + 1, sy, sx);
      s0 += fabsf(d - dr) * expf(-color_absdiff_mean(col, plane, idx, idx + 1));
    }
    if (v + 1 < Hc) {
      const float dd = upsample_at(lr, h, w, v + 1, u, sy, sx);
      s1 += fabsf(d - dd) * expf(-color_absdiff_mean(col, plane, idx, idx + Wc));
    }
  }
  const float t0 = block_sum(s0, red);
  const float t1 = block_sum(s1, red);
  const float t2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    float* o = partial + ((size_t)b * gridDim.x + blockIdx.x) * 3;
    o[0] = t0; o[1] = t1; o[2] = t2;
  }
}

__device__ __forceinline__ void scatter_up(float* __restrict__ out, int h, int w, int v, int u, float sy, float sx,
                                           float g) {
  const UpTap ty = up_tap(v, sy, h), tx = up_tap(u, sx, w);
  atomicAdd(out + ty.i0 * w + tx.i0, g * ty.l0 * tx.l0);
  atomicAdd(out + ty.i0 * w + tx.i1, g * ty.l0 * tx.l1);
  atomicAdd(out + ty.i1 * w + tx.i0, g * ty.l1 * tx.l0);
  atomicAdd(out + ty.i1 * w + tx.i1, g * ty.l1 * tx.l1);
}

// g_sums[b] = upstream gradients of {sum_x, sum_y, sum_d}
__global__ void smooth_bwd_kernel(const float* __restrict__ disp_lr, const float* __restrict__ color, int h, int w,
                                  int Hc, int Wc, const float* __restrict__ g_sums, float* __restrict__ d_lr) {
  const int b = blockIdx.y;
  const float sy = (float)h / (float)Hc, sx = (float)w / (float)Wc;
  const float* lr = disp_lr + (size_t)b * h * w;
  float* out = d_lr + (size_t)b * h * w;
  const size_t plane = (size_t)Hc * Wc;
  const float* col = color + (size_t)b * 3 * plane;
  const float gxs = g_sums[b * 3 + 0], gys = g_sums[b * 3 + 1], gds = g_sums[b * 3 + 2];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < Hc * Wc; idx += gridDim.x * blockDim.x) {
    const int v = idx / Wc, u = idx - v * Wc;
    const float d = upsample_at(lr, h, w, v, u, sy, sx);
    float g = gds;  // gradient wrt the upsampled value at (v,u)
    if (u + 1 < Wc) {
      const float diff = d - upsample_at(lr, h, w, v, u + 1, sy, sx);
      const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      g += gxs * sg * expf(-color_absdiff_mean(col, plane, idx, idx + 1));
    }
    if (u > 0) {
      const float diff = upsample_at(lr, h, w, v, u - 1, sy, sx) - d;
      const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      g -= gxs * sg * expf(-color_absdiff_mean(col, plane, idx - 1, idx));
    }
    if (v + 1 < Hc) {
      const float diff = d - upsample_at(lr, h, w, v + 1, u, sy, sx);
      const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      g += gys * sg * expf(-color_absdiff_mean(col, plane, idx, idx + Wc));
    }
    if (v > 0) {
      const float diff = upsample_at(lr, h, w, v - 1, u, sy, sx) - d;
      const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      g -= gys * sg * expf(-color_absdiff_mean(col, plane, idx - Wc, idx));
    }
    if (g != 0.f) scatter_up(out, h, w, v, u, sy, sx, g);
  }
}

}  // namespace sqlx

using namespace sqlx;

namespace {
constexpr int kTH = 16, kTW = 32, kNT = 256;

template <int R>
int launch_photo_bwd(const PhotoBwdParams& p, cudaStream_t st) {
  using C = BwdCfg<R, kTH, kTW, kNT>;
  auto kern = photo_bwd_kernel<R, kTH, kTW, kNT>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
    configured = true;
  }
  dim3 grid(ceil_div(p.d.W, kTW), ceil_div(p.d.H, kTH), p.d.B);
  ProfScope prof("photo_bwd_kernel", st);
  kern<<<grid, kNT, C::smem_bytes, st>>>(p);
  return check_launch("photo_bwd_kernel");
}

template <int R>
int launch_photo_bwd2(const PhotoBwdParams& p, const float* coef, cudaStream_t st) {
  using C = Bwd2Cfg<R, kTH, kTW, kNT>;
  static const int minb = getenv("SQLX_BWD_MINB") ? atoi(getenv("SQLX_BWD_MINB")) : 4;  // resident CTAs per SM the register budget is capped for (tuning knob)
  dim3 grid(ceil_div(p.d.W, kTW), ceil_div(p.d.H, kTH), p.d.B);
  ProfScope prof("photo_bwd_kernel", st);
  if (minb >= 4) photo_bwd2_kernel<R, kTH, kTW, kNT, 4><<<grid, kNT, C::smem_bytes, st>>>(p, coef);
  else if (minb == 3) photo_bwd2_kernel<R, kTH, kTW, kNT, 3><<<grid, kNT, C::smem_bytes, st>>>(p, coef);
  else photo_bwd2_kernel<R, kTH, kTW, kNT, 2><<<grid, kNT, C::smem_bytes, st>>>(p, coef);
  return check_launch("photo_bwd2_kernel");
}

__global__ void reduce_rows_kernel(const float* __restrict__ partial, int per_b, int ncol, float* __restrict__ out) {
  __shared__ double sh[256];
  const int b = blockIdx.x, col = blockIdx.y;
  double acc = 0.0;
  for (int i = threadIdx.x; i < per_b; i += blockDim.x) acc += (double)partial[((size_t)b * per_b + i) * ncol + col];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[b * ncol + col] = (float)sh[0];
}
}  // namespace

extern "C" int sqlx_photo_bwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                              const float* const* sources, const float* K, const float* inv_K, const float* T,
                              const uint8_t* argmin, const float* ssim_coef, const float* g_loss, float scale,
                              float* d_depth_lr, float* d_T, void* workspace, size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(desc != nullptr, "desc is NULL");
  SQLX_REQUIRE(desc->B > 0 && desc->H > 0 && desc->W > 0 && desc->h > 0 && desc->w > 0, "non-positive shape");
  SQLX_REQUIRE(desc->S >= 1 && desc->S <= SQLX_MAX_SOURCES, "S=%d outside 1..%d", desc->S, SQLX_MAX_SOURCES);
  SQLX_REQUIRE(desc->h <= desc->H && desc->w <= desc->W, "depth map larger than the image is not supported");
  SQLX_REQUIRE(depth_lr && target && sources && K && inv_K && T && argmin && g_loss && d_depth_lr && d_T,
               "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sqlx_photo_workspace_bytes(desc), "workspace too small");
  const int r = (desc->flags & SQLX_NO_SSIM) ? 0 : desc->ssim_radius;
  SQLX_REQUIRE(r == 0 || r == 1 || r == 3, "ssim_radius must be 1 or 3 (got %d)", desc->ssim_radius);
  SQLX_REQUIRE(desc->H > 2 * r && desc->W > 2 * r, "image smaller than the SSIM window");
  PhotoBwdParams p;
  p.d = *desc;
  p.depth_lr = depth_lr; p.target = target;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s) p.src[s] = s < desc->S ? sources[s] : nullptr;
  for (int s = 0; s < desc->S; ++s) SQLX_REQUIRE(p.src[s], "source %d is NULL", s);
  p.K = K; p.invK = inv_K; p.T = T; p.argmin = argmin; p.g_loss = g_loss; p.scale = scale;
  p.d_depth_lr = d_depth_lr;
  // dP accumulators live after the forward partial sums in the workspace
  const size_t ctas = (size_t)ceil_div(desc->W, kTW) * ceil_div(desc->H, kTH) * desc->B;
  p.dP = reinterpret_cast<float*>(workspace) + ctas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(p.dP, 0, sizeof(float) * (size_t)desc->B * desc->S * 12, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(dP)");
  int e;
  if (ssim_coef || r == 0)   // saved-coefficient path (no recompute); r == 0 (--no_ssim) needs no coefficients at all
    e = r == 3 ? launch_photo_bwd2<3>(p, ssim_coef, st) : (r == 1 ? launch_photo_bwd2<1>(p, ssim_coef, st)
                                                                  : launch_photo_bwd2<0>(p, ssim_coef, st));
  else
    e = r == 3 ? launch_photo_bwd<3>(p, st) : (r == 1 ? launch_photo_bwd<1>(p, st) : launch_photo_bwd<0>(p, st));
  if (e) return e;
  const int n = desc->B * desc->S * 16;
  dT_from_dP_kernel<<<ceil_div(n, 128), 128, 0, st>>>(K, p.dP, desc->B, desc->S, d_T);
  return check_launch("dT_from_dP_kernel");
}

extern "C" size_t sqlx_smooth_workspace_bytes(int B, int Hc, int Wc) {
  (void)Hc; (void)Wc;
  return sizeof(float) * (size_t)B * kSmoothBlocksPerSample * 3;
}

extern "C" int sqlx_smooth_fwd(const float* disp_lr, const float* color, int B, int h, int w, int Hc, int Wc,
                               float* sums, void* workspace, size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(disp_lr && color && sums && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && Hc > 1 && Wc > 1 && h > 0 && w > 0 && h <= Hc && w <= Wc, "bad shape");
  SQLX_REQUIRE(workspace_bytes >= sqlx_smooth_workspace_bytes(B, Hc, Wc), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  ProfScope prof("smooth_fwd_kernel", st);
  smooth_fwd_kernel<<<dim3(kSmoothBlocksPerSample, B), 256, 0, st>>>(disp_lr, color, h, w, Hc, Wc, partial);
  if (int e = check_launch("smooth_fwd_kernel")) return e;
  reduce_rows_kernel<<<dim3(B, 3), 256, 0, st>>>(partial, kSmoothBlocksPerSample, 3, sums);
  return check_launch("reduce_rows_kernel");
}

extern "C" int sqlx_smooth_bwd(const float* disp_lr, const float* color, int B, int h, int w, int Hc, int Wc,
                               const float* g_sums, float* d_disp_lr, void* stream) {
  SQLX_REQUIRE(disp_lr && color && g_sums && d_disp_lr, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && Hc > 1 && Wc > 1 && h > 0 && w > 0 && h <= Hc && w <= Wc, "bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope prof("smooth_bwd_kernel", st);
  smooth_bwd_kernel<<<dim3(kSmoothBlocksPerSample, B), 256, 0, st>>>(disp_lr, color, h, w, Hc, Wc, g_sums, d_disp_lr);
  return check_launch("smooth_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// SSIM backward (module-level drop-in for layers.SSIM; autograd of layers.py:31-46)
// ------------------------------------------------------------------------------------------------
namespace sqlx {

// gx[b,c,v,u] = d/dx sum(g_out * SSIM(x,y)).  SSIM is symmetric, so the caller obtains the gradient wrt y by
// swapping x and y.  One CTA = one (tile, channel-plane); recomputes the box sums on a 2R halo.
template <int R, int TH, int TW, int NT>
__global__ void __launch_bounds__(NT) ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* __restrict__ g_out, int H, int W,
                                                      float* __restrict__ gx) {
  using C = BwdCfg<R, TH, TW, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);
  float* ys = xs + C::PLANE2;
  float* hb = ys + C::PLANE2;          // 5 * HB
  float* cf = hb + 5 * C::HB;          // 3 * CF
  const int pl = blockIdx.z;           // b*C + c
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  stage_plane<C::PH2, C::PW2, C::LD2>(x + pl * plane, H, W, v0 - 2 * R, u0 - 2 * R, xs);
  stage_plane<C::PH2, C::PW2, C::LD2>(y + pl * plane, H, W, v0 - 2 * R, u0 - 2 * R, ys);
  __syncthreads();
  hpass5<R, C::PH2, C::PW1, C::LD2, C::PW1, true>(xs, ys, hb, hb + C::HB, hb + 2 * C::HB, hb + 3 * C::HB,
                                                  hb + 4 * C::HB);
  __syncthreads();
  for (int idx = threadIdx.x; idx < C::CF; idx += NT) {
    const int lr = idx / C::PW1, lc = idx - lr * C::PW1;
    const int v = v0 - R + lr, u = u0 - R + lc;
    float a = 0.f, bb = 0.f, cc = 0.f;
    if (v >= 0 && v < H && u >= 0 && u < W) {
      const float g = __ldg(g_out + pl * plane + (size_t)v * W + u);
      const float Sx = vsum<R, C::PW1>(hb, lr, lc);
      const float Sxx = vsum<R, C::PW1>(hb + C::HB, lr, lc);
      const float Sxy = vsum<R, C::PW1>(hb + 2 * C::HB, lr, lc);
      const float Sy = vsum<R, C::PW1>(hb + 3 * C::HB, lr, lc);
      const float Syy = vsum<R, C::PW1>(hb + 4 * C::HB, lr, lc);
      const SsimGrad sg = ssim_grad(make_stats<R>(Sx, Sy, Sxx, Syy, Sxy));
      a = g * sg.dmx; bb = g * sg.dexx; cc = g * sg.dexy;
    }
    cf[idx] = a; cf[C::CF + idx] = bb; cf[2 * C::CF + idx] = cc;
  }
  __syncthreads();
  float* h2 = hb;
  for (int idx = threadIdx.x; idx < C::PH1 * TW; idx += NT) {
    const int lr = idx / TW, pc = idx - lr * TW;
    const int u = u0 + pc;
    const float* ca = cf + lr * C::PW1 + pc;
    float sa = 0.f, sb = 0.f, sc = 0.f;
    if (u < W) {
#pragma unroll
      for (int k = 0; k <= 2 * R; ++k) {
        const int qu = u - R + k;
        if (qu < 0 || qu >= W) continue;
        const float m = reflect_mult<R>(u, qu, W);
        sa += m * ca[k]; sb += m * ca[C::CF + k]; sc += m * ca[2 * C::CF + k];
      }
    }
    h2[idx] = sa; h2[C::PH1 * TW + idx] = sb; h2[2 * C::PH1 * TW + idx] = sc;
  }
  __syncthreads();
  for (int pix = threadIdx.x; pix < TH * TW; pix += NT) {
    const int pr = pix / TW, pc = pix - pr * TW;
    const int v = v0 + pr, u = u0 + pc;
    if (v >= H || u >= W) continue;
    const float* ha = h2 + pr * TW + pc;
    float sa = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll
    for (int j = 0; j <= 2 * R; ++j) {
      const int qv = v - R + j;
      if (qv < 0 || qv >= H) continue;
      const float m = reflect_mult<R>(v, qv, H);
      sa += m * ha[j * TW]; sb += m * ha[C::PH1 * TW + j * TW]; sc += m * ha[2 * C::PH1 * TW + j * TW];
    }
    const int o = (pr + 2 * R) * C::LD2 + pc + 2 * R;
    gx[pl * plane + (size_t)v * W + u] = ia * (sa + 2.f * xs[o] * sb + ys[o] * sc);
  }
}

}  // namespace sqlx

namespace {
template <int R>
int launch_ssim_bwd(const float* x, const float* y, const float* g, int planes, int H, int W, float* gx,
                    cudaStream_t st) {
  using C = BwdCfg<R, kTH, kTW, kNT>;
  auto kern = ssim_bwd_kernel<R, kTH, kTW, kNT>;
  const size_t smem = sizeof(float) * (2 * C::PLANE2 + 5 * C::HB + 3 * C::CF);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  dim3 grid(ceil_div(W, kTW), ceil_div(H, kTH), planes);
  kern<<<grid, kNT, smem, st>>>(x, y, g, H, W, gx);
  return check_launch("ssim_bwd_kernel");
}
}  // namespace

extern "C" int sqlx_ssim_bwd(const float* x, const float* y, const float* g_out, int B, int C, int H, int W,
                             int ssim_radius, float* gx, float* gy, void* stream) {
  SQLX_REQUIRE(x && y && g_out && (gx || gy), "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "non-positive shape");
  SQLX_REQUIRE(ssim_radius == 1 || ssim_radius == 3, "ssim_radius must be 1 or 3");
  SQLX_REQUIRE(H > 2 * ssim_radius && W > 2 * ssim_radius, "image smaller than the SSIM window");
  SQLX_REQUIRE((long long)B * C <= 65535, "B*C too large for one launch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int pass = 0; pass < 2; ++pass) {
    float* out = pass == 0 ? gx : gy;
    if (!out) continue;
    const float* a = pass == 0 ? x : y;   // SSIM(x,y) == SSIM(y,x): d/dy is d/dx with the roles swapped
    const float* b = pass == 0 ? y : x;
    int e = ssim_radius == 3 ? launch_ssim_bwd<3>(a, b, g_out, B * C, H, W, out, st)
                             : launch_ssim_bwd<1>(a, b, g_out, B * C, H, W, out, st);
    if (e) return e;
  }
  return SQLX_OK;
}
