// Fused photometric kernels (forward and saved-coefficient backward), third generation.
//
//   photo_fwd3_kernel   upsample -> backproject -> project -> border-clamped bilinear warp of S sources ->
//                       7x7 SSIM + L1 vs target -> running min with the identity (+noise) losses -> partial sums,
//                       arg-min and the SSIM derivative coefficients of every (source, channel)
//   photo_bwd3_kernel   arg-min masked adjoint box filter of the saved coefficients -> bilinear-sample adjoint ->
//                       projection / back-projection adjoint -> bilinear-upsample adjoint (d depth_lr) and d P
//   pack_rgba_kernel    [B,3,H,W] planar frame -> [B,H,W,4] pixel-interleaved frame: one 16-byte load per bilinear
//                       tap instead of three 4-byte loads from three planes
//
// Reference lines: layers.py:31-46 (SSIM), 210-215 (backproject), 247-258 (project); trainer.py:395-396 (upsample),
// 431-435 (grid_sample border / align_corners), 444-451 (0.85 SSIM + 0.15 L1), 516-532 (noise, min, mean).
//
// Instruction-count driven design (the previous generation issued 2430 thread instructions per pixel and launch and
// was issue-bound at 7 % of the HBM roofline):
//   * 32-wide tiles with the R halo staged once; region loops carry (row, col) incrementally (no div/mod) and
//     take the reflect path only on tiles that touch the frame border
//   * projection with one reciprocal (MUFU.RCP) instead of six IEEE divisions; the normalise / un-normalise round
//     trip of Project3D + grid_sample is the identity and is skipped (the module-level Project3D drop-in and
//     sqlx_warp_fwd keep the literal arithmetic)
//   * bilinear taps clamped to (W-2, H-2) so the four taps always sit at offsets {0, 1, W, W+1}: identical values
//     (the weight of the clamped-away tap is exactly 0 in ATen), 4 x LDG.128 per warped pixel
//   * separable box sums: register-blocked horizontal pass (4 outputs / item), PPT vertically adjacent outputs per
//     thread in the vertical pass, ping-pong horizontal buffers -> 4 block barriers per source
//   * one reciprocal per SSIM value, shared with the derivative coefficients
#include "photo_tile.cuh"

#include <stdlib.h>
#include <string.h>

namespace sqlx {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__global__ void pack_rgba_kernel(const float* __restrict__ img, int B, int plane, float4* __restrict__ out) {
  const size_t total = (size_t)B * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, o = i - b * plane;
    const float* p = img + b * 3 * plane + o;
    out[i] = make_float4(__ldg(p), __ldg(p + plane), __ldg(p + 2 * (size_t)plane), 0.f);
  }
}

// 1/x as a single MUFU.RCP (rcp.approx.ftz): every caller's argument is far from the denormal range
// (SSIM denominators >= C1*C2 = 9e-8; camera depths), so the range fix-up of __fdividef is dead weight
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// region loop: f(lr, lc) for every element of an RH x RW region, thread-strided, (lr, lc) carried incrementally
template <int RH, int RW, int NT, class F>
__device__ __forceinline__ void for_region(F&& f) {
  int lr = threadIdx.x / RW, lc = threadIdx.x - lr * RW;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < RH * RW; idx += NT) {
    f(lr, lc, idx);
    lc += NT % RW;
    lr += NT / RW;
    if (lc >= RW) { lc -= RW; ++lr; }
  }
}

// advance (lr, lc) by NT flattened positions of an RW-wide region
template <int RW, int NT>
__device__ __forceinline__ void region_advance(int& lr, int& lc) {
  lc += NT % RW;
  lr += NT / RW;
  if (lc >= RW) { lc -= RW; ++lr; }
}

// Two region elements per trip so that the global loads of both are in flight together:
//   pre(j, lr, lc, live) issues the loads of element j into caller-owned registers, post(j, lr, lc) consumes them.
template <int RH, int RW, int NT, class Pre, class Post>
__device__ __forceinline__ void for_region2(Pre&& pre, Post&& post) {
  int lr0 = threadIdx.x / RW, lc0 = threadIdx.x - lr0 * RW;
  for (int idx = threadIdx.x; idx < RH * RW; idx += 2 * NT) {
    int lr1 = lr0, lc1 = lc0;
    region_advance<RW, NT>(lr1, lc1);
    const bool has1 = idx + NT < RH * RW;
    pre(0, lr0, lc0, true);
    pre(1, has1 ? lr1 : lr0, has1 ? lc1 : lc0, has1);
    post(0, lr0, lc0);
    if (has1) post(1, lr1, lc1);
    lr0 = lr1; lc0 = lc1;
    region_advance<RW, NT>(lr0, lc0);
  }
}

// Per-CTA tables of the region's rows / columns: reflected frame coordinate and the bilinear-upsample taps of the
// low-resolution depth map (align_corners=False).  x = first tap offset, y = second tap offset, z = bits of the
// second tap's weight, w = frame coordinate.
template <int N>
__device__ __forceinline__ void fill_axis_table(int4* tab, int first, int size, int lr_size, float scale, int lr_stride,
                                                int tid0) {
  const int i = (int)threadIdx.x - tid0;
  if (i >= 0 && i < N) {
    const int c = clamp_reflect(first + i, size);
    const UpTap t = up_tap(c, scale, lr_size);
    tab[i] = make_int4(t.i0 * lr_stride, t.i1 * lr_stride, __float_as_int(t.l1), c);
  }
}

struct Proj {
  float pu, pv;      // projected pixel coordinates (= the un-normalised grid_sample coordinates)
  float rz;          // 1 / (z + eps)
  float X0, X1, X2;  // camera point
};
__device__ __forceinline__ Proj project_fast(const Camera& cam, float u, float v, float d, float eps) {
  Proj s;
  const float r0 = fmaf(cam.iK[0], u, fmaf(cam.iK[1], v, cam.iK[2]));
  const float r1 = fmaf(cam.iK[3], u, fmaf(cam.iK[4], v, cam.iK[5]));
  const float r2 = fmaf(cam.iK[6], u, fmaf(cam.iK[7], v, cam.iK[8]));
  s.X0 = d * r0; s.X1 = d * r1; s.X2 = d * r2;
  const float c0 = fmaf(cam.P[0], s.X0, fmaf(cam.P[1], s.X1, fmaf(cam.P[2], s.X2, cam.P[3])));
  const float c1 = fmaf(cam.P[4], s.X0, fmaf(cam.P[5], s.X1, fmaf(cam.P[6], s.X2, cam.P[7])));
  const float c2 = fmaf(cam.P[8], s.X0, fmaf(cam.P[9], s.X1, fmaf(cam.P[10], s.X2, cam.P[11])));
  s.rz = rcp_fast(c2 + eps);
  s.pu = c0 * s.rz;
  s.pv = c1 * s.rz;
  return s;
}

// Forward-only form of the same projection with back-projection and projection composed:
//   (K T)[:3,:] (d * inv_K (u,v,1), 1) = d * (M (u,v,1)) + t,  M = P[:, :3] inv_K[:3,:3],  t = P[:, 3]
// 12 constants and 9 FMAs per pixel instead of 21 and 18 (the backward needs the camera point and keeps project_fast).
struct CamM {
  float M[9], t[3];
};
__device__ __forceinline__ CamM compose_camera(const Camera& cam) {
  CamM c;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.M[i * 3 + j] = cam.P[i * 4 + 0] * cam.iK[0 * 3 + j] + cam.P[i * 4 + 1] * cam.iK[1 * 3 + j] +
                       cam.P[i * 4 + 2] * cam.iK[2 * 3 + j];
    c.t[i] = cam.P[i * 4 + 3];
  }
  return c;
}
__device__ __forceinline__ void project_composed(const CamM& c, float u, float v, float d, float eps, float& pu,
                                                 float& pv) {
  const float m0 = fmaf(c.M[0], u, fmaf(c.M[1], v, c.M[2]));
  const float m1 = fmaf(c.M[3], u, fmaf(c.M[4], v, c.M[5]));
  const float m2 = fmaf(c.M[6], u, fmaf(c.M[7], v, c.M[8]));
  const float rz = rcp_fast(fmaf(d, m2, c.t[2]) + eps);
  pu = fmaf(d, m0, c.t[0]) * rz;
  pv = fmaf(d, m1, c.t[1]) * rz;
}

struct Taps4 {
  int o;         // y0 * W + x0 with x0 <= W-2, y0 <= H-2
  float fx, fy;  // may equal 1 on the last column / row
};
__device__ __forceinline__ Taps4 make_taps4(float pu, float pv, int H, int W) {
  const float ix = fminf((float)(W - 1), fmaxf(pu, 0.f));   // padding_mode="border" (NaN -> 0 as fmaxf does)
  const float iy = fminf((float)(H - 1), fmaxf(pv, 0.f));
  const int x0 = min((int)ix, W - 2), y0 = min((int)iy, H - 2);
  Taps4 t;
  t.fx = ix - (float)x0;
  t.fy = iy - (float)y0;
  t.o = y0 * W + x0;
  return t;
}

// Horizontal (2R+1)-tap sums of ONE plane, 4 outputs per item (adjoint pass of the backward kernel).
template <int R, int ROWS, int OUTW, int LD, int OLD>
__device__ __forceinline__ void hsum_blocked(const float* __restrict__ X, float* __restrict__ out) {
  static_assert(OUTW % 4 == 0 && LD % 2 == 0 && OLD % 4 == 0, "blocked horizontal pass alignment");
  constexpr int NIN = 4 + 2 * R;
  constexpr int ITEMS = ROWS * (OUTW / 4);
  for (int idx = threadIdx.x; idx < ITEMS; idx += blockDim.x) {
    const int r = idx / (OUTW / 4), c = (idx - r * (OUTW / 4)) * 4;
    float x[NIN];
    const float2* xp = reinterpret_cast<const float2*>(X + r * LD + c);
#pragma unroll
    for (int k = 0; k < NIN / 2; ++k) { const float2 v = xp[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
    float o[4];
    float a = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) a += x[k];
    o[0] = a;
#pragma unroll
    for (int j = 1; j < 4; ++j) { a += x[j + 2 * R] - x[j - 1]; o[j] = a; }
    *reinterpret_cast<float4*>(out + r * OLD + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (instead of 5 per value): after the call
// lanes 2i and 2i+1 both hold the warp total of v[i] in v[0].
__device__ __forceinline__ void warp_reduce16(float (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int w = 16, n = 8; n >= 1; w >>= 1, n >>= 1) {
    const bool upper = lane & w;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = upper ? v[i] : v[i + n];
      const float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct PhotoFwdParams {
  sqlx_photo_desc d;
  const float* depth_up;   // optional [B,H,W]: upsampled depth (multiscale.cu materialises it once per scale)
  const float* depth_lr;
  const float* target;
  const float4* src[SQLX_MAX_SOURCES];   // pixel-interleaved [B,H,W,4]
  const float* K;
  const float* invK;
  const float* T;
  const float* identity;
  const float* noise;
  float* partial;      // [gridDim.z*gridDim.y*gridDim.x]
  uint8_t* argmin;
  float* coef;         // optional [B][S][3 ch][3][H][W]: d SSIM / d(mean_x, E[x^2], E[xy])
  // indoor variant (OCC kernels only; trainer_indoor.py:583-587, 636-651)
  const float* ref[SQLX_MAX_SOURCES];    // depth of each source frame, planar [B,H,W]
  float* partial_reg;                    // per-CTA partial sums of diff_depth * valid_mask
  // all loss scales in one CTA pass (MS kernels only): the target tile, its box statistics and the identity losses are
  // staged once and shared by the scales; scale i reads ms_depth_up[i] / ms_noise[i] / T + i * ms_T_stride and writes
  // ms_argmin[i] / partial + i * ms_partial_stride / coef + i * ms_coef_stride
  int ns;
  const float* ms_depth_up[SQLX_MAX_SCALES];
  const float* ms_noise[SQLX_MAX_SCALES];
  uint8_t* ms_argmin[SQLX_MAX_SCALES];
  size_t ms_T_stride, ms_partial_stride, ms_coef_stride;   // in floats
};

// Depth-consistency terms of the indoor loss at one pixel (trainer_indoor.py:636-648): pd = the source frame's depth
// sampled at the projected position, diff = |d - pd| / (d + pd), weight = (1 - sqrt(1 - (diff - 1)^2)) * valid.
struct OccTerm {
  float pd, diff, weight;
  bool valid;
};
__device__ __forceinline__ OccTerm occ_term(float d, float pd, float r, float g, float b) {
  OccTerm o;
  o.pd = pd;
  o.valid = (fabsf(r) + fabsf(g) + fabsf(b)) / 3.f > 1e-3f;
  o.diff = fabsf(d - pd) / (d + pd);
  const float e = o.diff - 1.f;
  o.weight = o.valid ? 1.f - sqrtf(1.f - e * e) : 0.f;
  return o;
}

template <int R, int TH, int TW, int NT, bool MERGED = false>
struct Fwd3Cfg {
  static constexpr int PH = TH + 2 * R, PW = TW + 2 * R;
  static constexpr int LD = ((PW + 3) & ~3) + 2;   // even, = 10 mod 32 for a 32-wide tile: conflict-free 8-byte rows
  static constexpr int PLANE = PH * LD;
  static constexpr int HB = PH * TW;
  static constexpr int PPT = (TH * TW) / NT;
  static constexpr int TS = TH * TW;              // one plane of per-pixel target statistics
  static constexpr int NHB = MERGED ? 9 : 6;       // horizontal-sum planes: all channels at once, or ping-pong
  static constexpr size_t smem_bytes = sizeof(float) * (7 * PLANE + NHB * HB + 6 * TS + 32) +
                                       sizeof(Camera) * SQLX_MAX_SOURCES + sizeof(int4) * (PH + PW);
  static_assert((TH * TW) % NT == 0 && NT % TW == 0 && NT >= PH + PW, "tile / block shape");
};

template <int R, int TH, int TW, int NT, int MINB, bool MERGED, bool OCC = false, bool MS = false>
__global__ void __launch_bounds__(NT, MINB) photo_fwd3_kernel(const __grid_constant__ PhotoFwdParams p) {
  using C = Fwd3Cfg<R, TH, TW, NT, MERGED>;
  constexpr int PPT = C::PPT;
  constexpr int RR = R > 0 ? R : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* dpl = reinterpret_cast<float*>(smem_raw);
  float* tg = dpl + C::PLANE;           // 3 planes
  float* wp = tg + 3 * C::PLANE;        // 3 planes
  float* hbA = wp + 3 * C::PLANE;       // 3 planes of HB
  float* hbB = hbA + 3 * C::HB;         // 3 planes of HB
  float* tstat = hbA + C::NHB * C::HB; // [3 ch][mean, variance + C2][TH*TW]: target statistics of the owned pixels
  float* red = tstat + 6 * C::TS;
  int4* rowt = reinterpret_cast<int4*>(red + 32);     // [PH]
  int4* colt = rowt + C::PH;                           // [PW]
  Camera* cams = reinterpret_cast<Camera*>(colt + C::PW);

  const int H = p.d.H, W = p.d.W, S = p.d.S;
  const int b = blockIdx.z;
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  const bool automask = p.d.flags & SQLX_AUTOMASK;
  const bool avg = p.d.flags & SQLX_AVG_REPROJ;
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));

  fill_axis_table<C::PH>(rowt, v0 - R, H, p.d.h, (float)p.d.h / (float)H, p.d.w, 0);
  fill_axis_table<C::PW>(colt, u0 - R, W, p.d.w, (float)p.d.w / (float)W, 1, C::PH);

  // owned pixels: PPT vertically adjacent rows of one column
  const int pcol = threadIdx.x % TW;
  const int prow0 = (threadIdx.x / TW) * PPT;
  const bool col_in = u0 + pcol < W;
  const size_t pix0 = (size_t)(v0 + prow0) * W + (u0 + pcol);   // offset of the first owned pixel in a plane
  const int n_ident = automask ? (avg ? 1 : S) : 0;
  float* tsp = tstat + prow0 * TW + pcol;
  const size_t cta_index = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

  const int nsc = MS ? p.ns : 1;
  for (int sc = 0; sc < nsc; ++sc) {
  const bool first = !MS || sc == 0;
  const float* T_sc = MS ? p.T + (size_t)sc * p.ms_T_stride : p.T;
  const float* noise_sc = MS ? p.ms_noise[sc] : p.noise;
  uint8_t* argmin_sc = MS ? p.ms_argmin[sc] : p.argmin;
  float* coef_sc = (MS && p.coef) ? p.coef + (size_t)sc * p.ms_coef_stride : p.coef;
  if (threadIdx.x < S)
    load_camera(p.K + b * 16, p.invK + b * 16, T_sc + ((size_t)b * S + threadIdx.x) * 16, cams[threadIdx.x]);
  __syncthreads();

  {   // upsampled depth and (first scale only) the three target planes on the R halo, two elements per trip
    // (128-bit row-segment loads on interior tiles were measured SLOWER, 420 -> 440 us: the kernel sits at the 128-register
    // cap of 2 CTAs/SM and any extra live state spills, profiles/r02d_ab.log)
    const float* lr_map = p.depth_lr + (size_t)b * p.d.h * p.d.w;
    const float* dup = MS ? p.ms_depth_up[sc] : p.depth_up;
    const float* up_map = dup ? dup + (size_t)b * plane : nullptr;
    const float* tgb = p.target + (size_t)b * 3 * plane;
    {
      float dv[2][4], tv[2][3], wy[2], wx[2];
      for_region2<C::PH, C::PW, NT>(
          [&](int j, int lr, int lc, bool live) {
            const int4 rt = rowt[lr], ct = colt[lc];
            wy[j] = __int_as_float(rt.z); wx[j] = __int_as_float(ct.z);
            if (live) {
              if (up_map) {
                dv[j][0] = __ldg(up_map + rt.w * W + ct.w);
              } else {
                const float* r0 = lr_map + rt.x;
                const float* r1 = lr_map + rt.y;
                dv[j][0] = __ldg(r0 + ct.x); dv[j][1] = __ldg(r0 + ct.y);
                dv[j][2] = __ldg(r1 + ct.x); dv[j][3] = __ldg(r1 + ct.y);
              }
              if (first) {
                const float* tp = tgb + (size_t)(rt.w * W + ct.w);
                tv[j][0] = __ldg(tp); tv[j][1] = __ldg(tp + plane); tv[j][2] = __ldg(tp + 2 * plane);
              }
            }
          },
          [&](int j, int lr, int lc) {
            const int o = lr * C::LD + lc;
            // same expression as upsample_at (common.cuh): bit-identical depth in every kernel
            dpl[o] = up_map ? dv[j][0]
                            : (1.f - wy[j]) * ((1.f - wx[j]) * dv[j][0] + wx[j] * dv[j][1]) +
                                  wy[j] * ((1.f - wx[j]) * dv[j][2] + wx[j] * dv[j][3]);
            if (first) { tg[o] = tv[j][0]; tg[C::PLANE + o] = tv[j][1]; tg[2 * C::PLANE + o] = tv[j][2]; }
          });
    }
  }

  // identity (+noise) candidates: independent global loads issued before the staging barrier, consumed after it
  float best[PPT];
  int arg[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) { best[k] = INFINITY; arg[k] = 0; }
  float idv[SQLX_MAX_SOURCES][PPT], nzv[SQLX_MAX_SOURCES][PPT];
  if (automask) {
    const float* idp = p.identity + (size_t)b * S * plane + pix0;
    const float* nzp = noise_sc + (size_t)b * (avg ? 1 : S) * plane + pix0;
#pragma unroll
    for (int s = 0; s < SQLX_MAX_SOURCES; ++s) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const bool live = s < S && col_in && (v0 + prow0 + k < H);
        idv[s][k] = live ? __ldg(idp + (size_t)s * plane + k * W) : 0.f;
        nzv[s][k] = (live && (!avg || s == 0)) ? __ldg(nzp + (size_t)s * plane + k * W) : 0.f;
      }
    }
  }
  __syncthreads();
  if (automask) {
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (avg) {
        float m = 0.f;
#pragma unroll
        for (int s = 0; s < SQLX_MAX_SOURCES; ++s) m += idv[s][k];
        best[k] = m / (float)S + nzv[0][k] * p.d.noise_scale;
      } else {
#pragma unroll
        for (int s = 0; s < SQLX_MAX_SOURCES; ++s) {
          const float v = idv[s][k] + nzv[s][k] * p.d.noise_scale;
          if (s < S && v < best[k]) { best[k] = v; arg[k] = s; }
        }
      }
    }
  }

  // target statistics per channel (mean and variance + C2): written and read back by the owning thread only
  // (shared memory rather than 6*PPT registers that would stay live across the whole source loop)
  if (!first) {
    // the target statistics of the tile are already in shared memory
  } else if (R > 0 && MERGED) {
    // all channels in one pass: 6 horizontal planes, one barrier
#pragma unroll
    for (int c = 0; c < 3; ++c)
      hpass_blocked<RR, C::PH, TW, C::LD, TW, false>(nullptr, tg + c * C::PLANE, hbA + (2 * c) * C::HB,
                                                     hbA + (2 * c + 1) * C::HB, nullptr);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float Sy[PPT], Syy[PPT];
      vsum_multi<R, TW, PPT>(hbA + (2 * c) * C::HB, prow0, pcol, Sy);
      vsum_multi<R, TW, PPT>(hbA + (2 * c + 1) * C::HB, prow0, pcol, Syy);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const float m = Sy[k] * ia;
        tsp[(2 * c) * C::TS + k * TW] = m;
        tsp[(2 * c + 1) * C::TS + k * TW] = fmaf(-m, m, Syy[k] * ia) + kC2;
      }
    }
  } else if (R > 0) {
    hpass_blocked<RR, C::PH, TW, C::LD, TW, false>(nullptr, tg, hbA, hbA + C::HB, nullptr);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float* cur = (c & 1) ? hbB : hbA;
      float* nxt = (c & 1) ? hbA : hbB;
      if (c < 2) hpass_blocked<RR, C::PH, TW, C::LD, TW, false>(nullptr, tg + (c + 1) * C::PLANE, nxt, nxt + C::HB, nullptr);
      float Sy[PPT], Syy[PPT];
      vsum_multi<R, TW, PPT>(cur, prow0, pcol, Sy);
      vsum_multi<R, TW, PPT>(cur + C::HB, prow0, pcol, Syy);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const float m = Sy[k] * ia;
        tsp[(2 * c) * C::TS + k * TW] = m;
        tsp[(2 * c + 1) * C::TS + k * TW] = fmaf(-m, m, Syy[k] * ia) + kC2;
      }
      if (c < 2) __syncthreads();
    }
  }

  float avg_acc[PPT], reg_acc[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) { avg_acc[k] = 0.f; reg_acc[k] = 0.f; }

  for (int s = 0; s < S; ++s) {
    const CamM cam = compose_camera(cams[s]);
    {   // warp source s into the three shared planes (R halo included), two elements per trip
      const float4* src = p.src[s] + (size_t)b * plane;
      const float eps = p.d.eps;
      float4 ta[2], tb[2], tc[2], td[2];
      float fx[2], fy[2];
      for_region2<C::PH, C::PW, NT>(
          [&](int j, int lr, int lc, bool live) {
            float pu, pv;
            project_composed(cam, (float)colt[lc].w, (float)rowt[lr].w, dpl[lr * C::LD + lc], eps, pu, pv);
            const Taps4 t = make_taps4(pu, pv, H, W);
            fx[j] = t.fx; fy[j] = t.fy;
            if (live) {
              const float4* q = src + t.o;
              ta[j] = __ldg(q); tb[j] = __ldg(q + 1); tc[j] = __ldg(q + W); td[j] = __ldg(q + W + 1);
            }
          },
          [&](int j, int lr, int lc) {
            const int o = lr * C::LD + lc;
            const float w00 = (1.f - fx[j]) * (1.f - fy[j]), w01 = fx[j] * (1.f - fy[j]);
            const float w10 = (1.f - fx[j]) * fy[j], w11 = fx[j] * fy[j];
            wp[o] = ta[j].x * w00 + tb[j].x * w01 + tc[j].x * w10 + td[j].x * w11;
            wp[C::PLANE + o] = ta[j].y * w00 + tb[j].y * w01 + tc[j].y * w10 + td[j].y * w11;
            wp[2 * C::PLANE + o] = ta[j].z * w00 + tb[j].z * w01 + tc[j].z * w10 + td[j].z * w11;
          });
    }
    __syncthreads();

    float occ_w[PPT];
    if (OCC) {   // own pixels: sample the source frame's depth at the projected position (4 scalar taps)
      const float* rd = p.ref[s] + (size_t)b * plane;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int o = (prow0 + k + R) * C::LD + pcol + R;
        const float d = dpl[o];
        float pu, pv;
        project_composed(cam, (float)(u0 + pcol), (float)(v0 + prow0 + k), d, p.d.eps, pu, pv);
        const Taps4 t = make_taps4(pu, pv, H, W);
        const bool live = col_in && (v0 + prow0 + k < H);
        float pd = d;
        if (live) {
          const float* q = rd + t.o;
          pd = (__ldg(q) * (1.f - t.fx) + __ldg(q + 1) * t.fx) * (1.f - t.fy) +
               (__ldg(q + W) * (1.f - t.fx) + __ldg(q + W + 1) * t.fx) * t.fy;
        }
        const OccTerm oc = occ_term(d, pd, wp[o], wp[C::PLANE + o], wp[2 * C::PLANE + o]);
        occ_w[k] = oc.weight;
        if (live && oc.valid) reg_acc[k] += oc.diff;
      }
    }

    float ssim_acc[PPT], l1_acc[PPT], l1w[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int o = (prow0 + k + R) * C::LD + pcol + R;
      l1_acc[k] = fabsf(tg[o] - wp[o]) + fabsf(tg[C::PLANE + o] - wp[C::PLANE + o]) +
                  fabsf(tg[2 * C::PLANE + o] - wp[2 * C::PLANE + o]);
      l1w[k] = p.d.w_l1 * (l1_acc[k] * (1.f / 3.f));
      ssim_acc[k] = 0.f;
    }
    if (R > 0) {
      if (MERGED) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          hpass_blocked<RR, C::PH, TW, C::LD, TW, true>(wp + c * C::PLANE, tg + c * C::PLANE, hbA + (3 * c) * C::HB,
                                                        hbA + (3 * c + 1) * C::HB, hbA + (3 * c + 2) * C::HB);
      } else {
        hpass_blocked<RR, C::PH, TW, C::LD, TW, true>(wp, tg, hbA, hbA + C::HB, hbA + 2 * C::HB);
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float* cur = MERGED ? hbA + (3 * c) * C::HB : ((c & 1) ? hbB : hbA);
        float* nxt = (c & 1) ? hbA : hbB;
        if (!MERGED && c < 2)
          hpass_blocked<RR, C::PH, TW, C::LD, TW, true>(wp + (c + 1) * C::PLANE, tg + (c + 1) * C::PLANE, nxt,
                                                        nxt + C::HB, nxt + 2 * C::HB);
        float Sx[PPT], Sxx[PPT], Sxy[PPT];
        vsum_multi<R, TW, PPT>(cur, prow0, pcol, Sx);
        vsum_multi<R, TW, PPT>(cur + C::HB, prow0, pcol, Sxx);
        vsum_multi<R, TW, PPT>(cur + 2 * C::HB, prow0, pcol, Sxy);
        float* cbase = coef_sc ? coef_sc + ((((size_t)b * S + s) * 3 + c) * 3) * plane + pix0 : nullptr;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          // layers.py:37-46 with one reciprocal, shared by the value and its derivative coefficients
          const float mx = Sx[k] * ia, myk = tsp[(2 * c) * C::TS + k * TW];
          const float sxx = fmaf(-mx, mx, Sxx[k] * ia);
          const float sxy = fmaf(-mx, myk, Sxy[k] * ia);
          const float n1 = fmaf(2.f * mx, myk, kC1), n2 = fmaf(2.f, sxy, kC2);
          const float d1 = fmaf(mx, mx, fmaf(myk, myk, kC1)), d2 = sxx + tsp[(2 * c + 1) * C::TS + k * TW];
          const float inv_d = rcp_fast(d1 * d2);
          const float q = n1 * n2 * inv_d;
          const float val = 0.5f - 0.5f * q;
          ssim_acc[k] += fminf(fmaxf(val, 0.f), 1.f);
          // Lower bound of this source's loss from the channels seen so far (the SSIM terms still to come are >= 0, and
          // the float sums / products below are monotone): once it reaches the running minimum the source cannot win this
          // pixel any more, and its coefficients -- which the backward reads only where the arg-min selects the source --
          // are neither computed nor stored.  (--avg_reprojection averages the sources: every coefficient is needed.)
          float lb = fmaf(p.d.w_ssim, ssim_acc[k] * (1.f / 3.f), l1w[k]);
          if (OCC) lb *= occ_w[k];
          if (cbase && col_in && v0 + prow0 + k < H && (avg || lb < best[k])) {
            // torch.clamp backward: zero outside [0,1] (NaN -> 0); branch-free
            const bool ok = val >= 0.f && val <= 1.f;
            const float dS_dn = -0.5f * inv_d, dS_dd = 0.5f * q * inv_d;
            float gmx = ok ? dS_dn * (2.f * myk * (n2 - n1)) + dS_dd * (2.f * mx * (d2 - d1)) : 0.f;
            float gxx = ok ? dS_dd * d1 : 0.f;
            float gxy = ok ? dS_dn * 2.f * n1 : 0.f;
            if (OCC) { gmx *= occ_w[k]; gxx *= occ_w[k]; gxy *= occ_w[k]; }   // the (detached) per-pixel loss weight
            float* cp = cbase + (size_t)k * W;
            __stcs(cp, gmx); __stcs(cp + plane, gxx); __stcs(cp + 2 * plane, gxy);
          }
        }
        if (!MERGED && c < 2) __syncthreads();
      }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      float rho;
      if (R > 0) rho = fmaf(p.d.w_ssim, ssim_acc[k] * (1.f / 3.f), l1w[k]);   // (same expression as the pruning bound)
      else rho = l1_acc[k] * (1.f / 3.f);
      if (OCC) rho *= occ_w[k];
      if (avg) {
        avg_acc[k] += rho;
      } else if (rho < best[k]) {
        best[k] = rho; arg[k] = n_ident + s;
      }
    }
    // the next source's warp may overwrite wp now: every read of wp happened before the last barrier above
    // (R > 0) or happens before the barrier below (R == 0)
    if (R == 0) __syncthreads();
  }
  if (avg) {
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const float rho = avg_acc[k] / (float)S;
      if (rho < best[k]) { best[k] = rho; arg[k] = n_ident; }
    }
  }

  float local = 0.f;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    if (col_in && v0 + prow0 + k < H) {
      local += best[k];
      argmin_sc[(size_t)b * plane + pix0 + (size_t)k * W] = (uint8_t)arg[k];
    }
  }
  const float tot = block_sum(local, red);
  if (threadIdx.x == 0) p.partial[(MS ? (size_t)sc * p.ms_partial_stride : 0) + cta_index] = tot;
  if (OCC) {
    float lr = 0.f;
#pragma unroll
    for (int k = 0; k < PPT; ++k) lr += reg_acc[k];
    __syncthreads();
    const float treg = block_sum(lr, red);
    if (threadIdx.x == 0) p.partial_reg[cta_index] = treg;
  }
  if (MS) __syncthreads();   // the next scale rewrites the camera slots, the depth plane and `red`
  }  // scales
}

// ------------------------------------------------------------------------------------------------
// identity (auto-mask) losses of every source in one launch: compute_reprojection_loss(source_f, target) for all f
// (trainer.py:480-493).  Same tile machinery as the fused forward without the warp: the target tile and its box
// statistics are staged / computed once per CTA and shared by the S sources.
// ------------------------------------------------------------------------------------------------
struct IdentParams {
  int B, H, W, S;
  float w_ssim, w_l1;
  const float* target;                    // [B,3,H,W]
  const float* src[SQLX_MAX_SOURCES];     // planar [B,3,H,W]
  float* out;                             // [B,S,H,W]
};

template <int R, int TH, int TW, int NT>
struct Ident3Cfg {
  static constexpr int PH = TH + 2 * R, PW = TW + 2 * R;
  static constexpr int LD = ((PW + 3) & ~3) + 2;
  static constexpr int PLANE = PH * LD;
  static constexpr int HB = PH * TW;
  static constexpr int PPT = (TH * TW) / NT;
  static constexpr int TS = TH * TW;
  static constexpr size_t smem_bytes = sizeof(float) * (6 * PLANE + 9 * HB + 6 * TS) + sizeof(int4) * (PH + PW);
  static_assert((TH * TW) % NT == 0 && NT % TW == 0 && NT >= PH + PW, "tile / block shape");
};

template <int R, int TH, int TW, int NT>
__global__ void __launch_bounds__(NT, 2) identity3_kernel(const IdentParams p) {
  using C = Ident3Cfg<R, TH, TW, NT>;
  constexpr int PPT = C::PPT;
  constexpr int RR = R > 0 ? R : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tg = reinterpret_cast<float*>(smem_raw);   // 3 planes
  float* wp = tg + 3 * C::PLANE;                    // 3 planes
  float* hb = wp + 3 * C::PLANE;                    // 9 planes of HB
  float* tstat = hb + 9 * C::HB;                    // [3 ch][mean, variance + C2][TH*TW]
  int4* rowt = reinterpret_cast<int4*>(tstat + 6 * C::TS);
  int4* colt = rowt + C::PH;
  const int H = p.H, W = p.W, S = p.S;
  const int b = blockIdx.z;
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  fill_axis_table<C::PH>(rowt, v0 - R, H, H, 1.f, 0, 0);
  fill_axis_table<C::PW>(colt, u0 - R, W, W, 1.f, 0, C::PH);
  __syncthreads();
  auto stage3 = [&](const float* __restrict__ img, float* __restrict__ dst) {
    float tv[2][3];
    for_region2<C::PH, C::PW, NT>(
        [&](int j, int lr, int lc, bool live) {
          if (live) {
            const float* tp = img + (size_t)(rowt[lr].w * W + colt[lc].w);
            tv[j][0] = __ldg(tp); tv[j][1] = __ldg(tp + plane); tv[j][2] = __ldg(tp + 2 * plane);
          }
        },
        [&](int j, int lr, int lc) {
          const int o = lr * C::LD + lc;
          dst[o] = tv[j][0]; dst[C::PLANE + o] = tv[j][1]; dst[2 * C::PLANE + o] = tv[j][2];
        });
  };
  stage3(p.target + (size_t)b * 3 * plane, tg);
  __syncthreads();
  const int pcol = threadIdx.x % TW;
  const int prow0 = (threadIdx.x / TW) * PPT;
  const bool col_in = u0 + pcol < W;
  const size_t pix0 = (size_t)(v0 + prow0) * W + (u0 + pcol);
  float* tsp = tstat + prow0 * TW + pcol;
  if (R > 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      hpass_blocked<RR, C::PH, TW, C::LD, TW, false>(nullptr, tg + c * C::PLANE, hb + (2 * c) * C::HB,
                                                     hb + (2 * c + 1) * C::HB, nullptr);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float Sy[PPT], Syy[PPT];
      vsum_multi<R, TW, PPT>(hb + (2 * c) * C::HB, prow0, pcol, Sy);
      vsum_multi<R, TW, PPT>(hb + (2 * c + 1) * C::HB, prow0, pcol, Syy);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const float m = Sy[k] * ia;
        tsp[(2 * c) * C::TS + k * TW] = m;
        tsp[(2 * c + 1) * C::TS + k * TW] = fmaf(-m, m, Syy[k] * ia) + kC2;
      }
    }
  }
  for (int s = 0; s < S; ++s) {
    stage3(p.src[s] + (size_t)b * 3 * plane, wp);     // (wp / hb readers of the previous source are past its barriers)
    __syncthreads();
    float ssim_acc[PPT], l1_acc[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int o = (prow0 + k + R) * C::LD + pcol + R;
      l1_acc[k] = fabsf(tg[o] - wp[o]) + fabsf(tg[C::PLANE + o] - wp[C::PLANE + o]) +
                  fabsf(tg[2 * C::PLANE + o] - wp[2 * C::PLANE + o]);
      ssim_acc[k] = 0.f;
    }
    if (R > 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        hpass_blocked<RR, C::PH, TW, C::LD, TW, true>(wp + c * C::PLANE, tg + c * C::PLANE, hb + (3 * c) * C::HB,
                                                      hb + (3 * c + 1) * C::HB, hb + (3 * c + 2) * C::HB);
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float Sx[PPT], Sxx[PPT], Sxy[PPT];
        vsum_multi<R, TW, PPT>(hb + (3 * c) * C::HB, prow0, pcol, Sx);
        vsum_multi<R, TW, PPT>(hb + (3 * c + 1) * C::HB, prow0, pcol, Sxx);
        vsum_multi<R, TW, PPT>(hb + (3 * c + 2) * C::HB, prow0, pcol, Sxy);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          const float mx = Sx[k] * ia, myk = tsp[(2 * c) * C::TS + k * TW];
          const float sxx = fmaf(-mx, mx, Sxx[k] * ia);
          const float sxy = fmaf(-mx, myk, Sxy[k] * ia);
          const float n1 = fmaf(2.f * mx, myk, kC1), n2 = fmaf(2.f, sxy, kC2);
          const float d1 = fmaf(mx, mx, fmaf(myk, myk, kC1)), d2 = sxx + tsp[(2 * c + 1) * C::TS + k * TW];
          const float val = 0.5f - 0.5f * (n1 * n2 * rcp_fast(d1 * d2));
          ssim_acc[k] += fminf(fmaxf(val, 0.f), 1.f);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (!(col_in && v0 + prow0 + k < H)) continue;
      const float rho = R > 0 ? p.w_ssim * (ssim_acc[k] * (1.f / 3.f)) + p.w_l1 * (l1_acc[k] * (1.f / 3.f))
                              : l1_acc[k] * (1.f / 3.f);
      p.out[((size_t)b * S + s) * plane + pix0 + (size_t)k * W] = rho;
    }
    __syncthreads();   // wp and hb are rewritten by the next source
  }
}

// Deterministic final reduction of per-CTA partial sums (one block; double accumulation).
__global__ void finalize_sum3_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sh[0];
}

// ------------------------------------------------------------------------------------------------
// backward from the saved SSIM coefficients
// ------------------------------------------------------------------------------------------------
struct PhotoBwdParams {
  sqlx_photo_desc d;
  const float* depth_up;   // optional [B,H,W]
  const float* depth_lr;
  const float* target;
  const float4* src[SQLX_MAX_SOURCES];
  const float* K;
  const float* invK;
  const float* T;
  const uint8_t* argmin;
  const float* coef;
  const float* g_loss;
  float scale;
  float* d_depth_lr;   // [B,h,w] accumulated through the bilinear-upsample adjoint (atomics), or NULL when g_up is used
  float* g_up;         // [B,H,W] gradient wrt the UPSAMPLED depth, one plain store per pixel (no atomics); the caller
                       // applies the upsample adjoint as a gather (multiscale.cu).  NULL -> d_depth_lr path
  float* q_up;         // optional [B,H,W]: 1 / d_up^2 (the upsample-adjoint kernel needs it for the mean-inverse-depth term)
  int g_up_accumulate; // 1: g_up += (the smoothness backward already wrote the plane), 0: g_up =
  float* dP;  // [B,S,12] accumulators (zeroed by the host wrapper)
  // indoor variant (OCC kernels only)
  const float* ref[SQLX_MAX_SOURCES];    // depth of each source frame, planar [B,H,W]
  float* d_ref[SQLX_MAX_SOURCES];        // its gradient, accumulated with atomics (caller zeroes)
  const float* g_reg;                    // device scalar: upstream gradient of the regularisation sum
  // all loss scales in one launch (MS kernels only): blockIdx.z = scale * B + sample
  int ns;
  const float* ms_depth_up[SQLX_MAX_SCALES];
  const uint8_t* ms_argmin[SQLX_MAX_SCALES];
  float* ms_g_up[SQLX_MAX_SCALES];
  float* ms_q_up[SQLX_MAX_SCALES];       // entries may be NULL
  unsigned ms_accumulate;                // bit i: g_up of scale i is accumulated into
  size_t ms_T_stride, ms_coef_stride, ms_dP_stride;   // in floats
};

// multiplicity with which the window of output q (coordinate qi) covers pixel i under reflection padding
template <int R>
__device__ __forceinline__ float reflect_mult3(int i, int qi, int n) {
  float m = 1.f;  // |qi - i| <= R is guaranteed by the caller's loop bounds
  if (i >= 1 && i <= R && qi + i <= R) m += 1.f;
  if (i <= n - 2 && i >= n - 1 - R && 2 * (n - 1) - i - qi <= R) m += 1.f;
  return m;
}

template <int R, int TH, int TW, int NT>
struct Bwd3Cfg {
  static constexpr int PH1 = TH + 2 * R, PW1 = TW + 2 * R;
  static constexpr int LD = ((PW1 + 3) & ~3) + 2;
  static constexpr int CF = PH1 * LD;             // one coefficient plane on the R halo
  static constexpr int H2 = PH1 * TW;             // one horizontally filtered plane
  static constexpr int PPT = (TH * TW) / NT;
  static constexpr int LRH = TH + 2, LRW = TW + 2;
  static constexpr int SCR = (9 * H2 > LRH * LRW) ? 9 * H2 : LRH * LRW;
  static constexpr size_t smem_bytes = sizeof(float) * (9 * CF + SCR + 32 + 16 * SQLX_MAX_SOURCES) +
                                       sizeof(Camera) * SQLX_MAX_SOURCES + PH1 * PW1 + 16;
  static_assert((TH * TW) % NT == 0 && NT % TW == 0, "tile / block shape");
};

template <int R, int TH, int TW, int NT, int MINB, bool OCC = false, bool MS = false>
__global__ void __launch_bounds__(NT, MINB) photo_bwd3_kernel(const __grid_constant__ PhotoBwdParams p) {
  using C = Bwd3Cfg<R, TH, TW, NT>;
  constexpr int PPT = C::PPT;
  constexpr int RR = R > 0 ? R : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* cf = reinterpret_cast<float*>(smem_raw);   // 9 planes on the R halo (3 channels x 3 coefficients)
  float* h2 = cf + 9 * C::CF;                        // 9 planes PH1 x TW ; later the low-res accumulation scratch
  float* red = h2 + C::SCR;
  float* dPs = red + 32;                             // [S][16]
  Camera* cams = reinterpret_cast<Camera*>(dPs + 16 * SQLX_MAX_SOURCES);
  uint8_t* amin = reinterpret_cast<uint8_t*>(cams + SQLX_MAX_SOURCES);

  const int H = p.d.H, W = p.d.W, S = p.d.S;
  const int sc = MS ? (int)blockIdx.z / p.d.B : 0;
  const int b = MS ? (int)blockIdx.z - sc * p.d.B : (int)blockIdx.z;
  const float* depth_up = MS ? p.ms_depth_up[sc] : p.depth_up;
  const uint8_t* argmin_p = MS ? p.ms_argmin[sc] : p.argmin;
  const float* T_p = MS ? p.T + (size_t)sc * p.ms_T_stride : p.T;
  const float* coef_p = (MS && p.coef) ? p.coef + (size_t)sc * p.ms_coef_stride : p.coef;
  float* dP_p = MS ? p.dP + (size_t)sc * p.ms_dP_stride : p.dP;
  float* g_up_p = MS ? p.ms_g_up[sc] : p.g_up;
  float* q_up_p = MS ? p.ms_q_up[sc] : p.q_up;
  const int g_up_acc = MS ? (int)((p.ms_accumulate >> sc) & 1u) : p.g_up_accumulate;
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  const bool automask = p.d.flags & SQLX_AUTOMASK;
  const bool avg = p.d.flags & SQLX_AVG_REPROJ;
  const int n_ident = automask ? (avg ? 1 : S) : 0;
  const float gscale = __ldg(p.g_loss) * p.scale;
  constexpr float ia = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  const int h = p.d.h, w = p.d.w;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  // every own pixel lies more than R from the frame border: all reflect-pad multiplicities are 1
  const bool interior = (v0 > R) && (u0 > R) && (v0 + TH + R < H) && (u0 + TW + R < W);

  if (threadIdx.x < S)
    load_camera(p.K + b * 16, p.invK + b * 16, T_p + ((size_t)b * S + threadIdx.x) * 16, cams[threadIdx.x]);
  if (threadIdx.x < 16 * SQLX_MAX_SOURCES) dPs[threadIdx.x] = 0.f;
  {   // arg-min of the tile's R halo, two elements per trip (loads of both in flight)
    uint8_t av[2];
    for_region2<C::PH1, C::PW1, NT>(
        [&](int j, int lr, int lc, bool live) {
          const int v = v0 - R + lr, u = u0 - R + lc;
          av[j] = (live && v >= 0 && v < H && u >= 0 && u < W) ? argmin_p[(size_t)b * plane + (size_t)v * W + u]
                                                               : (uint8_t)255;
        },
        [&](int j, int lr, int lc) { amin[lr * C::PW1 + lc] = av[j]; });
  }

  const int pcol = threadIdx.x % TW;
  const int prow0 = (threadIdx.x / TW) * PPT;
  const bool col_in = u0 + pcol < W;
  const size_t pix0 = (size_t)(v0 + prow0) * W + (u0 + pcol);
  bool pin[PPT];
  float gd[PPT], dep[PPT], Yv[3][PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    pin[k] = col_in && (v0 + prow0 + k < H);
    gd[k] = 0.f;
    dep[k] = 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) Yv[c][k] = 0.f;
    if (pin[k]) {
      dep[k] = depth_up ? __ldg(depth_up + (size_t)b * plane + pix0 + (size_t)k * W)
                          : upsample_at(p.depth_lr + (size_t)b * h * w, h, w, v0 + prow0 + k, u0 + pcol, sy, sx);
#pragma unroll
      for (int c = 0; c < 3; ++c) Yv[c][k] = __ldg(p.target + ((size_t)b * 3 + c) * plane + pix0 + (size_t)k * W);
    }
  }
  __syncthreads();

  for (int s = 0; s < S; ++s) {
    const float sel_w = avg ? 1.f / (float)S : 1.f;
    const int sel_idx = avg ? n_ident : n_ident + s;
    if (!OCC) {   // nothing on this tile's halo selects source s (auto-masked or won by another source): no gradient
                  // at all (the indoor regularisation term has a gradient at every valid pixel: never skipped)
      int any = 0;
      for (int idx = threadIdx.x; idx < C::PH1 * C::PW1; idx += NT) any |= (amin[idx] == sel_idx);
      if (!__syncthreads_or(any)) continue;
    }
    const Camera cam = cams[s];
    const float4* srcb = p.src[s] + (size_t)b * plane;
    // own pixels: warped value and its spatial derivatives per channel
    float Xv[3][PPT], dXx[3][PPT], dXy[3][PPT];
    float gix[PPT], giy[PPT], occ_w[PPT];
    const float greg = OCC ? __ldg(p.g_reg) * p.scale : 0.f;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      gix[k] = 0.f; giy[k] = 0.f; occ_w[k] = 1.f;
      const Proj pr = project_fast(cam, (float)(u0 + pcol), (float)(v0 + prow0 + k), dep[k], p.d.eps);
      const Taps4 t = make_taps4(pr.pu, pr.pv, H, W);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = a, cq = a, dq = a;
      float ra = 0.f, rb = 0.f, rc = 0.f, rd = 0.f;
      if (pin[k]) {
        const float4* q = srcb + t.o;
        a = __ldg(q); bq = __ldg(q + 1); cq = __ldg(q + W); dq = __ldg(q + W + 1);
        if (OCC) {
          const float* r = p.ref[s] + (size_t)b * plane + t.o;
          ra = __ldg(r); rb = __ldg(r + 1); rc = __ldg(r + W); rd = __ldg(r + W + 1);
        }
      }
      const float w00 = (1.f - t.fx) * (1.f - t.fy), w01 = t.fx * (1.f - t.fy);
      const float w10 = (1.f - t.fx) * t.fy, w11 = t.fx * t.fy;
      Xv[0][k] = a.x * w00 + bq.x * w01 + cq.x * w10 + dq.x * w11;
      Xv[1][k] = a.y * w00 + bq.y * w01 + cq.y * w10 + dq.y * w11;
      Xv[2][k] = a.z * w00 + bq.z * w01 + cq.z * w10 + dq.z * w11;
      dXx[0][k] = (bq.x - a.x) * (1.f - t.fy) + (dq.x - cq.x) * t.fy;
      dXx[1][k] = (bq.y - a.y) * (1.f - t.fy) + (dq.y - cq.y) * t.fy;
      dXx[2][k] = (bq.z - a.z) * (1.f - t.fy) + (dq.z - cq.z) * t.fy;
      dXy[0][k] = (cq.x - a.x) * (1.f - t.fx) + (dq.x - bq.x) * t.fx;
      dXy[1][k] = (cq.y - a.y) * (1.f - t.fx) + (dq.y - bq.y) * t.fx;
      dXy[2][k] = (cq.z - a.z) * (1.f - t.fx) + (dq.z - bq.z) * t.fx;
      if (OCC && pin[k]) {
        // depth-consistency terms (trainer_indoor.py:636-651, 699): the loss weight is detached; the regularisation
        // sum diff * valid has a gradient wrt the depth, the sampled source depth and (through it) the position
        const float d = dep[k];
        const float pd = (ra * (1.f - t.fx) + rb * t.fx) * (1.f - t.fy) + (rc * (1.f - t.fx) + rd * t.fx) * t.fy;
        const OccTerm oc = occ_term(d, pd, Xv[0][k], Xv[1][k], Xv[2][k]);
        occ_w[k] = oc.weight;
        if (oc.valid && greg != 0.f) {
          const float inv = 1.f / (d + pd);
          const float sg = d > pd ? 1.f : (d < pd ? -1.f : 0.f);
          gd[k] += greg * (sg - oc.diff) * inv;
          const float gpd = greg * (-sg - oc.diff) * inv;
          gix[k] = gpd * ((rb - ra) * (1.f - t.fy) + (rd - rc) * t.fy);
          giy[k] = gpd * ((rc - ra) * (1.f - t.fx) + (rd - rb) * t.fx);
          float* q = p.d_ref[s] + (size_t)b * plane + t.o;
          atomicAdd(q, gpd * w00); atomicAdd(q + 1, gpd * w01); atomicAdd(q + W, gpd * w10); atomicAdd(q + W + 1, gpd * w11);
        }
      }
    }
    const float alpha = (p.d.w_ssim / 3.f) * sel_w * gscale;
    const float wl1 = ((R > 0) ? p.d.w_l1 : 1.f) / 3.f * sel_w * gscale;
    float gx[3][PPT];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const float diff = Yv[c][k] - Xv[c][k];
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const bool sel = amin[(prow0 + k + R) * C::PW1 + pcol + R] == sel_idx;
        gx[c][k] = sel ? -wl1 * sgn * (OCC ? occ_w[k] : 1.f) : 0.f;
      }
    }
    if (R > 0) {
      // all nine coefficient planes (3 channels x {d/d mean, d/d E[x^2], d/d E[xy]}) of this source in ONE staging
      // pass: nine independent masked loads per halo pixel in flight, two block barriers per source
      const float* cb9 = coef_p + ((size_t)b * S + s) * 9 * plane;
      float cv[2][9];
      for_region2<C::PH1, C::PW1, NT>(
          [&](int j, int lr, int lc, bool live) {
            if (live && amin[lr * C::PW1 + lc] == sel_idx) {
              const float* q = cb9 + (size_t)(v0 - R + lr) * W + (u0 - R + lc);
#pragma unroll
              for (int m = 0; m < 9; ++m) cv[j][m] = __ldg(q + (size_t)m * plane);
            } else {
#pragma unroll
              for (int m = 0; m < 9; ++m) cv[j][m] = 0.f;
            }
          },
          [&](int j, int lr, int lc) {
            const int o = lr * C::LD + lc;
#pragma unroll
            for (int m = 0; m < 9; ++m) cf[m * C::CF + o] = cv[j][m];
          });
      __syncthreads();
      if (interior) {
#pragma unroll
        for (int m = 0; m < 9; ++m) hsum_blocked<RR, C::PH1, TW, C::LD, TW>(cf + m * C::CF, h2 + m * C::H2);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sa[PPT], sb[PPT], sc[PPT];
          vsum_multi<R, TW, PPT>(h2 + (3 * c) * C::H2, prow0, pcol, sa);
          vsum_multi<R, TW, PPT>(h2 + (3 * c + 1) * C::H2, prow0, pcol, sb);
          vsum_multi<R, TW, PPT>(h2 + (3 * c + 2) * C::H2, prow0, pcol, sc);
#pragma unroll
          for (int k = 0; k < PPT; ++k)
            gx[c][k] += alpha * ia * (sa[k] + 2.f * Xv[c][k] * sb[k] + Yv[c][k] * sc[k]);
        }
      } else {
        // frame-border tiles: adjoint of the reflection padding = per-tap multiplicities
        for (int idx = threadIdx.x; idx < C::PH1 * TW; idx += NT) {
          const int lr = idx / TW, pc = idx - lr * TW;
          const int u = u0 + pc;
          const float* ca = cf + lr * C::LD + pc;
          float t9[9];
#pragma unroll
          for (int m = 0; m < 9; ++m) t9[m] = 0.f;
          if (u < W) {
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
              const int qu = u - R + k;
              if (qu < 0 || qu >= W) continue;
              const float mult = reflect_mult3<R>(u, qu, W);
#pragma unroll
              for (int m = 0; m < 9; ++m) t9[m] += mult * ca[m * C::CF + k];
            }
          }
#pragma unroll
          for (int m = 0; m < 9; ++m) h2[m * C::H2 + idx] = t9[m];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          const int v = v0 + prow0 + k;
          const float* ha = h2 + (prow0 + k) * TW + pcol;
          float t9[9];
#pragma unroll
          for (int m = 0; m < 9; ++m) t9[m] = 0.f;
          if (v < H) {
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) {
              const int qv = v - R + j;
              if (qv < 0 || qv >= H) continue;
              const float mult = reflect_mult3<R>(v, qv, H);
#pragma unroll
              for (int m = 0; m < 9; ++m) t9[m] += mult * ha[m * C::H2 + j * TW];
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c)
            gx[c][k] += alpha * ia * (t9[3 * c] + 2.f * Xv[c][k] * t9[3 * c + 1] + Yv[c][k] * t9[3 * c + 2]);
        }
      }
      // cf is rewritten by the next source's staging: its readers finished before the second barrier; h2 is rewritten
      // only after the next staging barrier
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        gix[k] = fmaf(gx[c][k], dXx[c][k], gix[k]);
        giy[k] = fmaf(gx[c][k], dXy[c][k], giy[k]);
      }
    }
    float dPacc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dPacc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (!pin[k]) continue;
      if (gix[k] == 0.f && giy[k] == 0.f) continue;
      const float uf = (float)(u0 + pcol), vf = (float)(v0 + prow0 + k);
      const Proj pr = project_fast(cam, uf, vf, dep[k], p.d.eps);
      // grid_sample clips the coordinate: no gradient through a clipped axis (ATen clip_coordinates_set_grad)
      const float gx_ = (pr.pu > 0.f && pr.pu < (float)(W - 1)) ? gix[k] : 0.f;
      const float gy_ = (pr.pv > 0.f && pr.pv < (float)(H - 1)) ? giy[k] : 0.f;
      const float g0 = gx_ * pr.rz, g1 = gy_ * pr.rz, g2 = -(gx_ * pr.pu + gy_ * pr.pv) * pr.rz;
      const float r0 = fmaf(cam.iK[0], uf, fmaf(cam.iK[1], vf, cam.iK[2]));
      const float r1 = fmaf(cam.iK[3], uf, fmaf(cam.iK[4], vf, cam.iK[5]));
      const float r2 = fmaf(cam.iK[6], uf, fmaf(cam.iK[7], vf, cam.iK[8]));
      gd[k] += g0 * (cam.P[0] * r0 + cam.P[1] * r1 + cam.P[2] * r2) + g1 * (cam.P[4] * r0 + cam.P[5] * r1 + cam.P[6] * r2) +
               g2 * (cam.P[8] * r0 + cam.P[9] * r1 + cam.P[10] * r2);
      dPacc[0] += g0 * pr.X0; dPacc[1] += g0 * pr.X1; dPacc[2] += g0 * pr.X2; dPacc[3] += g0;
      dPacc[4] += g1 * pr.X0; dPacc[5] += g1 * pr.X1; dPacc[6] += g1 * pr.X2; dPacc[7] += g1;
      dPacc[8] += g2 * pr.X0; dPacc[9] += g2 * pr.X1; dPacc[10] += g2 * pr.X2; dPacc[11] += g2;
    }
    warp_reduce16(dPacc);
    {
      const int lane = threadIdx.x & 31;
      if (!(lane & 1) && (lane >> 1) < 12 && dPacc[0] != 0.f) atomicAdd(&dPs[s * 16 + (lane >> 1)], dPacc[0]);
    }
  }

  __syncthreads();
  if (threadIdx.x < 12 * S) {
    const int s = threadIdx.x / 12, i = threadIdx.x - s * 12;
    const float t = dPs[s * 16 + i];
    if (t != 0.f) atomicAdd(dP_p + ((size_t)b * S + s) * 12 + i, t);
  }
  if (g_up_p) {   // hand the per-pixel gradient to the gather-style upsample adjoint
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (!pin[k]) continue;
      const size_t o = (size_t)b * plane + pix0 + (size_t)k * W;
      float* q = g_up_p + o;
      *q = g_up_acc ? *q + gd[k] : gd[k];
      if (q_up_p) q_up_p[o] = __fdividef(1.f, dep[k] * dep[k]);
    }
    return;
  }
  // adjoint of the bilinear upsampling: per-tile accumulation in shared memory, one global atomic per touched cell
  const int vend = min(v0 + TH, H) - 1, uend = min(u0 + TW, W) - 1;
  const int i_lo = up_tap(v0, sy, h).i0, i_hi = up_tap(vend, sy, h).i1;
  const int j_lo = up_tap(u0, sx, w).i0, j_hi = up_tap(uend, sx, w).i1;
  const int nh = i_hi - i_lo + 1, nw = j_hi - j_lo + 1;
  float* acc = h2;
  for (int idx = threadIdx.x; idx < nh * nw; idx += NT) acc[idx] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    if (!pin[k] || gd[k] == 0.f) continue;
    const UpTap ty = up_tap(v0 + prow0 + k, sy, h), tx = up_tap(u0 + pcol, sx, w);
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
__global__ void dT_from_dP3_kernel(const float* __restrict__ K, const float* __restrict__ dP, int B, int S,
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

}  // namespace sqlx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace sqlx;

namespace {
// Tile configurations, picked on a B200 (A/B of round 1 and round 2, profiles/r02a_candidates.log: every alternative
// -- 16x32 forward tiles at 3 or 4 CTAs/SM, one-trip staging, 32x32 or uncapped backward tiles -- measured slower or
// equal and was removed):
//   forward : 32x32 tile, 256 threads (4 px/thread), 2 CTAs/SM        backward: 16x32 tile, 256 threads, 3 CTAs/SM
constexpr int kMinTH = 16, kMinTW = 32;   // smallest tile of any configuration: sizes the per-CTA partial buffer

template <int R, int TH, int TW, int NT, int MINB, bool MERGED = false, bool OCC = false, bool MS = false>
int launch_photo_fwd3(const PhotoFwdParams& p, int* ctas, cudaStream_t st) {
  using C = Fwd3Cfg<R, TH, TW, NT, MERGED>;
  auto kern = photo_fwd3_kernel<R, TH, TW, NT, MINB, MERGED, OCC, MS>;
  if (int e = ensure_dyn_smem(kern, C::smem_bytes)) return e;
  dim3 grid(ceil_div(p.d.W, TW), ceil_div(p.d.H, TH), p.d.B);
  *ctas = (int)(grid.x * grid.y * grid.z);
  ProfScope prof(OCC ? "photo_occ_fwd_kernel" : (MS ? "photo_fwd_ms_kernel" : "photo_fwd_kernel"), st);
  kern<<<grid, NT, C::smem_bytes, st>>>(p);
  return check_launch("photo_fwd3_kernel");
}

template <int R>
int dispatch_photo_fwd3(const PhotoFwdParams& p, int* ctas, cudaStream_t st) {
  if (p.partial_reg) return launch_photo_fwd3<R, 32, 32, 256, 2, false, true>(p, ctas, st);   // indoor variant
  if (p.ns > 0) return launch_photo_fwd3<R, 32, 32, 256, 2, false, false, true>(p, ctas, st);   // all scales per CTA
  return launch_photo_fwd3<R, 32, 32, 256, 2>(p, ctas, st);
}

template <int R, int TH, int TW, int NT, int MINB, bool OCC = false, bool MS = false>
int launch_photo_bwd3(const PhotoBwdParams& p, cudaStream_t st) {
  using C = Bwd3Cfg<R, TH, TW, NT>;
  auto kern = photo_bwd3_kernel<R, TH, TW, NT, MINB, OCC, MS>;
  if (int e = ensure_dyn_smem(kern, C::smem_bytes)) return e;
  dim3 grid(ceil_div(p.d.W, TW), ceil_div(p.d.H, TH), p.d.B * (MS ? p.ns : 1));
  ProfScope prof(OCC ? "photo_occ_bwd_kernel" : (MS ? "photo_bwd_ms_kernel" : "photo_bwd_kernel"), st);
  kern<<<grid, NT, C::smem_bytes, st>>>(p);
  return check_launch("photo_bwd3_kernel");
}

template <int R>
int dispatch_photo_bwd3(const PhotoBwdParams& p, cudaStream_t st) {
  if (p.g_reg) return launch_photo_bwd3<R, 16, 32, 256, 2, true>(p, st);   // indoor variant
  if (p.ns > 0) return launch_photo_bwd3<R, 16, 32, 256, 3, false, true>(p, st);   // all scales in one launch
  return launch_photo_bwd3<R, 16, 32, 256, 3>(p, st);
}

int check_desc(const sqlx_photo_desc* d) {
  SQLX_REQUIRE(d != nullptr, "desc is NULL");
  SQLX_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->h > 0 && d->w > 0, "non-positive shape");
  SQLX_REQUIRE(d->S >= 1 && d->S <= SQLX_MAX_SOURCES, "S=%d outside 1..%d", d->S, SQLX_MAX_SOURCES);
  SQLX_REQUIRE(d->h <= d->H && d->w <= d->W, "depth map larger than the image is not supported");
  SQLX_REQUIRE((d->flags & SQLX_NO_SSIM) || d->ssim_radius == 1 || d->ssim_radius == 3,
               "ssim_radius must be 1 or 3 (got %d)", d->ssim_radius);
  const int r = (d->flags & SQLX_NO_SSIM) ? 0 : d->ssim_radius;
  SQLX_REQUIRE(d->H > 2 * r && d->W > 2 * r && d->H >= 2 && d->W >= 2, "image smaller than the SSIM window");
  SQLX_REQUIRE((long long)d->H * d->W < (1ll << 30), "frame too large for 32-bit pixel offsets");
  return SQLX_OK;
}

size_t fwd_ctas(const sqlx_photo_desc* d) { return (size_t)ceil_div(d->W, kMinTW) * ceil_div(d->H, kMinTH) * d->B; }
}  // namespace

extern "C" int sqlx_pack_rgba(const float* image, int B, int H, int W, float* out, void* stream) {
  SQLX_REQUIRE(image && out, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "bad shape");
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "output must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * H * W;
  const int blocks = (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs);
  pack_rgba_kernel<<<blocks, 256, 0, st>>>(image, B, H * W, reinterpret_cast<float4*>(out));
  return check_launch("pack_rgba_kernel");
}

extern "C" size_t sqlx_photo_workspace_bytes(const sqlx_photo_desc* d) {
  if (!d) return 0;
  // forward: per-CTA partial sums; backward: dP accumulators [B,S,12] (+ scratch)
  return sizeof(float) * (fwd_ctas(d) + (size_t)d->B * SQLX_MAX_SOURCES * 16 + 64);
}

extern "C" size_t sqlx_photo_occ_workspace_bytes(const sqlx_photo_desc* d) {
  if (!d) return 0;
  // two per-CTA partial-sum arrays (photometric, regularisation), then the dP accumulators
  return sizeof(float) * (2 * fwd_ctas(d) + (size_t)d->B * SQLX_MAX_SOURCES * 16 + 64);
}

extern "C" size_t sqlx_photo_coef_bytes(const sqlx_photo_desc* d) {
  if (!d || (d->flags & SQLX_NO_SSIM)) return 0;
  return sizeof(float) * 9 * (size_t)d->B * d->S * d->H * d->W;
}

// internal entry points shared with multiscale.cu (declared in photo_v3.h)
namespace sqlx {
int photo_fwd3_launch(const sqlx_photo_desc* desc, const float* depth_lr, const float* depth_up, const float* target,
                      const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                      const float* identity, const float* noise, float* partial, int* ctas, uint8_t* argmin,
                      float* ssim_coef, cudaStream_t st, const float* const* ref_depths, float* partial_reg) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(depth_lr && target && sources_rgba && K && inv_K && T && partial && ctas && argmin, "NULL pointer argument");
  const bool automask = desc->flags & SQLX_AUTOMASK;
  SQLX_REQUIRE(!automask || (identity && noise), "automask needs identity and noise");
  PhotoFwdParams p;
  p.d = *desc;
  p.depth_lr = depth_lr; p.depth_up = depth_up; p.target = target;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s)
    p.src[s] = s < desc->S ? reinterpret_cast<const float4*>(sources_rgba[s]) : nullptr;
  for (int s = 0; s < desc->S; ++s) {
    SQLX_REQUIRE(p.src[s], "source %d is NULL", s);
    SQLX_REQUIRE((reinterpret_cast<uintptr_t>(p.src[s]) & 15) == 0, "source %d is not 16-byte aligned", s);
  }
  p.K = K; p.invK = inv_K; p.T = T; p.identity = identity; p.noise = noise;
  p.partial = partial;
  p.argmin = argmin;
  p.coef = (desc->flags & SQLX_NO_SSIM) ? nullptr : ssim_coef;
  p.partial_reg = partial_reg;
  p.ns = 0;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s) p.ref[s] = (ref_depths && s < desc->S) ? ref_depths[s] : nullptr;
  if (partial_reg) {
    SQLX_REQUIRE(ref_depths, "the indoor variant needs the source frames' depth maps");
    for (int s = 0; s < desc->S; ++s) SQLX_REQUIRE(p.ref[s], "ref_depths[%d] is NULL", s);
  }
  const int r = (desc->flags & SQLX_NO_SSIM) ? 0 : desc->ssim_radius;
  return r == 3 ? dispatch_photo_fwd3<3>(p, ctas, st)
                : (r == 1 ? dispatch_photo_fwd3<1>(p, ctas, st) : dispatch_photo_fwd3<0>(p, ctas, st));
}

int photo_fwd3_ms_launch(const sqlx_photo_desc* desc, int ns, const float* const* depth_up, const float* target,
                         const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                         size_t T_stride, const float* identity, const float* const* noise, float* partial,
                         size_t partial_stride, int* ctas, uint8_t* const* argmin, float* ssim_coef, size_t coef_stride,
                         cudaStream_t st) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(ns >= 1 && ns <= SQLX_MAX_SCALES, "num_scales %d outside 1..%d", ns, SQLX_MAX_SCALES);
  SQLX_REQUIRE(depth_up && target && sources_rgba && K && inv_K && T && partial && ctas && argmin, "NULL pointer argument");
  const bool automask = desc->flags & SQLX_AUTOMASK;
  SQLX_REQUIRE(!automask || (identity && noise), "automask needs identity and noise");
  PhotoFwdParams p;
  memset(&p, 0, sizeof(p));
  p.d = *desc;
  p.target = target;
  for (int s = 0; s < desc->S; ++s) {
    p.src[s] = reinterpret_cast<const float4*>(sources_rgba[s]);
    SQLX_REQUIRE(p.src[s], "source %d is NULL", s);
    SQLX_REQUIRE((reinterpret_cast<uintptr_t>(p.src[s]) & 15) == 0, "source %d is not 16-byte aligned", s);
  }
  p.K = K; p.invK = inv_K; p.T = T; p.identity = identity;
  p.partial = partial;
  p.coef = (desc->flags & SQLX_NO_SSIM) ? nullptr : ssim_coef;
  p.ns = ns;
  p.ms_T_stride = T_stride; p.ms_partial_stride = partial_stride; p.ms_coef_stride = coef_stride;
  for (int i = 0; i < ns; ++i) {
    SQLX_REQUIRE(depth_up[i] && argmin[i] && (!automask || noise[i]), "scale %d: NULL pointer", i);
    p.ms_depth_up[i] = depth_up[i];
    p.ms_noise[i] = automask ? noise[i] : nullptr;
    p.ms_argmin[i] = argmin[i];
  }
  p.depth_lr = depth_up[0];   // never read: every scale has its upsampled plane
  const int r = (desc->flags & SQLX_NO_SSIM) ? 0 : desc->ssim_radius;
  return r == 3 ? dispatch_photo_fwd3<3>(p, ctas, st)
                : (r == 1 ? dispatch_photo_fwd3<1>(p, ctas, st) : dispatch_photo_fwd3<0>(p, ctas, st));
}

int photo_bwd3_launch(const sqlx_photo_desc* desc, const float* depth_lr, const float* depth_up, const float* target,
                      const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                      const uint8_t* argmin, const float* ssim_coef, const float* g_loss, float scale,
                      float* d_depth_lr, float* g_up, float* q_up, int g_up_accumulate, float* dP, cudaStream_t st,
                      const float* const* ref_depths, float* const* d_ref_depths, const float* g_reg) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(depth_lr && target && sources_rgba && K && inv_K && T && argmin && g_loss && (d_depth_lr || g_up) && dP,
               "NULL pointer argument");
  const int r = (desc->flags & SQLX_NO_SSIM) ? 0 : desc->ssim_radius;
  SQLX_REQUIRE(r == 0 || ssim_coef, "the backward needs the SSIM coefficients exported by sqlx_photo_fwd");
  PhotoBwdParams p;
  p.d = *desc;
  p.depth_lr = depth_lr; p.depth_up = depth_up; p.target = target;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s)
    p.src[s] = s < desc->S ? reinterpret_cast<const float4*>(sources_rgba[s]) : nullptr;
  for (int s = 0; s < desc->S; ++s) {
    SQLX_REQUIRE(p.src[s], "source %d is NULL", s);
    SQLX_REQUIRE((reinterpret_cast<uintptr_t>(p.src[s]) & 15) == 0, "source %d is not 16-byte aligned", s);
  }
  p.K = K; p.invK = inv_K; p.T = T; p.argmin = argmin; p.coef = ssim_coef; p.g_loss = g_loss; p.scale = scale;
  p.d_depth_lr = d_depth_lr; p.g_up = g_up; p.q_up = q_up; p.g_up_accumulate = g_up_accumulate;
  p.dP = dP;
  p.ns = 0;
  p.g_reg = g_reg;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s) {
    p.ref[s] = (g_reg && ref_depths && s < desc->S) ? ref_depths[s] : nullptr;
    p.d_ref[s] = (g_reg && d_ref_depths && s < desc->S) ? d_ref_depths[s] : nullptr;
  }
  if (g_reg) {
    SQLX_REQUIRE(ref_depths && d_ref_depths, "the indoor variant needs the source frames' depth maps and their gradients");
    for (int s = 0; s < desc->S; ++s) SQLX_REQUIRE(p.ref[s] && p.d_ref[s], "ref_depths[%d] / d_ref_depths[%d] is NULL", s, s);
  }
  return r == 3 ? dispatch_photo_bwd3<3>(p, st) : (r == 1 ? dispatch_photo_bwd3<1>(p, st) : dispatch_photo_bwd3<0>(p, st));
}

int photo_bwd3_ms_launch(const sqlx_photo_desc* desc, int ns, const float* const* depth_up, const float* target,
                         const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                         size_t T_stride, const uint8_t* const* argmin, const float* ssim_coef, size_t coef_stride,
                         const float* g_loss, float scale, float* const* g_up, float* const* q_up, unsigned accumulate_mask,
                         float* dP, size_t dP_stride, cudaStream_t st) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(ns >= 1 && ns <= SQLX_MAX_SCALES, "num_scales %d outside 1..%d", ns, SQLX_MAX_SCALES);
  SQLX_REQUIRE(depth_up && target && sources_rgba && K && inv_K && T && argmin && g_loss && g_up && dP, "NULL pointer argument");
  const int r = (desc->flags & SQLX_NO_SSIM) ? 0 : desc->ssim_radius;
  SQLX_REQUIRE(r == 0 || ssim_coef, "the backward needs the SSIM coefficients exported by the forward");
  SQLX_REQUIRE((long long)desc->B * ns <= 65535, "batch x scales exceeds the grid's z extent");
  PhotoBwdParams p;
  memset(&p, 0, sizeof(p));
  p.d = *desc;
  p.target = target;
  for (int s = 0; s < desc->S; ++s) {
    p.src[s] = reinterpret_cast<const float4*>(sources_rgba[s]);
    SQLX_REQUIRE(p.src[s], "source %d is NULL", s);
    SQLX_REQUIRE((reinterpret_cast<uintptr_t>(p.src[s]) & 15) == 0, "source %d is not 16-byte aligned", s);
  }
  p.K = K; p.invK = inv_K; p.T = T; p.coef = ssim_coef; p.g_loss = g_loss; p.scale = scale; p.dP = dP;
  p.ns = ns;
  p.ms_accumulate = accumulate_mask;
  p.ms_T_stride = T_stride; p.ms_coef_stride = coef_stride; p.ms_dP_stride = dP_stride;
  for (int i = 0; i < ns; ++i) {
    SQLX_REQUIRE(depth_up[i] && argmin[i] && g_up[i], "scale %d: NULL pointer", i);
    p.ms_depth_up[i] = depth_up[i];
    p.ms_argmin[i] = argmin[i];
    p.ms_g_up[i] = g_up[i];
    p.ms_q_up[i] = q_up ? q_up[i] : nullptr;
  }
  p.depth_lr = depth_up[0];   // never read: every scale has its upsampled plane
  p.g_up = g_up[0];
  return r == 3 ? dispatch_photo_bwd3<3>(p, st) : (r == 1 ? dispatch_photo_bwd3<1>(p, st) : dispatch_photo_bwd3<0>(p, st));
}

size_t photo_max_ctas(const sqlx_photo_desc* d) { return fwd_ctas(d); }
}  // namespace sqlx

extern "C" int sqlx_photo_fwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                              const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                              const float* identity, const float* noise, float* loss_sum, uint8_t* argmin,
                              float* ssim_coef, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(loss_sum, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sqlx_photo_workspace_bytes(desc), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  int ctas = 0;
  if (int e = photo_fwd3_launch(desc, depth_lr, nullptr, target, sources_rgba, K, inv_K, T, identity, noise, partial, &ctas, argmin,
                                ssim_coef, st, nullptr, nullptr))
    return e;
  finalize_sum3_kernel<<<1, 256, 0, st>>>(partial, ctas, loss_sum);
  return check_launch("finalize_sum_kernel");
}

extern "C" int sqlx_photo_bwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                              const float* const* sources_rgba, const float* K, const float* inv_K, const float* T,
                              const uint8_t* argmin, const float* ssim_coef, const float* g_loss, float scale,
                              float* d_depth_lr, float* d_T, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(d_depth_lr && d_T, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sqlx_photo_workspace_bytes(desc), "workspace too small");
  // dP accumulators live after the forward partial sums in the workspace
  float* dP = reinterpret_cast<float*>(workspace) + fwd_ctas(desc);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)desc->B * desc->S * 12, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(dP)");
  if (int e = photo_bwd3_launch(desc, depth_lr, nullptr, target, sources_rgba, K, inv_K, T, argmin, ssim_coef, g_loss, scale,
                                d_depth_lr, nullptr, nullptr, 0, dP, st, nullptr, nullptr, nullptr))
    return e;
  const int n = desc->B * desc->S * 16;
  dT_from_dP3_kernel<<<ceil_div(n, 128), 128, 0, st>>>(K, dP, desc->B, desc->S, d_T);
  return check_launch("dT_from_dP_kernel");
}

/* Indoor variant: see include/sqlx.h */
extern "C" int sqlx_photo_occ_fwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                                  const float* const* sources_rgba, const float* const* ref_depths, const float* K,
                                  const float* inv_K, const float* T, const float* identity, const float* noise,
                                  float* sums, uint8_t* argmin, float* ssim_coef, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(sums && ref_depths, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sqlx_photo_occ_workspace_bytes(desc), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  float* partial_reg = partial + fwd_ctas(desc);
  int ctas = 0;
  if (int e = photo_fwd3_launch(desc, depth_lr, nullptr, target, sources_rgba, K, inv_K, T, identity, noise, partial, &ctas,
                                argmin, ssim_coef, st, ref_depths, partial_reg))
    return e;
  finalize_sum3_kernel<<<1, 256, 0, st>>>(partial, ctas, sums);
  finalize_sum3_kernel<<<1, 256, 0, st>>>(partial_reg, ctas, sums + 1);
  return check_launch("finalize_sum_kernel");
}

extern "C" int sqlx_photo_occ_bwd(const sqlx_photo_desc* desc, const float* depth_lr, const float* target,
                                  const float* const* sources_rgba, const float* const* ref_depths, const float* K,
                                  const float* inv_K, const float* T, const uint8_t* argmin, const float* ssim_coef,
                                  const float* g_sums, float scale, float* d_depth_lr, float* d_T,
                                  float* const* d_ref_depths, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_desc(desc)) return e;
  SQLX_REQUIRE(d_depth_lr && d_T && g_sums, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sqlx_photo_occ_workspace_bytes(desc), "workspace too small");
  float* dP = reinterpret_cast<float*>(workspace) + 2 * fwd_ctas(desc);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)desc->B * desc->S * 12, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(dP)");
  if (int e = photo_bwd3_launch(desc, depth_lr, nullptr, target, sources_rgba, K, inv_K, T, argmin, ssim_coef, g_sums, scale,
                                d_depth_lr, nullptr, nullptr, 0, dP, st, ref_depths, d_ref_depths, g_sums + 1))
    return e;
  const int n = desc->B * desc->S * 16;
  dT_from_dP3_kernel<<<ceil_div(n, 128), 128, 0, st>>>(K, dP, desc->B, desc->S, d_T);
  return check_launch("dT_from_dP_kernel");
}

namespace {
template <int R>
int launch_identity3(const IdentParams& p, cudaStream_t st) {
  using C = Ident3Cfg<R, 32, 32, 256>;
  auto kern = identity3_kernel<R, 32, 32, 256>;
  if (int e = ensure_dyn_smem(kern, C::smem_bytes)) return e;
  dim3 grid(ceil_div(p.W, 32), ceil_div(p.H, 32), p.B);
  ProfScope prof("identity_loss_kernel", st);
  kern<<<grid, 256, C::smem_bytes, st>>>(p);
  return check_launch("identity3_kernel");
}
}  // namespace

extern "C" int sqlx_identity_losses_fwd(const float* target, const float* const* sources, int S, int B, int H, int W,
                                        int ssim_radius, float w_ssim, float w_l1, int no_ssim, float* identity,
                                        void* stream) {
  SQLX_REQUIRE(target && sources && identity, "NULL pointer argument");
  SQLX_REQUIRE(S >= 1 && S <= SQLX_MAX_SOURCES && B > 0 && H > 0 && W > 0, "bad shape");
  const int r = no_ssim ? 0 : ssim_radius;
  SQLX_REQUIRE(r == 0 || r == 1 || r == 3, "ssim_radius must be 1 or 3");
  SQLX_REQUIRE(H > 2 * r && W > 2 * r, "image smaller than the SSIM window");
  SQLX_REQUIRE((long long)H * W < (1ll << 30), "frame too large for 32-bit pixel offsets");
  IdentParams p;
  p.B = B; p.H = H; p.W = W; p.S = S; p.w_ssim = w_ssim; p.w_l1 = w_l1; p.target = target; p.out = identity;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s) p.src[s] = s < S ? sources[s] : nullptr;
  for (int s = 0; s < S; ++s) SQLX_REQUIRE(p.src[s], "source %d is NULL", s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return r == 3 ? launch_identity3<3>(p, st) : (r == 1 ? launch_identity3<1>(p, st) : launch_identity3<0>(p, st));
}
