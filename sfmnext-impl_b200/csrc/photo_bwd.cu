// Smoothness forward / backward (layers.py:267-280, trainer.py:533-542) and the module-level SSIM backward
// (autograd of layers.py:31-46).  The fused photometric backward lives in photo_v3.cu.
#include "photo_tile.cuh"

#include <stdlib.h>

namespace sqlx {

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
  if (int e = ensure_dyn_smem(kern, smem)) return e;
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
