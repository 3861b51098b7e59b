// Photometric forward kernels:
//   (the fused forward / backward kernels live in photo_v3.cu)
//   reproj_loss_kernel      reprojection loss of two given images (identity losses, compute_reprojection_loss)
//   ssim_map_kernel         full SSIM map (module-level layers.SSIM drop-in)
//   warp_kernel             materialise depth_up / sample grid / warped colour (what Trainer.log reads)
//   depth_stats_kernel      per-sample mean(d_up), mean(1/d_up)
// Reference lines: layers.py:13-46,186-258; trainer.py:386-453,474-532.
#include "photo_tile.cuh"

namespace sqlx {

template <int R, int TH, int TW, int NT>
struct FwdCfg {
  static constexpr int PH = TH + 2 * R, PW = TW + 2 * R;
  static constexpr int LD = ((PW + 3) & ~3) + 2;   // even; 42 for a 32-wide tile: conflict-free 8-byte row accesses
  static constexpr int PLANE = PH * LD;
  static constexpr int HB = PH * TW;
  static constexpr int PPT = (TH * TW) / NT;
  static_assert((TH * TW) % NT == 0, "tile must be a multiple of the block size");
};

// Deterministic final reduction of per-CTA partial sums (one block; double accumulation).
__global__ void finalize_sum_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
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

// per-sample variant: partial laid out [B][per_b]; out[b*stride+col]
__global__ void finalize_rows_kernel(const float* __restrict__ partial, int per_b, int ncol, float* __restrict__ out,
                                     float scale) {
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
  if (threadIdx.x == 0) out[b * ncol + col] = (float)(sh[0] * (double)scale);
}

// ------------------------------------------------------------------------------------------------
// reprojection loss of two given images / SSIM map
// ------------------------------------------------------------------------------------------------
template <int R, int TH, int TW, int NT, bool MAP>
__global__ void __launch_bounds__(NT) reproj_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                         int C_, int H, int W, float w_ssim, float w_l1,
                                                         float* __restrict__ out, size_t out_bstride) {
  using C = FwdCfg<R, TH, TW, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);
  float* ys = xs + C::PLANE;
  float* hb = ys + C::PLANE;
  const int b = blockIdx.z;
  const int v0 = blockIdx.y * TH, u0 = blockIdx.x * TW;
  const size_t plane = (size_t)H * W;
  int prow[C::PPT], pcol[C::PPT];
  bool pin[C::PPT];
  float ssim_acc[C::PPT], l1_acc[C::PPT];
#pragma unroll
  for (int k = 0; k < C::PPT; ++k) {
    prow[k] = (threadIdx.x / TW) * C::PPT + k; pcol[k] = threadIdx.x % TW;
    pin[k] = (v0 + prow[k] < H) && (u0 + pcol[k] < W);
    ssim_acc[k] = 0.f; l1_acc[k] = 0.f;
  }
  for (int c = 0; c < C_; ++c) {
    stage_plane<C::PH, C::PW, C::LD>(pred + ((size_t)b * C_ + c) * plane, H, W, v0 - R, u0 - R, xs);
    stage_plane<C::PH, C::PW, C::LD>(target + ((size_t)b * C_ + c) * plane, H, W, v0 - R, u0 - R, ys);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < C::PPT; ++k) {
      const int o = (prow[k] + R) * C::LD + pcol[k] + R;
      l1_acc[k] += fabsf(ys[o] - xs[o]);
    }
    if (R > 0) {
      hpass_blocked<(R > 0 ? R : 1), C::PH, TW, C::LD, TW, true>(xs, ys, hb, hb + C::HB, hb + 2 * C::HB);
      hpass_blocked<(R > 0 ? R : 1), C::PH, TW, C::LD, TW, false>(nullptr, ys, hb + 3 * C::HB, hb + 4 * C::HB, nullptr);
      __syncthreads();
      float Sxv[C::PPT], Sxxv[C::PPT], Sxyv[C::PPT], Syv[C::PPT], Syyv[C::PPT];
      vsum_multi<R, TW, C::PPT>(hb, prow[0], pcol[0], Sxv);
      vsum_multi<R, TW, C::PPT>(hb + C::HB, prow[0], pcol[0], Sxxv);
      vsum_multi<R, TW, C::PPT>(hb + 2 * C::HB, prow[0], pcol[0], Sxyv);
      vsum_multi<R, TW, C::PPT>(hb + 3 * C::HB, prow[0], pcol[0], Syv);
      vsum_multi<R, TW, C::PPT>(hb + 4 * C::HB, prow[0], pcol[0], Syyv);
#pragma unroll
      for (int k = 0; k < C::PPT; ++k) {
        const float v = ssim_value(make_stats<R>(Sxv[k], Syv[k], Sxxv[k], Syyv[k], Sxyv[k]));
        if (MAP) {
          if (pin[k]) out[((size_t)b * C_ + c) * plane + (size_t)(v0 + prow[k]) * W + (u0 + pcol[k])] = v;
        } else {
          ssim_acc[k] += v;
        }
      }
    }
    __syncthreads();
  }
  if (!MAP) {
#pragma unroll
    for (int k = 0; k < C::PPT; ++k) {
      if (!pin[k]) continue;
      float rho;
      if (R > 0) rho = w_ssim * (ssim_acc[k] / (float)C_) + w_l1 * (l1_acc[k] / (float)C_);
      else rho = l1_acc[k] / (float)C_;
      out[(size_t)b * out_bstride + (size_t)(v0 + prow[k]) * W + (u0 + pcol[k])] = rho;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// warp only
// ------------------------------------------------------------------------------------------------
__global__ void warp_kernel(const float* __restrict__ depth_lr, const float* __restrict__ source,
                            const float* __restrict__ K, const float* __restrict__ invK, const float* __restrict__ T,
                            int T_stride, int h, int w, int H, int W, float eps, float* __restrict__ depth_up,
                            float* __restrict__ sample, float* __restrict__ color) {
  __shared__ Camera cam;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) load_camera(K + b * 16, invK + b * 16, T + (size_t)b * T_stride, cam);
  __syncthreads();
  const size_t plane = (size_t)H * W;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < H * W; idx += gridDim.x * blockDim.x) {
    const int v = idx / W, u = idx - v * W;
    const float d = upsample_at(depth_lr + (size_t)b * h * w, h, w, v, u, sy, sx);
    if (depth_up) depth_up[(size_t)b * plane + idx] = d;
    if (!sample && !color) continue;
    const Sample sp = project_pixel(cam, (float)u, (float)v, d, H, W, eps);
    if (sample) {
      sample[((size_t)b * plane + idx) * 2 + 0] = sp.gx;
      sample[((size_t)b * plane + idx) * 2 + 1] = sp.gy;
    }
    if (color) {
      const Taps t = make_taps(sp.ix, sp.iy, H, W);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* p = source + ((size_t)b * 3 + c) * plane;
        color[((size_t)b * 3 + c) * plane + idx] =
            __ldg(p + t.o00) * t.w00 + __ldg(p + t.o01) * t.w01 + __ldg(p + t.o10) * t.w10 + __ldg(p + t.o11) * t.w11;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depth statistics
// ------------------------------------------------------------------------------------------------
constexpr int kStatsBlocksPerSample = 32;

__global__ void depth_stats_kernel(const float* __restrict__ depth_lr, int h, int w, int H, int W,
                                   float* __restrict__ partial /*[B][blocks][2]*/) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float* lr = depth_lr + (size_t)b * h * w;
  float s0 = 0.f, s1 = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < H * W; idx += gridDim.x * blockDim.x) {
    const int v = idx / W, u = idx - v * W;
    const float d = upsample_at(lr, h, w, v, u, sy, sx);
    s0 += d;
    s1 += 1.f / d;
  }
  const float t0 = block_sum(s0, red);
  const float t1 = block_sum(s1, red);
  if (threadIdx.x == 0) {
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 0] = t0;
    partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = t1;
  }
}

// d_lr[i,j] += sum over hi-res pixels of tap weight * (g0/N + g1/N * (-1/d^2))
__global__ void depth_stats_bwd_kernel(const float* __restrict__ depth_lr, int h, int w, int H, int W,
                                       const float* __restrict__ g_stats, float* __restrict__ d_lr) {
  const int b = blockIdx.y;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float* lr = depth_lr + (size_t)b * h * w;
  float* out = d_lr + (size_t)b * h * w;
  const float invN = 1.f / ((float)H * (float)W);
  const float g0 = g_stats[b * 2 + 0] * invN, g1 = g_stats[b * 2 + 1] * invN;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < H * W; idx += gridDim.x * blockDim.x) {
    const int v = idx / W, u = idx - v * W;
    const UpTap ty = up_tap(v, sy, h), tx = up_tap(u, sx, w);
    const float d = ty.l0 * (tx.l0 * lr[ty.i0 * w + tx.i0] + tx.l1 * lr[ty.i0 * w + tx.i1]) +
                    ty.l1 * (tx.l0 * lr[ty.i1 * w + tx.i0] + tx.l1 * lr[ty.i1 * w + tx.i1]);
    const float g = g0 - g1 / (d * d);
    atomicAdd(out + ty.i0 * w + tx.i0, g * ty.l0 * tx.l0);
    atomicAdd(out + ty.i0 * w + tx.i1, g * ty.l0 * tx.l1);
    atomicAdd(out + ty.i1 * w + tx.i0, g * ty.l1 * tx.l0);
    atomicAdd(out + ty.i1 * w + tx.i1, g * ty.l1 * tx.l1);
  }
}

}  // namespace sqlx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace sqlx;

namespace {
constexpr int kTH = 16, kTW = 32, kNT = 256;

template <int R, bool MAP>
int launch_reproj(const float* pred, const float* target, int B, int C_, int H, int W, float w_ssim, float w_l1,
                  float* out, cudaStream_t st, size_t out_bstride = 0) {
  if (out_bstride == 0) out_bstride = (size_t)H * W;
  using C = FwdCfg<R, kTH, kTW, kNT>;
  auto kern = reproj_loss_kernel<R, kTH, kTW, kNT, MAP>;
  const size_t smem = sizeof(float) * (2 * C::PLANE + 5 * C::HB);
  if (int e = ensure_dyn_smem(kern, smem)) return e;
  dim3 grid(ceil_div(W, kTW), ceil_div(H, kTH), B);
  ProfScope prof("reproj_loss_kernel", st);
  kern<<<grid, kNT, smem, st>>>(pred, target, C_, H, W, w_ssim, w_l1, out, out_bstride);
  return check_launch("reproj_loss_kernel");
}
}  // namespace

extern "C" int sqlx_reprojection_loss_fwd(const float* pred, const float* target, int B, int H, int W,
                                          int ssim_radius, float w_ssim, float w_l1, int no_ssim, float* out,
                                          void* stream) {
  SQLX_REQUIRE(pred && target && out, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 0 && W > 0, "non-positive shape");
  const int r = no_ssim ? 0 : ssim_radius;
  SQLX_REQUIRE(r == 0 || r == 1 || r == 3, "ssim_radius must be 1 or 3");
  SQLX_REQUIRE(H > 2 * r && W > 2 * r, "image smaller than the SSIM window");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (r == 3) return launch_reproj<3, false>(pred, target, B, 3, H, W, w_ssim, w_l1, out, st);
  if (r == 1) return launch_reproj<1, false>(pred, target, B, 3, H, W, w_ssim, w_l1, out, st);
  return launch_reproj<0, false>(pred, target, B, 3, H, W, w_ssim, w_l1, out, st);
}

extern "C" int sqlx_ssim_fwd(const float* x, const float* y, int B, int C, int H, int W, int ssim_radius, float* out,
                             void* stream) {
  SQLX_REQUIRE(x && y && out, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "non-positive shape");
  SQLX_REQUIRE(ssim_radius == 1 || ssim_radius == 3, "ssim_radius must be 1 or 3");
  SQLX_REQUIRE(H > 2 * ssim_radius && W > 2 * ssim_radius, "image smaller than the SSIM window");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (ssim_radius == 3) return launch_reproj<3, true>(x, y, B, C, H, W, 0.f, 0.f, out, st);
  return launch_reproj<1, true>(x, y, B, C, H, W, 0.f, 0.f, out, st);
}

extern "C" int sqlx_warp_fwd(const float* depth_lr, const float* source, const float* K, const float* inv_K,
                             const float* T, int T_stride, int B, int h, int w, int H, int W, float eps,
                             float* depth_up, float* sample, float* color, void* stream) {
  SQLX_REQUIRE(depth_lr && K && inv_K && T, "NULL pointer argument");
  SQLX_REQUIRE(!color || source, "color output needs a source image");
  SQLX_REQUIRE(B > 0 && H > 1 && W > 1 && h > 0 && w > 0 && h <= H && w <= W, "bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(min(ceil_div(H * W, 256), 4 * kNumSMs), B);
  warp_kernel<<<grid, 256, 0, st>>>(depth_lr, source, K, inv_K, T, T_stride, h, w, H, W, eps, depth_up, sample, color);
  return check_launch("warp_kernel");
}

extern "C" size_t sqlx_depth_stats_workspace_bytes(int B, int H, int W) {
  (void)H; (void)W;
  return sizeof(float) * (size_t)B * kStatsBlocksPerSample * 2;
}

extern "C" int sqlx_depth_stats_fwd(const float* depth_lr, int B, int h, int w, int H, int W, float* stats,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(depth_lr && stats && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= H && w <= W, "bad shape");
  SQLX_REQUIRE(workspace_bytes >= sqlx_depth_stats_workspace_bytes(B, H, W), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  ProfScope prof("depth_stats_kernel", st);
  depth_stats_kernel<<<dim3(kStatsBlocksPerSample, B), 256, 0, st>>>(depth_lr, h, w, H, W, partial);
  if (int e = check_launch("depth_stats_kernel")) return e;
  finalize_rows_kernel<<<dim3(B, 2), 256, 0, st>>>(partial, kStatsBlocksPerSample, 2, stats, 1.f / ((float)H * (float)W));
  return check_launch("finalize_rows_kernel");
}

extern "C" int sqlx_depth_stats_bwd(const float* depth_lr, int B, int h, int w, int H, int W, const float* g_stats,
                                    float* d_depth_lr, void* stream) {
  SQLX_REQUIRE(depth_lr && g_stats && d_depth_lr, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= H && w <= W, "bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope prof("depth_stats_bwd_kernel", st);
  depth_stats_bwd_kernel<<<dim3(kStatsBlocksPerSample, B), 256, 0, st>>>(depth_lr, h, w, H, W, g_stats, d_depth_lr);
  return check_launch("depth_stats_bwd_kernel");
}
