// Scale-invariant log loss of the supervised fine-tuning path:
//   finetune/loss.py:29-42   SILogLoss.forward(input, target, mask, interpolate)
//     input  = F.interpolate(input, target.shape[-2:], mode="bilinear", align_corners=True)
//     g      = log(input[mask]) - log(target[mask])
//     loss   = 10 * sqrt(var(g) + 0.15 * mean(g)^2)            (torch.var: unbiased)
// One pass over the target-resolution pixels: the upsample, the mask gather and the two reductions are fused
// (the reference materialises the upsampled map, two boolean-mask gathers and three reductions); fixed-order
// double-precision partial sums, finalised by the last block.  Backward: one pass, atomics into the low-res map.
#include "common.cuh"

namespace sqlx {

// source index rule of F.interpolate(mode="bilinear", align_corners=True): src = dst * (in - 1) / (out - 1)
__device__ __forceinline__ UpTap up_tap_ac(int dst, float scale, int in_size) {
  const float src = scale * (float)dst;
  int i0 = (int)src;
  i0 = i0 > in_size - 1 ? in_size - 1 : i0;
  UpTap t;
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  t.l1 = src - (float)i0;
  t.l0 = 1.f - t.l1;
  return t;
}

constexpr int kSilogBlocks = 4 * kNumSMs;

// partial [blocks][3] doubles: sum g, sum g^2, count.  saved [4] floats: mean, n, Dg, loss
__global__ void silog_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                 const uint8_t* __restrict__ mask, int B, int h, int w, int H, int W, float vf,
                                 double* __restrict__ partial, unsigned int* __restrict__ counter,
                                 float* __restrict__ loss, float* __restrict__ saved) {
  __shared__ double sh[3][256];
  __shared__ int is_last;
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  const int rows = B * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / H, v = row - b * H;
    const UpTap ty = up_tap_ac(v, sy, h);
    const float* r0 = pred + ((size_t)b * h + ty.i0) * w;
    const float* r1 = pred + ((size_t)b * h + ty.i1) * w;
    const size_t o = (size_t)row * W;
    float a1 = 0.f, a2 = 0.f, ac = 0.f;     // per-row float partials, promoted to double once per row
    for (int u = threadIdx.x; u < W; u += blockDim.x) {
      if (mask && !mask[o + u]) continue;
      const UpTap tx = up_tap_ac(u, sx, w);
      const float p = ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1)) +
                      ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));
      const float g = logf(p) - logf(__ldg(gt + o + u));
      a1 += g; a2 = fmaf(g, g, a2); ac += 1.f;
    }
    s1 += (double)a1; s2 += (double)a2; cnt += (double)ac;
  }
  sh[0][threadIdx.x] = s1; sh[1][threadIdx.x] = s2; sh[2][threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3 + 0] = sh[0][0]; partial[blockIdx.x * 3 + 1] = sh[1][0]; partial[blockIdx.x * 3 + 2] = sh[2][0];
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    t0 += __ldcg(partial + i * 3); t1 += __ldcg(partial + i * 3 + 1); t2 += __ldcg(partial + i * 3 + 2);
  }
  sh[0][threadIdx.x] = t0; sh[1][threadIdx.x] = t1; sh[2][threadIdx.x] = t2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = sh[2][0], mean = sh[0][0] / n;
    const double var = (sh[1][0] - sh[0][0] * mean) / (n - 1.0);     // unbiased, as torch.var
    const double Dg = var + (double)vf * mean * mean;
    const double l = 10.0 * sqrt(Dg);
    loss[0] = (float)l;
    saved[0] = (float)mean; saved[1] = (float)n; saved[2] = (float)Dg; saved[3] = (float)l;
    *counter = 0u;
  }
}

// d loss / d g_i = 10 / (2 sqrt(Dg)) * ( 2 (g_i - mean) / (n - 1) + vf * 2 mean / n );  d g_i / d pred_up = 1 / pred_up
__global__ void silog_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                 const uint8_t* __restrict__ mask, int B, int h, int w, int H, int W, float vf,
                                 const float* __restrict__ saved, const float* __restrict__ g_loss,
                                 float* __restrict__ d_pred) {
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const float mean = saved[0], n = saved[1], Dg = saved[2];
  const float k = __ldg(g_loss) * 10.f * 0.5f * rsqrtf(Dg);
  const float c1 = k * 2.f / (n - 1.f), c2 = k * vf * 2.f * mean / n;
  const int rows = B * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / H, v = row - b * H;
    const UpTap ty = up_tap_ac(v, sy, h);
    const float* r0 = pred + ((size_t)b * h + ty.i0) * w;
    const float* r1 = pred + ((size_t)b * h + ty.i1) * w;
    float* o0 = d_pred + ((size_t)b * h + ty.i0) * w;
    float* o1 = d_pred + ((size_t)b * h + ty.i1) * w;
    const size_t o = (size_t)row * W;
    for (int u = threadIdx.x; u < W; u += blockDim.x) {
      if (mask && !mask[o + u]) continue;
      const UpTap tx = up_tap_ac(u, sx, w);
      const float p = ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1)) +
                      ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));
      const float g = logf(p) - logf(__ldg(gt + o + u));
      const float gp = (c1 * (g - mean) + c2) / p;
      atomicAdd(o0 + tx.i0, gp * ty.l0 * tx.l0);
      atomicAdd(o0 + tx.i1, gp * ty.l0 * tx.l1);
      atomicAdd(o1 + tx.i0, gp * ty.l1 * tx.l0);
      atomicAdd(o1 + tx.i1, gp * ty.l1 * tx.l1);
    }
  }
}

}  // namespace sqlx

using namespace sqlx;

extern "C" size_t sqlx_silog_workspace_bytes(void) { return sizeof(double) * 3 * kSilogBlocks + 256; }

extern "C" int sqlx_silog_fwd(const float* pred, const float* gt, const uint8_t* mask, int B, int h, int w, int H, int W,
                              float variance_focus, float* loss, float* saved, void* workspace, size_t workspace_bytes,
                              void* stream) {
  SQLX_REQUIRE(pred && gt && loss && saved && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "non-positive shape");
  SQLX_REQUIRE(workspace_bytes >= sqlx_silog_workspace_bytes(), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  if (cudaMemsetAsync(counter, 0, 256, st) != cudaSuccess) return check_launch("cudaMemsetAsync(silog counter)");
  const int blocks = B * H < kSilogBlocks ? B * H : kSilogBlocks;
  ProfScope prof("silog_fwd_kernel", st);
  silog_fwd_kernel<<<blocks, 256, 0, st>>>(pred, gt, mask, B, h, w, H, W, variance_focus, partial, counter, loss, saved);
  return check_launch("silog_fwd_kernel");
}

extern "C" int sqlx_silog_bwd(const float* pred, const float* gt, const uint8_t* mask, int B, int h, int w, int H, int W,
                              float variance_focus, const float* saved, const float* g_loss, float* d_pred,
                              void* stream) {
  SQLX_REQUIRE(pred && gt && saved && g_loss && d_pred, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "non-positive shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(d_pred, 0, sizeof(float) * (size_t)B * h * w, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(d_pred)");
  const int blocks = B * H < kSilogBlocks ? B * H : kSilogBlocks;
  ProfScope prof("silog_bwd_kernel", st);
  silog_bwd_kernel<<<blocks, 256, 0, st>>>(pred, gt, mask, B, h, w, H, W, variance_focus, saved, g_loss, d_pred);
  return check_launch("silog_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// Flip test-time augmentation blend: evaluate_depth_config.py:51-59 batch_post_process_disparity
// fused with the flip of the second pass (evaluate_depth_config.py:131,157): one kernel, no host round trip.
//   l [N,h,w] prediction of the frames, r [N,h,w] prediction of the horizontally flipped frames
//   r_is_flipped = 1: r still is in the flipped frame (as the network returned it) and is read mirrored
//   out = r_mask * l + l_mask * r' + (1 - l_mask - r_mask) * (l + r') / 2,  l_mask(u) = 1 - clip(20 (u/(w-1) - .05), 0, 1)
// ------------------------------------------------------------------------------------------------
namespace sqlx {
__global__ void postprocess_disp_kernel(const float* __restrict__ l, const float* __restrict__ r, int rows, int w,
                                        int r_is_flipped, float* __restrict__ out) {
  const float inv = w > 1 ? 1.f / (float)(w - 1) : 0.f;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const size_t o = (size_t)row * w;
    for (int u = threadIdx.x; u < w; u += blockDim.x) {
      const float lm = 1.f - fminf(fmaxf(20.f * ((float)u * inv - 0.05f), 0.f), 1.f);
      const float rm = 1.f - fminf(fmaxf(20.f * ((float)(w - 1 - u) * inv - 0.05f), 0.f), 1.f);
      const float a = l[o + u], b = r[o + (r_is_flipped ? w - 1 - u : u)];
      out[o + u] = rm * a + lm * b + (1.f - lm - rm) * (0.5f * (a + b));
    }
  }
}
}  // namespace sqlx

extern "C" int sqlx_postprocess_disparity(const float* l_disp, const float* r_disp, int N, int h, int w, int r_is_flipped,
                                          float* out, void* stream) {
  SQLX_REQUIRE(l_disp && r_disp && out, "NULL pointer argument");
  SQLX_REQUIRE(N > 0 && h > 0 && w > 0, "non-positive shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = N * h;
  postprocess_disp_kernel<<<rows < 8 * kNumSMs ? rows : 8 * kNumSMs, 128, 0, st>>>(l_disp, r_disp, rows, w, r_is_flipped, out);
  return check_launch("postprocess_disp_kernel");
}
