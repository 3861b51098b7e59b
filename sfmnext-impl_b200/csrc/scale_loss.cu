// One loss scale of Trainer.generate_images_pred + compute_losses (trainer.py:386-439, 455-549) as ONE C-ABI
// call forward and ONE backward: the host-side chaining of the kernels (depth statistics -> pose matrices ->
// fused photometric kernel -> smoothness -> scalar) lives here instead of in ~70 tiny PyTorch ops per scale.
//
//   loss_s = mean_{b,v,u} min(identity + noise, reprojection) + (smooth_weight) * smooth(disp / mean(disp), color_s)
#include "pose.cuh"

namespace sqlx {

// T[b,s] = pose matrix of source s (scaled by mean inverse depth when rescale) or the fixed transform
struct PoseSources {
  const float* axisangle[SQLX_MAX_SOURCES];     // [B,3] or NULL -> fixed
  const float* translation[SQLX_MAX_SOURCES];   // [B,3]
  const float* fixed_T[SQLX_MAX_SOURCES];       // [B,4,4] when axisangle is NULL
  float* d_axisangle[SQLX_MAX_SOURCES];         // backward outputs (may be NULL)
  float* d_translation[SQLX_MAX_SOURCES];
  uint32_t invert_mask;
};

__global__ void pose_multi_fwd_kernel(PoseSources ps, const float* __restrict__ stats /*[B,2] or NULL*/, int B, int S,
                                      float* __restrict__ T /*[B,S,4,4]*/) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int b = idx / S, s = idx - b * S;
  float M[16];
  if (ps.axisangle[s]) {
    const float a[3] = {ps.axisangle[s][b * 3], ps.axisangle[s][b * 3 + 1], ps.axisangle[s][b * 3 + 2]};
    const float t[3] = {ps.translation[s][b * 3], ps.translation[s][b * 3 + 1], ps.translation[s][b * 3 + 2]};
    pose_eval<float>(a, t, stats ? stats[b * 2 + 1] : 1.f, (ps.invert_mask >> s) & 1u, M);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = ps.fixed_T[s][b * 16 + i];
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) T[(size_t)idx * 16 + i] = M[i];
}

// one thread per (sample, source, input j): j = 0..2 axisangle, 3..5 translation, 6 scale.
// d_scale contributions are summed over sources into g_stats[b][1] (g_stats[b][0] = 0).
__global__ void pose_multi_bwd_kernel(PoseSources ps, const float* __restrict__ stats, int B, int S,
                                      const float* __restrict__ dT /*[B,S,4,4]*/, float* __restrict__ g_stats /*[B,2]*/) {
  __shared__ float sh_scale[256];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // blockDim = 7 * S * (samples per block)
  const int per_b = 7 * S;
  const int b = idx / per_b, rem = idx - b * per_b, s = rem / 7, j = rem - s * 7;
  float g = 0.f;
  const bool live = b < B && ps.axisangle[s] != nullptr;
  if (live) {
    Dual a[3], t[3], sc;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      a[i] = {ps.axisangle[s][b * 3 + i], j == i ? 1.f : 0.f};
      t[i] = {ps.translation[s][b * 3 + i], j == 3 + i ? 1.f : 0.f};
    }
    sc = {stats ? stats[b * 2 + 1] : 1.f, j == 6 ? 1.f : 0.f};
    Dual M[16];
    pose_eval<Dual>(a, t, sc, (ps.invert_mask >> s) & 1u, M);
#pragma unroll
    for (int i = 0; i < 12; ++i) g += dT[((size_t)b * S + s) * 16 + i] * M[i].d;
    if (j < 3) {
      if (ps.d_axisangle[s]) ps.d_axisangle[s][b * 3 + j] = g;
    } else if (j < 6) {
      if (ps.d_translation[s]) ps.d_translation[s][b * 3 + (j - 3)] = g;
    }
  }
  sh_scale[threadIdx.x] = (live && j == 6) ? g : 0.f;
  __syncthreads();
  if (b < B && rem == 0 && g_stats) {   // first thread of each sample sums the per-source scale gradients
    float tot = 0.f;
    for (int k = 0; k < per_b; ++k) tot += sh_scale[threadIdx.x + k];
    g_stats[b * 2 + 0] = 0.f;
    g_stats[b * 2 + 1] = stats ? tot : 0.f;
  }
}

// loss = loss_sum / (B*H*W) + w * sum_b [ sx_b / (Nx) + sy_b / (Ny) ] / (sd_b / N + 1e-7)
__global__ void scale_finalize_kernel(const float* __restrict__ loss_sum, const float* __restrict__ sums /*[B,3]*/, int B,
                                      float inv_bhw, float w, float Nx, float Ny, float N, float* __restrict__ loss) {
  if (threadIdx.x != 0) return;
  double acc = 0.0;
  for (int b = 0; b < B; ++b) {
    const float inv = 1.f / (sums[b * 3 + 2] / N + 1e-7f);
    acc += (double)(sums[b * 3 + 0] * inv) / Nx + (double)(sums[b * 3 + 1] * inv) / Ny;
  }
  loss[0] = loss_sum[0] * inv_bhw + w * (float)acc;
}

// upstream gradients of the smoothness sums given g = d(total)/d(loss_s)
__global__ void scale_bwd_prep_kernel(const float* __restrict__ g_loss, const float* __restrict__ sums, int B, float w,
                                      float Nx, float Ny, float N, float* __restrict__ g_sums /*[B,3]*/) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float g = g_loss[0] * w;
  const float inv = 1.f / (sums[b * 3 + 2] / N + 1e-7f);
  g_sums[b * 3 + 0] = g * inv / Nx;
  g_sums[b * 3 + 1] = g * inv / Ny;
  g_sums[b * 3 + 2] = -g * (sums[b * 3 + 0] / Nx + sums[b * 3 + 1] / Ny) * inv * inv / N;
}

}  // namespace sqlx

using namespace sqlx;

namespace {
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ScaleWs {
  float* T;         // [B,S,16]
  float* stats;     // [B,2]
  float* sums;      // [B,3]
  float* loss_sum;  // [1]
  float* coef;      // [B,S,3,3,H,W] SSIM derivative coefficients (absent with --no_ssim)
  float* g_sums;    // [B,3]
  float* g_stats;   // [B,2]
  float* dT;        // [B,S,16]
  uint8_t* scratch; // kernels' own workspaces
  size_t scratch_bytes;
};

size_t scratch_bytes_for(const sqlx_scale_desc* d) {
  size_t m = sqlx_photo_workspace_bytes(&d->photo);
  const size_t a = sqlx_depth_stats_workspace_bytes(d->photo.B, d->photo.H, d->photo.W);
  const size_t b = sqlx_smooth_workspace_bytes(d->photo.B, d->Hc, d->Wc);
  if (a > m) m = a;
  if (b > m) m = b;
  return align256(m);
}

ScaleWs carve_ws(const sqlx_scale_desc* d, void* saved, void* workspace) {
  const int B = d->photo.B, S = d->photo.S;
  ScaleWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(saved);       // persists from forward to backward
  w.T = reinterpret_cast<float*>(p); p += align256(sizeof(float) * B * S * 16);
  w.stats = reinterpret_cast<float*>(p); p += align256(sizeof(float) * B * 2);
  w.sums = reinterpret_cast<float*>(p); p += align256(sizeof(float) * B * 3);
  w.loss_sum = reinterpret_cast<float*>(p); p += 256;
  w.coef = (d->photo.flags & SQLX_NO_SSIM) ? nullptr : reinterpret_cast<float*>(p);
  uint8_t* q = reinterpret_cast<uint8_t*>(workspace);   // scratch
  w.g_sums = reinterpret_cast<float*>(q); q += align256(sizeof(float) * B * 3);
  w.g_stats = reinterpret_cast<float*>(q); q += align256(sizeof(float) * B * 2);
  w.dT = reinterpret_cast<float*>(q); q += align256(sizeof(float) * B * S * 16);
  w.scratch = q;
  w.scratch_bytes = scratch_bytes_for(d);
  return w;
}

int check_scale(const sqlx_scale_desc* d, const sqlx_pose_inputs* poses) {
  SQLX_REQUIRE(d && poses, "NULL descriptor");
  SQLX_REQUIRE(d->photo.S >= 1 && d->photo.S <= SQLX_MAX_SOURCES, "S=%d outside 1..%d", d->photo.S, SQLX_MAX_SOURCES);
  SQLX_REQUIRE(d->Hc > 1 && d->Wc > 1 && d->photo.h <= d->Hc && d->photo.w <= d->Wc, "bad colour-pyramid shape %dx%d", d->Hc, d->Wc);
  for (int s = 0; s < d->photo.S; ++s)
    SQLX_REQUIRE((poses->axisangle[s] && poses->translation[s]) || poses->fixed_T[s],
                 "source %d has neither (axisangle, translation) nor a fixed transform", s);
  return SQLX_OK;
}

PoseSources to_sources(const sqlx_pose_inputs* poses, int S, float* const* d_aa, float* const* d_tr) {
  PoseSources ps;
  for (int s = 0; s < SQLX_MAX_SOURCES; ++s) {
    ps.axisangle[s] = s < S ? poses->axisangle[s] : nullptr;
    ps.translation[s] = s < S ? poses->translation[s] : nullptr;
    ps.fixed_T[s] = s < S ? poses->fixed_T[s] : nullptr;
    ps.d_axisangle[s] = (d_aa && s < S) ? d_aa[s] : nullptr;
    ps.d_translation[s] = (d_tr && s < S) ? d_tr[s] : nullptr;
  }
  ps.invert_mask = poses->invert_mask;
  return ps;
}
}  // namespace

extern "C" size_t sqlx_scale_saved_bytes(const sqlx_scale_desc* d) {
  if (!d) return 0;
  const int B = d->photo.B, S = d->photo.S;
  const size_t coef = (d->photo.flags & SQLX_NO_SSIM) ? 0 : sizeof(float) * 9 * (size_t)B * S * d->photo.H * d->photo.W;
  return align256(sizeof(float) * B * S * 16) + align256(sizeof(float) * B * 2) + align256(sizeof(float) * B * 3) + 256 +
         align256(coef);
}

extern "C" size_t sqlx_scale_workspace_bytes(const sqlx_scale_desc* d) {
  if (!d) return 0;
  const int B = d->photo.B, S = d->photo.S;
  return align256(sizeof(float) * B * 3) + align256(sizeof(float) * B * 2) + align256(sizeof(float) * B * S * 16) +
         scratch_bytes_for(d) + 256;
}

extern "C" int sqlx_scale_loss_fwd(const sqlx_scale_desc* d, const float* depth_lr, const float* target,
                                   const float* const* sources, const float* color_s, const float* K, const float* inv_K,
                                   const sqlx_pose_inputs* poses, const float* identity, const float* noise, float* loss,
                                   uint8_t* argmin, void* saved, size_t saved_bytes, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (int e = check_scale(d, poses)) return e;
  SQLX_REQUIRE(depth_lr && target && sources && color_s && K && inv_K && loss && argmin && saved && workspace,
               "NULL pointer argument");
  SQLX_REQUIRE(saved_bytes >= sqlx_scale_saved_bytes(d) && workspace_bytes >= sqlx_scale_workspace_bytes(d),
               "saved / workspace buffer too small");
  const int B = d->photo.B, S = d->photo.S, H = d->photo.H, W = d->photo.W, h = d->photo.h, w = d->photo.w;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ScaleWs ws = carve_ws(d, saved, workspace);
  bool any_pose = false;
  for (int s = 0; s < S; ++s) any_pose |= poses->axisangle[s] != nullptr;
  const bool rescale = d->rescale_translation && any_pose;
  if (rescale) {   // trainer.py:417-418  mean inverse depth of the upsampled map
    if (int e = sqlx_depth_stats_fwd(depth_lr, B, h, w, H, W, ws.stats, ws.scratch, ws.scratch_bytes, stream)) return e;
  }
  pose_multi_fwd_kernel<<<ceil_div(B * S, 64), 64, 0, st>>>(to_sources(poses, S, nullptr, nullptr),
                                                             rescale ? ws.stats : nullptr, B, S, ws.T);
  if (int e = check_launch("pose_multi_fwd_kernel")) return e;
  if (int e = sqlx_photo_fwd(&d->photo, depth_lr, target, sources, K, inv_K, ws.T, identity, noise, ws.loss_sum, argmin,
                             ws.coef, ws.scratch, ws.scratch_bytes, stream))
    return e;
  if (int e = sqlx_smooth_fwd(depth_lr, color_s, B, h, w, d->Hc, d->Wc, ws.sums, ws.scratch, ws.scratch_bytes, stream))
    return e;
  const float N = (float)d->Hc * (float)d->Wc;
  scale_finalize_kernel<<<1, 32, 0, st>>>(ws.loss_sum, ws.sums, B, 1.f / ((float)B * H * W), d->smooth_weight,
                                          (float)B * d->Hc * (d->Wc - 1), (float)B * (d->Hc - 1) * d->Wc, N, loss);
  return check_launch("scale_finalize_kernel");
}

extern "C" int sqlx_scale_loss_bwd(const sqlx_scale_desc* d, const float* depth_lr, const float* target,
                                   const float* const* sources, const float* color_s, const float* K, const float* inv_K,
                                   const sqlx_pose_inputs* poses, const uint8_t* argmin, const float* g_loss,
                                   const void* saved, float* d_depth_lr, float* const* d_axisangle,
                                   float* const* d_translation, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_scale(d, poses)) return e;
  SQLX_REQUIRE(depth_lr && target && sources && color_s && K && inv_K && argmin && g_loss && saved && d_depth_lr &&
               workspace, "NULL pointer argument");
  SQLX_REQUIRE(workspace_bytes >= sqlx_scale_workspace_bytes(d), "workspace too small");
  const int B = d->photo.B, S = d->photo.S, H = d->photo.H, W = d->photo.W, h = d->photo.h, w = d->photo.w;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ScaleWs ws = carve_ws(d, const_cast<void*>(saved), workspace);
  bool any_pose = false;
  for (int s = 0; s < S; ++s) any_pose |= poses->axisangle[s] != nullptr;
  const bool rescale = d->rescale_translation && any_pose;
  const float N = (float)d->Hc * (float)d->Wc;
  scale_bwd_prep_kernel<<<ceil_div(B, 64), 64, 0, st>>>(g_loss, ws.sums, B, d->smooth_weight,
                                                         (float)B * d->Hc * (d->Wc - 1), (float)B * (d->Hc - 1) * d->Wc,
                                                         N, ws.g_sums);
  if (int e = check_launch("scale_bwd_prep_kernel")) return e;
  if (cudaMemsetAsync(d_depth_lr, 0, sizeof(float) * (size_t)B * h * w, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(d_depth_lr)");
  if (int e = sqlx_smooth_bwd(depth_lr, color_s, B, h, w, d->Hc, d->Wc, ws.g_sums, d_depth_lr, stream)) return e;
  if (int e = sqlx_photo_bwd(&d->photo, depth_lr, target, sources, K, inv_K, ws.T, argmin, ws.coef, g_loss,
                             1.f / ((float)B * H * W), d_depth_lr, ws.dT, ws.scratch, ws.scratch_bytes, stream))
    return e;
  if (any_pose) {
    const int per_b = 7 * S;
    const int spb = 252 / per_b;   // samples per block (blockDim <= 256 = shared array size)
    pose_multi_bwd_kernel<<<ceil_div(B, spb), spb * per_b, 0, st>>>(to_sources(poses, S, d_axisangle, d_translation),
                                                                    rescale ? ws.stats : nullptr, B, S, ws.dT,
                                                                    rescale ? ws.g_stats : nullptr);
    if (int e = check_launch("pose_multi_bwd_kernel")) return e;
    if (rescale) {
      if (int e = sqlx_depth_stats_bwd(depth_lr, B, h, w, H, W, ws.g_stats, d_depth_lr, stream)) return e;
    }
  }
  return SQLX_OK;
}

