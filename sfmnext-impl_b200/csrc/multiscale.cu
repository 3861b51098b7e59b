// All loss scales of Trainer.generate_images_pred + compute_losses (trainer.py:386-439, 455-549) as ONE forward
// and ONE backward library call.  The reference loops `for scale in self.opt.scales` in Python and issues ~70 tiny
// ATen kernels per scale; here the per-scale work that is not the fused photometric kernel is batched ACROSS scales:
//
//   forward    ms_stats_pose_kernel       mean inverse depth of every (scale, sample) (trainer.py:417-418) and, by the
//                                         last block of each (scale, sample), the S pose matrices (layers.py:75-150)
//              photo_fwd3_kernel x scales (photo_v3.cu)
//              ms_smooth_loss_kernel      edge-aware smoothness sums of every (scale, sample) (layers.py:267-280,
//                                         trainer.py:533-542); the last block reduces the photometric per-CTA
//                                         partials and writes loss/s and the total (trainer.py:532,543-549)
//   backward   ms_smooth_bwd_kernel       smoothness gradient per pixel (plain stores)
//              photo_bwd3_kernel x scales gradient wrt the upsampled depth per pixel (plain stores) and dP
//              ms_pose_bwd_kernel         dT = K^T dP -> d axisangle / d translation summed over scales, d(mean 1/d)
//              ms_upsample_adjoint_kernel gather-style adjoint of F.interpolate(bilinear, align_corners=False)
//                                         (+ the mean-inverse-depth term): deterministic, overwrites, no atomics
//
// 12 launches (+2 memsets) per step for 4 scales instead of 64.  Every reduction has a fixed summation order.
#include "photo_v3.h"

#include <stdlib.h>
#include "pose.cuh"

namespace sqlx {

constexpr int kMsBlocks = 96;   // blocks per (scale, sample) of the batched reduction kernels

struct MsShapes {
  int ns, B, S, H, W;
  int h[SQLX_MAX_SCALES], w[SQLX_MAX_SCALES];
  int Hc[SQLX_MAX_SCALES], Wc[SQLX_MAX_SCALES];
  float smooth_weight[SQLX_MAX_SCALES];
  const float* depth[SQLX_MAX_SCALES];
  const float* color[SQLX_MAX_SCALES];
  float* d_up[SQLX_MAX_SCALES];        // [B,H,W] upsampled depth, written once by ms_stats_pose_kernel and read by every
                                       // later kernel of the scale (one load instead of four loads and a blend)
  const float* dmap[SQLX_MAX_SCALES];  // [B,Hc,Wc] the map the smoothness term differentiates: depth itself or d_up
};

struct MsPose {
  const float* axisangle[SQLX_MAX_SOURCES];     // [B,3] or NULL -> fixed
  const float* translation[SQLX_MAX_SOURCES];   // [B,3]
  const float* fixed_T[SQLX_MAX_SOURCES];       // [B,4,4] when axisangle is NULL
  float* d_axisangle[SQLX_MAX_SOURCES];         // backward outputs (may be NULL)
  float* d_translation[SQLX_MAX_SOURCES];
  uint32_t invert_mask;
};

// Pixel loop of the batched per-sample kernels: block x of gridDim.x takes rows x, x + gridDim.x, ...; a thread keeps
// the (row-independent) column taps of its <= KC columns in registers and evaluates the row taps once per row, so a
// pixel costs four L1 loads and a blend instead of ~60 instructions of index arithmetic.
constexpr int kKC = 4;   // columns per thread held in registers (frames up to 4 * blockDim wide; wider ones loop)

__device__ __forceinline__ float blend4(const float* __restrict__ lr, int w, const UpTap& ty, const UpTap& tx) {
  const float* r0 = lr + ty.i0 * w;
  const float* r1 = lr + ty.i1 * w;
  return ty.l0 * (tx.l0 * __ldg(r0 + tx.i0) + tx.l1 * __ldg(r0 + tx.i1)) +
         ty.l1 * (tx.l0 * __ldg(r1 + tx.i0) + tx.l1 * __ldg(r1 + tx.i1));   // same expression as upsample_at
}

// Integer upsampling factor (1 or even): one work item = (low-resolution cell (i, j), one row of the F x F block of
// pixels whose first bilinear tap is that cell; the whole block for F <= 2) -- four loads per item, the tap weights
// (r + .5) / F are exactly the ones up_tap() derives from src = (v + .5) / F - .5 for these factors.  Writes d_up and
// returns this thread's sum of 1 / d_up.  F = 0: runtime factor `fr` (same arithmetic, loops not unrolled).
template <int F>
__device__ __forceinline__ float upsample_blocks(const float* __restrict__ lr, float* __restrict__ dup, int h, int w,
                                                 int H, int W, int fr = 0) {
  const int f = F > 0 ? F : fr;
  const float rf = 1.f / (float)f;
  const int half = f >> 1;
  float s1 = 0.f;
  // A block walks whole rows of cells; a thread owns a cell column j and, when the row is narrower than the block, one of
  // `nrg` interleaved subsets of the cell's F pixel rows: ONE index division per thread (not two per item as in the first
  // version, 83 instructions per output pixel), and the F pixels of a row segment leave next to the neighbouring
  // thread's.
  const int nrg = w >= (int)blockDim.x ? 1 : min(f, (int)blockDim.x / w);
  const int rg = nrg == 1 ? 0 : (int)threadIdx.x / w;
  const int jstep = nrg == 1 ? (int)blockDim.x : w;
  const int j0 = (int)threadIdx.x - rg * w;
  if (rg >= nrg) return s1;
  for (int i = blockIdx.x; i < h; i += gridDim.x) {
    const int i1 = min(i + 1, h - 1);
    const float* r0 = lr + (size_t)i * w;
    const float* r1 = lr + (size_t)i1 * w;
    for (int j = j0; j < w; j += jstep) {
      const int j1 = min(j + 1, w - 1);
      const float a = __ldg(r0 + j), bq = __ldg(r0 + j1), c = __ldg(r1 + j), dq = __ldg(r1 + j1);
      if (F > 1 && i > 0 && j > 0 && i < h - 1 && j < w - 1) {
        // interior cell: compile-time column weights, no range checks
        float* blk = dup + (size_t)(F * i + half) * W + (F * j + half);
        for (int r = rg; r < F; r += nrg) {
          const float wy = ((float)r + 0.5f) * rf;
          const float top = 1.f - wy;
          float* row = blk + (size_t)r * W;
#pragma unroll
          for (int q = 0; q < F; ++q) {
            const float wx = ((float)q + 0.5f) * rf;
            const float d = top * ((1.f - wx) * a + wx * bq) + wy * ((1.f - wx) * c + wx * dq);   // as upsample_at
            row[q] = d;
            s1 += __fdividef(1.f, d);
          }
        }
        continue;
      }
      // border cells (and the runtime-factor path): the first cell row / column also owns the clamped pixels before it
      // (row-group 0 takes the extra rows)
      const int c_lo = (j == 0) ? -half : 0;
      for (int r = (i == 0 && rg == 0) ? -half : rg; r < f; r += (r < 0 ? 1 : nrg)) {
        const int v = f * i + half + r;
        if (v >= H) break;
        const float wy = f == 1 ? 0.f : fmaxf(((float)r + 0.5f) * rf, 0.f);
        float* row = dup + (size_t)v * W;
        for (int q = c_lo; q < f; ++q) {
          const int u = f * j + half + q;
          if (u >= W) break;
          const float wx = f == 1 ? 0.f : fmaxf(((float)q + 0.5f) * rf, 0.f);
          const float d = (1.f - wy) * ((1.f - wx) * a + wx * bq) + wy * ((1.f - wx) * c + wx * dq);
          row[u] = d;
          s1 += __fdividef(1.f, d);
        }
      }
    }
  }
  return s1;
}

// ------------------------------------------------------------------------------------------------
// forward: depth statistics + pose matrices
// ------------------------------------------------------------------------------------------------
// grid (kMsBlocks, B, ns).  Also materialises the upsampled depth d_up [B,H,W] of every scale.
// partial [ns][B][kMsBlocks], counter [ns][B] (zero on entry, left zero), stats [ns][B] = mean_{HxW} 1/d_up,
// T [ns][B][S][16]
__global__ void ms_stats_pose_kernel(MsShapes sh, MsPose ps, int rescale, float* __restrict__ partial,
                                     unsigned int* __restrict__ counter, float* __restrict__ stats,
                                     float* __restrict__ T) {
  __shared__ float red[32];
  __shared__ int is_last;
  const int sc = blockIdx.z, b = blockIdx.y;
  const int B = sh.B, S = sh.S, H = sh.H, W = sh.W;
  float mean_inv = 1.f;
  {
    const int h = sh.h[sc], w = sh.w[sc];
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const float* lr = sh.depth[sc] + (size_t)b * h * w;
    float* dup = sh.d_up[sc] + (size_t)b * H * W;
    float s1 = 0.f;
    const int f = H / h;
    if (H == f * h && W == f * w && (f == 1 || (f & 1) == 0)) {
      // Integer upsampling factor (1, 2, 4, ...): one thread per low-resolution cell (i, j) produces the f x f block of
      // pixels whose first bilinear tap is that cell -- four loads per BLOCK, three FMAs per pixel.  The tap weights
      // (r + .5) / f are exactly the ones up_tap() derives from src = (v + .5) / f - .5 for these factors.
      if (f == 2) s1 = upsample_blocks<2>(lr, dup, h, w, H, W);
      else if (f == 4) s1 = upsample_blocks<4>(lr, dup, h, w, H, W);
      else if (f == 8) s1 = upsample_blocks<8>(lr, dup, h, w, H, W);
      else if (f == 16) s1 = upsample_blocks<16>(lr, dup, h, w, H, W);
      else if (f == 1) s1 = upsample_blocks<1>(lr, dup, h, w, H, W);
      else s1 = upsample_blocks<0>(lr, dup, h, w, H, W, f);
    } else {
      for (int u0 = 0; u0 < W; u0 += kKC * blockDim.x) {
        UpTap tx[kKC];
#pragma unroll
        for (int k = 0; k < kKC; ++k) tx[k] = up_tap(min(u0 + k * (int)blockDim.x + (int)threadIdx.x, W - 1), sx, w);
        for (int v = blockIdx.x; v < H; v += gridDim.x) {
          const UpTap ty = up_tap(v, sy, h);
#pragma unroll
          for (int k = 0; k < kKC; ++k) {
            const int u = u0 + k * (int)blockDim.x + (int)threadIdx.x;
            if (u < W) {
              const float d = blend4(lr, w, ty, tx[k]);
              dup[v * W + u] = d;
              s1 += __fdividef(1.f, d);
            }
          }
        }
      }
    }
    const float t1 = block_sum(s1, red);
    if (threadIdx.x == 0) {
      partial[((size_t)sc * B + b) * kMsBlocks + blockIdx.x] = t1;
      __threadfence();
      is_last = atomicAdd(&counter[sc * B + b], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 32) {   // fixed-order sum of the block partials
      const volatile float* pp = partial + ((size_t)sc * B + b) * kMsBlocks;
      float v = 0.f;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) v += pp[i];
      v = warp_sum(v);
      if (threadIdx.x == 0) {
        red[0] = v / ((float)H * (float)W);
        stats[sc * B + b] = red[0];
        counter[sc * B + b] = 0u;
      }
    }
    __syncthreads();
    mean_inv = rescale ? red[0] : 1.f;
  }
  if (threadIdx.x < S) {
    const int s = threadIdx.x;
    float M[16];
    if (ps.axisangle[s]) {
      const float a[3] = {ps.axisangle[s][b * 3], ps.axisangle[s][b * 3 + 1], ps.axisangle[s][b * 3 + 2]};
      const float t[3] = {ps.translation[s][b * 3], ps.translation[s][b * 3 + 1], ps.translation[s][b * 3 + 2]};
      pose_eval<float>(a, t, mean_inv, (ps.invert_mask >> s) & 1u, M);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) M[i] = ps.fixed_T[s][b * 16 + i];
    }
    float* out = T + (((size_t)sc * B + b) * S + s) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = M[i];
  }
}

// ------------------------------------------------------------------------------------------------
// forward: smoothness sums + loss assembly
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float absdiff_mean3(const float* __restrict__ col, size_t plane, int o0, int o1) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s += fabsf(__ldg(col + c * plane + o0) - __ldg(col + c * plane + o1));
  return s / 3.f;
}

// grid (kMsBlocks, B, ns).  spartial [ns][B][kMsBlocks][3]; sums [ns][B][3] = {sum |dx d| e^-|dx I|, same in y, sum d};
// photo_partial [ns][max_ctas] with n_ctas[sc] valid entries; loss [1 + ns] = {total, loss/0, loss/1, ...}.
__global__ void ms_smooth_loss_kernel(MsShapes sh, float* __restrict__ spartial, float* __restrict__ sums,
                                      const float* __restrict__ photo_partial, int max_ctas, int ctas,
                                      unsigned int* __restrict__ counter, float* __restrict__ loss) {
  __shared__ float red3[3][32];
  __shared__ double dred[256];
  __shared__ float lsum[SQLX_MAX_SCALES];
  extern __shared__ float sterm[];        // [ns * B]
  __shared__ int is_last;
  const int sc = blockIdx.z, b = blockIdx.y, B = sh.B;
  {
    const int Hc = sh.Hc[sc], Wc = sh.Wc[sc];
    const size_t plane = (size_t)Hc * Wc;
    const float* dm = sh.dmap[sc] + (size_t)b * plane;
    const float* col = sh.color[sc] + (size_t)b * 3 * plane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    // three pixels of a row per thread and trip: their 36 loads are issued together
    constexpr int KU = 3;
    for (int v = blockIdx.x; v < Hc; v += gridDim.x) {
      const bool down = v + 1 < Hc;
      for (int u0 = threadIdx.x; u0 < Wc; u0 += KU * blockDim.x) {
        float d[KU], dr[KU], dd[KU], c[KU][3], cr[KU][3], cd[KU][3];
        bool on[KU], right[KU];
#pragma unroll
        for (int k = 0; k < KU; ++k) {
          const int u = u0 + k * blockDim.x;
          on[k] = u < Wc;
          right[k] = u + 1 < Wc;
          const int idx = v * Wc + (on[k] ? u : 0);
          d[k] = on[k] ? dm[idx] : 0.f;
          dr[k] = right[k] ? dm[idx + 1] : 0.f;
          dd[k] = (on[k] && down) ? dm[idx + Wc] : 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            c[k][ch] = on[k] ? __ldg(col + ch * plane + idx) : 0.f;
            cr[k][ch] = right[k] ? __ldg(col + ch * plane + idx + 1) : 0.f;
            cd[k][ch] = (on[k] && down) ? __ldg(col + ch * plane + idx + Wc) : 0.f;
          }
        }
#pragma unroll
        for (int k = 0; k < KU; ++k) {
          if (!on[k]) continue;
          s2 += d[k];
          if (right[k]) {
            const float e = (fabsf(c[k][0] - cr[k][0]) + fabsf(c[k][1] - cr[k][1]) + fabsf(c[k][2] - cr[k][2])) * (1.f / 3.f);
            s0 += fabsf(d[k] - dr[k]) * __expf(-e);
          }
          if (down) {
            const float e = (fabsf(c[k][0] - cd[k][0]) + fabsf(c[k][1] - cd[k][1]) + fabsf(c[k][2] - cd[k][2])) * (1.f / 3.f);
            s1 += fabsf(d[k] - dd[k]) * __expf(-e);
          }
        }
      }
    }
    // one shared stage for the three sums (two barriers instead of six)
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
      red3[0][threadIdx.x >> 5] = s0; red3[1][threadIdx.x >> 5] = s1; red3[2][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    if (threadIdx.x == 0) {
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t0 += red3[0][i]; t1 += red3[1][i]; t2 += red3[2][i]; }
    }
    if (threadIdx.x == 0) {
      float* o = spartial + (((size_t)sc * B + b) * kMsBlocks + blockIdx.x) * 3;
      o[0] = t0; o[1] = t1; o[2] = t2;
      __threadfence();
      // two-level arrival: 1) the blocks of this (scale, sample) -- the counters the statistics kernel left zero --
      // 2) the (scale, sample) groups of the launch.  (One counter for all ~1500 blocks of the launch serialised their
      // atomics on one address and left every reduction to a single block.)
      is_last = atomicAdd(&counter[64 + sc * B + b], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block of this (scale, sample): its three sums, fixed order (lane j owns partials j, j + 32, ...; shuffle tree)
    if (threadIdx.x < 96) {
      const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const float* q = spartial + ((size_t)sc * B + b) * kMsBlocks * 3 + k;
      float v = 0.f;
      for (int j = lane; j < (int)gridDim.x; j += 32) v += __ldcg(q + j * 3);
      v = warp_sum(v);
      if (lane == 0) sums[((size_t)sc * B + b) * 3 + k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      counter[64 + sc * B + b] = 0u;
      __threadfence();
      is_last = atomicAdd(counter, 1u) == gridDim.y * gridDim.z - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
  }
  // ---- last block of the launch: every reduction in a fixed order, spread over the block's threads
  const int ns = sh.ns;
  // photometric per-CTA partials of every scale: strided double sums, one tree for all scales
  double acc[SQLX_MAX_SCALES];
#pragma unroll
  for (int s = 0; s < SQLX_MAX_SCALES; ++s) {
    acc[s] = 0.0;
    if (s < ns) {
      const float* pp = photo_partial + (size_t)s * max_ctas;
      for (int base = 0; base < ctas; base += 8 * (int)blockDim.x) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = base + k * (int)blockDim.x + (int)threadIdx.x;
          v[k] = i < ctas ? pp[i] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[s] += (double)v[k];
      }
    }
  }
  __syncthreads();   // sums[] visible to the block
#pragma unroll
  for (int s = 0; s < SQLX_MAX_SCALES; ++s) {
    if (s >= ns) break;
    dred[threadIdx.x] = acc[s];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) lsum[s] = (float)dred[0];
    __syncthreads();
  }
  // smoothness term of every (scale, sample) in parallel, then a fixed-order sum per scale
  for (int i = threadIdx.x; i < ns * B; i += blockDim.x) {
    const int s = i / B;
    const float Hc = (float)sh.Hc[s], Wc = (float)sh.Wc[s];
    const float rN = 1.f / (Hc * Wc), rNx = 1.f / ((float)B * Hc * (Wc - 1.f)), rNy = 1.f / ((float)B * (Hc - 1.f) * Wc);
    const float* q = sums + (size_t)i * 3;
    const float q0 = __ldcg(q), q1 = __ldcg(q + 1), q2 = __ldcg(q + 2);      // written by other blocks of this launch
    const float inv = 1.f / (q2 * rN + 1e-7f);
    sterm[i] = (q0 * inv) * rNx + (q1 * inv) * rNy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int s = 0; s < ns; ++s) {
      double sm = 0.0;
      for (int bb = 0; bb < B; ++bb) sm += (double)sterm[s * B + bb];
      const float ls = lsum[s] * (1.f / ((float)B * (float)sh.H * (float)sh.W)) + sh.smooth_weight[s] * (float)sm;
      loss[1 + s] = ls;
      total += ls;
    }
    loss[0] = total / (float)ns;
    *counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct MsGrads {
  float* g_up[SQLX_MAX_SCALES];       // [B,H,W]   gradient wrt the upsampled depth (photometric; + smoothness when it upsamples)
  float* g_direct[SQLX_MAX_SCALES];   // [B,h,w]   smoothness gradient when the map is used at its own resolution, else NULL
  float* d_depth[SQLX_MAX_SCALES];    // [B,h,w]   outputs
};

// Smoothness gradient per pixel of the colour-resolution map.  grid (blocks, B, ns)
__global__ void ms_smooth_bwd_kernel(MsShapes sh, MsGrads g, const float* __restrict__ g_loss,
                                     const float* __restrict__ sums) {
  const int sc = blockIdx.z, b = blockIdx.y, B = sh.B;
  const int Hc = sh.Hc[sc], Wc = sh.Wc[sc];
  const size_t plane = (size_t)Hc * Wc;
  const float* dm = sh.dmap[sc] + (size_t)b * plane;
  const float* col = sh.color[sc] + (size_t)b * 3 * plane;
  float* out = (g.g_direct[sc] ? g.g_direct[sc] : g.g_up[sc]) + (size_t)b * plane;
  // d loss / d {sum_x, sum_y, sum_d} of this sample  (loss = total / ns ; loss_s = ... + w_s * smooth)
  const float N = (float)Hc * (float)Wc, Nx = (float)B * Hc * (Wc - 1.f), Ny = (float)B * (Hc - 1.f) * Wc;
  const float* q = sums + ((size_t)sc * B + b) * 3;
  const float gw = __ldg(g_loss) * (1.f / (float)sh.ns) * sh.smooth_weight[sc];
  const float inv = 1.f / (q[2] / N + 1e-7f);
  const float gxs = gw * inv / Nx, gys = gw * inv / Ny;
  const float gds = -gw * (q[0] / Nx + q[1] / Ny) * inv * inv / N;
  for (int v = blockIdx.x; v < Hc; v += gridDim.x) {
    for (int u = threadIdx.x; u < Wc; u += blockDim.x) {
      const int idx = v * Wc + u;
      const float d = dm[idx];
      const float c0 = __ldg(col + idx), c1 = __ldg(col + plane + idx), c2 = __ldg(col + 2 * plane + idx);
      float gg = gds;
#define SQLX_EDGE(o) ((fabsf(c0 - __ldg(col + (o))) + fabsf(c1 - __ldg(col + plane + (o))) + \
                       fabsf(c2 - __ldg(col + 2 * plane + (o)))) * (1.f / 3.f))
      if (u + 1 < Wc) {
        const float diff = d - dm[idx + 1];
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        gg += gxs * sg * __expf(-SQLX_EDGE(idx + 1));
      }
      if (u > 0) {
        const float diff = dm[idx - 1] - d;
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        gg -= gxs * sg * __expf(-SQLX_EDGE(idx - 1));
      }
      if (v + 1 < Hc) {
        const float diff = d - dm[idx + Wc];
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        gg += gys * sg * __expf(-SQLX_EDGE(idx + Wc));
      }
      if (v > 0) {
        const float diff = dm[idx - Wc] - d;
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        gg -= gys * sg * __expf(-SQLX_EDGE(idx - Wc));
      }
#undef SQLX_EDGE
      out[idx] = gg;
    }
  }
}

// One thread per (sample, source, input j): j = 0..2 axisangle, 3..5 translation, 6 scale (mean inverse depth).
// dT = K^T dP per scale; d axisangle / d translation are summed over the scales in scale order;
// g_stats [ns][B] = sum over sources of d loss / d mean_inv of that scale.
__global__ void ms_pose_bwd_kernel(MsShapes sh, MsPose ps, int rescale, const float* __restrict__ K,
                                   const float* __restrict__ stats, const float* __restrict__ dP /*[ns][B][S][12]*/,
                                   float* __restrict__ g_stats) {
  __shared__ float sh_scale[256];
  const int B = sh.B, S = sh.S, ns = sh.ns;
  const int per_b = 7 * S;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // blockDim = per_b * (samples per block)
  const int b = idx / per_b, rem = idx - b * per_b, s = rem / 7, j = rem - s * 7;
  const bool live = b < B && ps.axisangle[s] != nullptr;
  float gsum = 0.f;
  for (int sc = 0; sc < ns; ++sc) {
    float g = 0.f;
    if (live) {
      Dual a[3], t[3], scl;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a[i] = {ps.axisangle[s][b * 3 + i], j == i ? 1.f : 0.f};
        t[i] = {ps.translation[s][b * 3 + i], j == 3 + i ? 1.f : 0.f};
      }
      scl = {rescale ? stats[sc * B + b] : 1.f, j == 6 ? 1.f : 0.f};
      Dual M[16];
      pose_eval<Dual>(a, t, scl, (ps.invert_mask >> s) & 1u, M);
      const float* dp = dP + (((size_t)sc * B + b) * S + s) * 12;
      const float* Kb = K + b * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i)        // dT[i][c] = sum_k K[k][i] dP[k][c]; rows 0..2 of T carry derivatives
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (i == 3) continue;
          const float dT = Kb[0 * 4 + i] * dp[0 * 4 + c] + Kb[1 * 4 + i] * dp[1 * 4 + c] + Kb[2 * 4 + i] * dp[2 * 4 + c];
          g += dT * M[i * 4 + c].d;
        }
    }
    if (j < 6) gsum += g;
    // per-scale gradient of the mean inverse depth: summed over the sources of the sample by thread (s = 0, j = 6)
    sh_scale[threadIdx.x] = (live && j == 6) ? g : 0.f;
    __syncthreads();
    if (b < B && rem == 6 && g_stats) {
      float tot = 0.f;
      for (int k = 0; k < S; ++k) tot += sh_scale[threadIdx.x + 7 * k];
      g_stats[sc * B + b] = rescale ? tot : 0.f;
    }
    __syncthreads();
  }
  if (live && j < 3) { if (ps.d_axisangle[s]) ps.d_axisangle[s][b * 3 + j] = gsum; }
  else if (live && j < 6) { if (ps.d_translation[s]) ps.d_translation[s][b * 3 + (j - 3)] = gsum; }
}

// Gather-style adjoint of the bilinear upsampling, all scales in one launch (blockIdx.y = scale): `tpc` lanes per
// low-resolution cell, one footprint column per lane (column weight evaluated once), rows walked in a loop.
//   d_depth[cell] = sum_px wy(v) wx(u) * (g_up[px] - g_stat * q_up[px]) + g_direct[cell]
// q_up[px] = 1 / d_up(px)^2 is stored by photo_bwd3_kernel; g_stat = d loss / d mean(1/d_up) / (H W).
struct MsAdjoint {
  const float* q_up[SQLX_MAX_SCALES];
  int log2_tpc[SQLX_MAX_SCALES];
  int factor[SQLX_MAX_SCALES];   // integer upsampling factor (1 or even) or 0: generic path
  float sy[SQLX_MAX_SCALES], sx[SQLX_MAX_SCALES], ry[SQLX_MAX_SCALES], rx[SQLX_MAX_SCALES];
  float rHW;
};

// Even integer factor F, TPC lanes per cell: lane `sub` takes footprint columns sub, sub + TPC, ...; all trip counts
// are compile-time so the loads of a column are issued together.
template <int F, int TPC>
__device__ __forceinline__ float adjoint_even(const float* __restrict__ gp, const float* __restrict__ qp, int i, int j,
                                              int h, int w, int H, int W, int sub, float gs, bool rescale) {
  constexpr int half = F / 2;
  constexpr float rf = 1.f / (float)F;
  const int vb = F * i - half, ub = F * j - half;
  float acc = 0.f;
  if (i > 0 && i < h - 1 && j > 0 && j < w - 1) {
    // interior cell (all but the map's border ring): the whole footprint is inside the frame and every weight is the
    // compile-time tent value -- two loads and two FMAs per pixel, addresses are base + constant offsets
#pragma unroll
    for (int c = 0; c < 2 * F / TPC; ++c) {
      const int tu = sub + c * TPC;
      const float wx = tu < F ? ((float)tu + 0.5f) * rf : ((float)(2 * F - tu) - 0.5f) * rf;
      const float* gcol = gp + (size_t)vb * W + (ub + tu);
      const float* qcol = qp + (size_t)vb * W + (ub + tu);
      // rows in batches of (at most) 8: 16 loads in flight per thread keeps the kernel at 64 registers / 4 CTAs per SM
      // (with the whole 2F-row column in registers the launch ran 19 latency-bound waves at 2 CTAs per SM)
      constexpr int RB = 2 * F < 8 ? 2 * F : 8;
      float col = 0.f;
#pragma unroll
      for (int t0 = 0; t0 < 2 * F; t0 += RB) {
        float gv[RB], qv[RB];
#pragma unroll
        for (int t = 0; t < RB; ++t) {
          gv[t] = gcol[(t0 + t) * W];
          qv[t] = rescale ? qcol[(t0 + t) * W] : 0.f;
        }
#pragma unroll
        for (int t = 0; t < RB; ++t) {
          const int tv = t0 + t;
          const float wy = tv < F ? ((float)tv + 0.5f) * rf : ((float)(2 * F - tv) - 0.5f) * rf;
          col = fmaf(wy, fmaf(-gs, qv[t], gv[t]), col);
        }
      }
      acc = fmaf(wx, col, acc);
    }
    return acc;
  }
#pragma unroll
  for (int c = 0; c < 2 * F / TPC; ++c) {
    const int tu = sub + c * TPC;
    const int u = ub + tu;
    if (u < 0 || u >= W) continue;
    float wx = tu < F ? ((float)tu + 0.5f) * rf : ((float)(2 * F - tu) - 0.5f) * rf;
    if ((j == 0 && u < half) || (j == w - 1 && tu >= F)) wx = 1.f;
    const float* gcol = gp + u;
    const float* qcol = qp + u;
    float col = 0.f;
#pragma unroll 4
    for (int tv = 0; tv < 2 * F; ++tv) {      // border ring of the map only: not batched
      const int v = vb + tv;
      if (v < 0 || v >= H) continue;
      float wy = tv < F ? ((float)tv + 0.5f) * rf : ((float)(2 * F - tv) - 0.5f) * rf;
      if ((i == 0 && v < half) || (i == h - 1 && tv >= F)) wy = 1.f;
      const float gq = rescale ? qcol[v * W] : 0.f;
      col = fmaf(wy, fmaf(-gs, gq, gcol[v * W]), col);
    }
    acc = fmaf(wx, col, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(256, 4) ms_upsample_adjoint_kernel(MsShapes sh, MsGrads g, MsAdjoint ad, int rescale,
                                           const float* __restrict__ g_stats) {
  const int sc = blockIdx.y;
  const int B = sh.B, H = sh.H, W = sh.W, h = sh.h[sc], w = sh.w[sc];
  const int lt = ad.log2_tpc[sc], tpc = 1 << lt;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;      // < 2^31 (checked on the host)
  const unsigned ncell = (unsigned)B * h * w;
  if ((blockIdx.x * blockDim.x) >> lt >= ncell) return;            // whole block beyond this scale
  const unsigned cell = gid >> lt;
  const int sub = (int)(gid & (tpc - 1));
  const bool live = cell < ncell;
  float acc = 0.f;
  if (live) {
    const unsigned hw = (unsigned)h * w;
    const int b = (int)(cell / hw), rem = (int)(cell - (unsigned)b * hw);
    const int i = rem / w, j = rem - i * w;
    const float* gp = g.g_up[sc] + (size_t)b * H * W;
    const float* qp = ad.q_up[sc] + (size_t)b * H * W;
    const float gs = rescale ? g_stats[sc * B + b] * ad.rHW : 0.f;
    const int f = ad.factor[sc];
    if (f == 1) {                      // identity "upsampling": the footprint is the cell itself
      if (sub == 0) acc = fmaf(-gs, rescale ? qp[i * W + j] : 0.f, gp[i * W + j]);
    } else if (f == 2) {       // lanes per cell fixed by the host: 1, 2, 8, 32 for factors 2, 4, 8, 16
      acc = adjoint_even<2, 1>(gp, qp, i, j, h, w, H, W, sub, gs, rescale);
    } else if (f == 4) {
      acc = adjoint_even<4, 2>(gp, qp, i, j, h, w, H, W, sub, gs, rescale);
    } else if (f == 8) {
      acc = adjoint_even<8, 8>(gp, qp, i, j, h, w, H, W, sub, gs, rescale);
    } else if (f == 16) {
      acc = adjoint_even<16, 32>(gp, qp, i, j, h, w, H, W, sub, gs, rescale);
    } else if (f > 1) {
      // Even integer factor: the footprint is rows f*i - f/2 .. f*i + 3f/2 - 1 (same for columns) with the tent weights
      // (t + .5)/f, t < f, and (2f - t - .5)/f above; pixels clamped onto the first / last map row carry weight 1.
      // One footprint column per lane, rows in a loop.
      const int half = f >> 1;
      const float rf = 1.f / (float)f;
      const int vb = f * i - half, ub = f * j - half;
      for (int tu = sub; tu < 2 * f; tu += tpc) {
        const int u = ub + tu;
        if (u < 0 || u >= W) continue;
        float wx = tu < f ? ((float)tu + 0.5f) * rf : ((float)(2 * f - tu) - 0.5f) * rf;
        if ((j == 0 && u < half) || (j == w - 1 && tu >= f)) wx = 1.f;
        float col = 0.f;
        const float* gcol = gp + u;
        const float* qcol = qp + u;
#pragma unroll 4
        for (int tv = 0; tv < 2 * f; ++tv) {
          const int v = vb + tv;
          if (v < 0 || v >= H) continue;
          float wy = tv < f ? ((float)tv + 0.5f) * rf : ((float)(2 * f - tv) - 0.5f) * rf;
          if ((i == 0 && v < half) || (i == h - 1 && tv >= f)) wy = 1.f;
          float gg = gcol[v * W];
          if (rescale) gg = fmaf(-gs, qcol[v * W], gg);
          col = fmaf(wy, gg, col);
        }
        acc = fmaf(wx, col, acc);
      }
    } else {
      // generic scale: conservative footprint bounds (rows whose source coordinate sy*(v+.5)-.5 lies in [i-1, i+1),
      // with a margin), exact weights from up_tap (rows outside the true footprint evaluate to weight 0)
      const float sy = ad.sy[sc], sx = ad.sx[sc], ry = ad.ry[sc], rx = ad.rx[sc];   // h/H, w/W, H/h, W/w
      const int v_lo = max(0, (int)ceilf(((float)i - 0.5f) * ry - 0.5f - 1e-3f));
      const int v_hi = min(H - 1, (int)ceilf(((float)i + 1.5f) * ry - 0.5f + 1e-3f));
      const int u_lo = max(0, (int)ceilf(((float)j - 0.5f) * rx - 0.5f - 1e-3f));
      const int u_hi = min(W - 1, (int)ceilf(((float)j + 1.5f) * rx - 0.5f + 1e-3f));
      for (int u = u_lo + sub; u <= u_hi; u += tpc) {
        const UpTap tx = up_tap(u, sx, w);
        const float wx = (tx.i0 == j ? tx.l0 : 0.f) + (tx.i1 == j ? tx.l1 : 0.f);
        if (wx == 0.f) continue;
        float col = 0.f;
        const float* gcol = gp + u;
        const float* qcol = qp + u;
        for (int v = v_lo; v <= v_hi; ++v) {
          const UpTap ty = up_tap(v, sy, h);
          const float wy = (ty.i0 == i ? ty.l0 : 0.f) + (ty.i1 == i ? ty.l1 : 0.f);
          float gg = gcol[v * W];
          if (rescale) gg = fmaf(-gs, qcol[v * W], gg);
          col = fmaf(wy, gg, col);
        }
        acc = fmaf(wx, col, acc);
      }
    }
  }
  for (int o = tpc >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (live && sub == 0) {
    if (g.g_direct[sc]) acc += g.g_direct[sc][cell];
    g.d_depth[sc][cell] = acc;
  }
}

}  // namespace sqlx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace sqlx;

namespace {
size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct MsLayout {
  // saved (forward -> backward)
  size_t T, stats, sums, d_up, d_up_stride, coef, coef_stride, saved_total;
  // workspace
  size_t counters, partial, spartial, photo_partial, g_stats, dP, g_up, g_up_stride, q_up, g_direct, g_direct_stride, ws_total;
  int max_ctas;
};

MsLayout ms_layout(const sqlx_ms_desc* d) {
  MsLayout L;
  const int ns = d->num_scales, B = d->photo.B, S = d->photo.S;
  size_t o = 0;
  L.T = o; o += al256(sizeof(float) * (size_t)ns * B * S * 16);
  L.stats = o; o += al256(sizeof(float) * (size_t)ns * B);
  L.sums = o; o += al256(sizeof(float) * (size_t)ns * B * 3);
  L.d_up_stride = al256(sizeof(float) * (size_t)B * d->photo.H * d->photo.W);
  L.d_up = o; o += L.d_up_stride * ns;
  L.coef_stride = (d->photo.flags & SQLX_NO_SSIM) ? 0 : al256(sizeof(float) * 9 * (size_t)B * S * d->photo.H * d->photo.W);
  L.coef = o; o += L.coef_stride * ns;
  L.saved_total = o;
  sqlx_photo_desc pd = d->photo;
  L.max_ctas = (int)photo_max_ctas(&pd);
  o = 0;
  L.counters = o; o += 256 + al256(sizeof(unsigned int) * (size_t)ns * B);
  L.partial = o; o += al256(sizeof(float) * (size_t)ns * B * kMsBlocks);
  L.spartial = o; o += al256(sizeof(float) * (size_t)ns * B * kMsBlocks * 3);
  L.photo_partial = o; o += al256(sizeof(float) * (size_t)ns * L.max_ctas);
  L.g_stats = o; o += al256(sizeof(float) * (size_t)ns * B);
  L.dP = o; o += al256(sizeof(float) * (size_t)ns * B * S * 12);
  L.g_up_stride = al256(sizeof(float) * (size_t)B * d->photo.H * d->photo.W);
  L.g_up = o; o += L.g_up_stride * ns;
  L.q_up = o; o += L.g_up_stride * ns;
  size_t mx = 0;
  for (int s = 0; s < ns; ++s) {
    const size_t b = sizeof(float) * (size_t)B * d->h[s] * d->w[s];
    if (b > mx) mx = b;
  }
  L.g_direct_stride = al256(mx);
  L.g_direct = o; o += L.g_direct_stride * ns;
  L.ws_total = o;
  return L;
}

int check_ms(const sqlx_ms_desc* d, const sqlx_pose_inputs* poses) {
  SQLX_REQUIRE(d && poses, "NULL descriptor");
  SQLX_REQUIRE(d->num_scales >= 1 && d->num_scales <= SQLX_MAX_SCALES, "num_scales=%d outside 1..%d", d->num_scales,
               SQLX_MAX_SCALES);
  SQLX_REQUIRE(d->photo.B > 0 && d->photo.H > 1 && d->photo.W > 1, "bad frame shape");
  SQLX_REQUIRE(d->photo.S >= 1 && d->photo.S <= SQLX_MAX_SOURCES, "S=%d outside 1..%d", d->photo.S, SQLX_MAX_SOURCES);
  SQLX_REQUIRE(7 * d->photo.S <= 252, "too many sources");
  for (int s = 0; s < d->num_scales; ++s) {
    SQLX_REQUIRE(d->h[s] > 0 && d->w[s] > 0 && d->h[s] <= d->photo.H && d->w[s] <= d->photo.W,
                 "scale %d: depth map %dx%d does not fit the %dx%d frame", s, d->h[s], d->w[s], d->photo.H, d->photo.W);
    // trainer.py:533-534: the map is used as is when it has the colour image's shape, else upsampled to (H, W)
    SQLX_REQUIRE(d->Hc[s] > 1 && d->Wc[s] > 1 &&
                 ((d->Hc[s] == d->h[s] && d->Wc[s] == d->w[s]) || (d->Hc[s] == d->photo.H && d->Wc[s] == d->photo.W)),
                 "scale %d: colour image %dx%d matches neither the depth map nor the frame", s, d->Hc[s], d->Wc[s]);
  }
  for (int s = 0; s < d->photo.S; ++s)
    SQLX_REQUIRE((poses->axisangle[s] && poses->translation[s]) || poses->fixed_T[s],
                 "source %d has neither (axisangle, translation) nor a fixed transform", s);
  return SQLX_OK;
}

MsShapes make_shapes(const sqlx_ms_desc* d, const float* const* depth, const float* const* color, const MsLayout& L,
                     void* saved) {
  MsShapes sh;
  sh.ns = d->num_scales; sh.B = d->photo.B; sh.S = d->photo.S; sh.H = d->photo.H; sh.W = d->photo.W;
  for (int s = 0; s < SQLX_MAX_SCALES; ++s) {
    const bool on = s < d->num_scales;
    sh.h[s] = on ? d->h[s] : 0; sh.w[s] = on ? d->w[s] : 0;
    sh.Hc[s] = on ? d->Hc[s] : 0; sh.Wc[s] = on ? d->Wc[s] : 0;
    sh.smooth_weight[s] = on ? d->smooth_weight[s] : 0.f;
    sh.depth[s] = on ? depth[s] : nullptr;
    sh.color[s] = on ? color[s] : nullptr;
    sh.d_up[s] = on ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(saved) + L.d_up + L.d_up_stride * s) : nullptr;
    // trainer.py:533-534: the smoothness term takes the map as is when it has the colour image's shape
    sh.dmap[s] = !on ? nullptr : ((d->Hc[s] == d->h[s] && d->Wc[s] == d->w[s]) ? depth[s] : sh.d_up[s]);
  }
  return sh;
}

MsPose make_pose(const sqlx_pose_inputs* poses, int S, float* const* d_aa, float* const* d_tr) {
  MsPose ps;
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

// The one-launch multi-scale photometric kernels are instantiated and GPU-tested for the live 7x7 SSIM; the 3x3 and
// --no_ssim variants go through the per-scale launches.
bool ms_fusable(const sqlx_ms_desc* d) { return !(d->photo.flags & SQLX_NO_SSIM) && d->photo.ssim_radius == 3; }

// Blocks per (scale, sample) actually launched (<= kMsBlocks, which sizes the partial arrays): about 15 frame pixels
// per thread -- measured on a B200 at 192x640 x 4 scales x 12 samples: 96 blocks 36 + 40 us (statistics + smoothness
// kernels), 48: 32 + 34, 32: 30 + 31, 24: 29 + 33.
int ms_blocks(int H, int W) {
  int k = (int)(((long long)H * W + 256 * 15 - 1) / (256 * 15));
  if (k < 24) k = 24;
  return k > kMsBlocks ? kMsBlocks : k;
}
// the same, capped so that the (blocks, B, ns) grid of 256-thread blocks is ONE wave (8 blocks per SM): these kernels end
// with a block-level reduction and an arrival counter, and a partial second wave doubled their elapsed time
int ms_blocks_one_wave(int H, int W, int B, int ns) {
  int k = ms_blocks(H, W);
  const int cap = (8 * kNumSMs) / (B * ns > 0 ? B * ns : 1);
  if (k > cap) k = cap;
  return k < 4 ? 4 : k;
}

sqlx_photo_desc scale_photo_desc(const sqlx_ms_desc* d, int s) {
  sqlx_photo_desc pd = d->photo;
  pd.h = d->h[s]; pd.w = d->w[s];
  return pd;
}
}  // namespace

extern "C" size_t sqlx_ms_saved_bytes(const sqlx_ms_desc* d) {
  if (!d || d->num_scales < 1 || d->num_scales > SQLX_MAX_SCALES) return 0;
  return ms_layout(d).saved_total;
}

extern "C" size_t sqlx_ms_workspace_bytes(const sqlx_ms_desc* d) {
  if (!d || d->num_scales < 1 || d->num_scales > SQLX_MAX_SCALES) return 0;
  return ms_layout(d).ws_total;
}

extern "C" int sqlx_ms_loss_fwd(const sqlx_ms_desc* d, const float* const* depth_lr, const float* target,
                                const float* const* sources_rgba, const float* const* color, const float* K,
                                const float* inv_K, const sqlx_pose_inputs* poses, const float* identity,
                                const float* const* noise, float* loss, uint8_t* const* argmin, void* saved,
                                size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_ms(d, poses)) return e;
  SQLX_REQUIRE(depth_lr && target && sources_rgba && color && K && inv_K && loss && argmin && saved && workspace,
               "NULL pointer argument");
  const MsLayout L = ms_layout(d);
  SQLX_REQUIRE(saved_bytes >= L.saved_total && workspace_bytes >= L.ws_total, "saved / workspace buffer too small");
  const int ns = d->num_scales, B = d->photo.B, S = d->photo.S;
  const bool automask = d->photo.flags & SQLX_AUTOMASK;
  SQLX_REQUIRE(!automask || (identity && noise), "automask needs identity and noise");
  for (int s = 0; s < ns; ++s)
    SQLX_REQUIRE(depth_lr[s] && color[s] && argmin[s] && (!automask || noise[s]), "scale %d: NULL pointer", s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* sv = reinterpret_cast<uint8_t*>(saved);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* T = reinterpret_cast<float*>(sv + L.T);
  float* stats = reinterpret_cast<float*>(sv + L.stats);
  float* sums = reinterpret_cast<float*>(sv + L.sums);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws + L.counters);
  bool any_pose = false;
  for (int s = 0; s < S; ++s) any_pose |= poses->axisangle[s] != nullptr;
  const int rescale = (d->rescale_translation && any_pose) ? 1 : 0;
  const MsShapes sh = make_shapes(d, depth_lr, color, L, saved);
  if (cudaMemsetAsync(counters, 0, 256 + sizeof(unsigned int) * (size_t)ns * B, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(counters)");
  {
    ProfScope prof("ms_stats_pose_kernel", st);
    ms_stats_pose_kernel<<<dim3(ms_blocks_one_wave(d->photo.H, d->photo.W, B, ns), B, ns), 256, 0, st>>>(
        sh, make_pose(poses, S, nullptr, nullptr), rescale, reinterpret_cast<float*>(ws + L.partial), counters + 64,
        stats, T);
    if (int e = check_launch("ms_stats_pose_kernel")) return e;
  }
  int ctas = 0;
  if (ns > 1 && ms_fusable(d)) {
    // every scale in one launch: the target tile, its statistics and the identity losses are shared by the scales
    const sqlx_photo_desc pd = scale_photo_desc(d, 0);
    const float* dups[SQLX_MAX_SCALES];
    for (int s = 0; s < ns; ++s) dups[s] = sh.d_up[s];
    if (int e = photo_fwd3_ms_launch(&pd, ns, dups, target, sources_rgba, K, inv_K, T, (size_t)B * S * 16, identity,
                                     automask ? noise : nullptr, reinterpret_cast<float*>(ws + L.photo_partial), L.max_ctas,
                                     &ctas, argmin, L.coef_stride ? reinterpret_cast<float*>(sv + L.coef) : nullptr,
                                     L.coef_stride / sizeof(float), st))
      return e;
  } else
  for (int s = 0; s < ns; ++s) {
    const sqlx_photo_desc pd = scale_photo_desc(d, s);
    if (int e = photo_fwd3_launch(&pd, depth_lr[s], sh.d_up[s], target, sources_rgba, K, inv_K, T + (size_t)s * B * S * 16, identity,
                                  automask ? noise[s] : nullptr,
                                  reinterpret_cast<float*>(ws + L.photo_partial) + (size_t)s * L.max_ctas, &ctas, argmin[s],
                                  L.coef_stride ? reinterpret_cast<float*>(sv + L.coef + L.coef_stride * s) : nullptr, st))
      return e;
  }
  {
    ProfScope prof("ms_smooth_loss_kernel", st);
    ms_smooth_loss_kernel<<<dim3(ms_blocks_one_wave(d->photo.H, d->photo.W, B, ns), B, ns), 256, sizeof(float) * (size_t)ns * B, st>>>(sh, reinterpret_cast<float*>(ws + L.spartial), sums,
                                                                  reinterpret_cast<float*>(ws + L.photo_partial),
                                                                  L.max_ctas, ctas, counters, loss);
    if (int e = check_launch("ms_smooth_loss_kernel")) return e;
  }
  return SQLX_OK;
}

extern "C" int sqlx_ms_loss_bwd(const sqlx_ms_desc* d, const float* const* depth_lr, const float* target,
                                const float* const* sources_rgba, const float* const* color, const float* K,
                                const float* inv_K, const sqlx_pose_inputs* poses, const uint8_t* const* argmin,
                                const float* g_loss, const void* saved, float* const* d_depth_lr,
                                float* const* d_axisangle, float* const* d_translation, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (int e = check_ms(d, poses)) return e;
  SQLX_REQUIRE(depth_lr && target && sources_rgba && color && K && inv_K && argmin && g_loss && saved && d_depth_lr &&
               workspace, "NULL pointer argument");
  const MsLayout L = ms_layout(d);
  SQLX_REQUIRE(workspace_bytes >= L.ws_total, "workspace too small");
  const int ns = d->num_scales, B = d->photo.B, S = d->photo.S, H = d->photo.H, W = d->photo.W;
  for (int s = 0; s < ns; ++s) SQLX_REQUIRE(depth_lr[s] && color[s] && argmin[s] && d_depth_lr[s], "scale %d: NULL pointer", s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(saved);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const float* T = reinterpret_cast<const float*>(sv + L.T);
  const float* stats = reinterpret_cast<const float*>(sv + L.stats);
  const float* sums = reinterpret_cast<const float*>(sv + L.sums);
  float* g_stats = reinterpret_cast<float*>(ws + L.g_stats);
  float* dP = reinterpret_cast<float*>(ws + L.dP);
  bool any_pose = false;
  for (int s = 0; s < S; ++s) any_pose |= poses->axisangle[s] != nullptr;
  const int rescale = (d->rescale_translation && any_pose) ? 1 : 0;
  const MsShapes sh = make_shapes(d, depth_lr, color, L, const_cast<void*>(saved));
  MsGrads g;
  for (int s = 0; s < SQLX_MAX_SCALES; ++s) {
    const bool on = s < ns;
    const bool direct = on && d->Hc[s] == d->h[s] && d->Wc[s] == d->w[s] && !(d->h[s] == H && d->w[s] == W);
    g.g_up[s] = on ? reinterpret_cast<float*>(ws + L.g_up + L.g_up_stride * s) : nullptr;
    g.g_direct[s] = direct ? reinterpret_cast<float*>(ws + L.g_direct + L.g_direct_stride * s) : nullptr;
    g.d_depth[s] = on ? d_depth_lr[s] : nullptr;
  }
  {
    ProfScope prof("ms_smooth_bwd_kernel", st);
    ms_smooth_bwd_kernel<<<dim3(64, B, ns), 256, 0, st>>>(sh, g, g_loss, sums);
    if (int e = check_launch("ms_smooth_bwd_kernel")) return e;
  }
  if (cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)ns * B * S * 12, st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(dP)");
  if (ns > 1 && ms_fusable(d)) {
    const sqlx_photo_desc pd = scale_photo_desc(d, 0);
    const float* dups[SQLX_MAX_SCALES];
    float* qups[SQLX_MAX_SCALES];
    unsigned acc = 0;
    for (int s = 0; s < ns; ++s) {
      dups[s] = sh.d_up[s];
      qups[s] = rescale ? reinterpret_cast<float*>(ws + L.q_up + L.g_up_stride * s) : nullptr;
      if (!g.g_direct[s]) acc |= 1u << s;
    }
    if (int e = photo_bwd3_ms_launch(&pd, ns, dups, target, sources_rgba, K, inv_K, T, (size_t)B * S * 16, argmin,
                                     L.coef_stride ? reinterpret_cast<const float*>(sv + L.coef) : nullptr,
                                     L.coef_stride / sizeof(float), g_loss, 1.f / ((float)ns * (float)B * H * W), g.g_up,
                                     qups, acc, dP, (size_t)B * S * 12, st))
      return e;
  } else
  for (int s = 0; s < ns; ++s) {
    const sqlx_photo_desc pd = scale_photo_desc(d, s);
    if (int e = photo_bwd3_launch(&pd, depth_lr[s], sh.d_up[s], target, sources_rgba, K, inv_K, T + (size_t)s * B * S * 16, argmin[s],
                                  L.coef_stride ? reinterpret_cast<const float*>(sv + L.coef + L.coef_stride * s) : nullptr,
                                  g_loss, 1.f / ((float)ns * (float)B * H * W), nullptr, g.g_up[s],
                                  rescale ? reinterpret_cast<float*>(ws + L.q_up + L.g_up_stride * s) : nullptr,
                                  g.g_direct[s] ? 0 : 1, dP + (size_t)s * B * S * 12, st))
      return e;
  }
  if (any_pose) {
    const int per_b = 7 * S;
    const int spb = 252 / per_b;   // samples per block (blockDim <= 256 = shared array size)
    ProfScope prof("ms_pose_bwd_kernel", st);
    ms_pose_bwd_kernel<<<ceil_div(B, spb), spb * per_b, 0, st>>>(sh, make_pose(poses, S, d_axisangle, d_translation),
                                                                rescale, K, stats, dP, g_stats);
    if (int e = check_launch("ms_pose_bwd_kernel")) return e;
  }
  {
    MsAdjoint ad;
    long long max_threads = 0;
    ad.rHW = 1.f / ((float)H * (float)W);
    for (int s = 0; s < SQLX_MAX_SCALES; ++s) {
      ad.q_up[s] = s < ns ? reinterpret_cast<const float*>(ws + L.q_up + L.g_up_stride * s) : nullptr;
      ad.log2_tpc[s] = 0;
      ad.factor[s] = 0;
      ad.sy[s] = ad.sx[s] = ad.ry[s] = ad.rx[s] = 1.f;
      if (s >= ns) continue;
      ad.sy[s] = (float)d->h[s] / (float)H; ad.sx[s] = (float)d->w[s] / (float)W;
      ad.ry[s] = (float)H / (float)d->h[s]; ad.rx[s] = (float)W / (float)d->w[s];
      const int fw = ceil_div(W, d->w[s]);      // footprint columns of a cell ~ 2 * fw (+1)
      const int fi = H / d->h[s];
      if (H == fi * d->h[s] && W == fi * d->w[s] && (fi == 1 || (fi & 1) == 0)) ad.factor[s] = fi;
      int lt = 2;
      while (lt < 5 && (1 << lt) < 2 * fw) ++lt;
      if (ad.factor[s] == 1 || ad.factor[s] == 2) lt = 0;      // must match the adjoint_even<F, TPC> instantiations
      else if (ad.factor[s] == 4) lt = 1;
      else if (ad.factor[s] == 8) lt = 3;
      else if (ad.factor[s] == 16) lt = 5;
      ad.log2_tpc[s] = lt;
      const long long t = ((long long)B * d->h[s] * d->w[s]) << lt;
      if (t > max_threads) max_threads = t;
    }
    SQLX_REQUIRE(max_threads < (1ll << 31), "depth maps too large for the upsample-adjoint launch");
    ProfScope prof("ms_upsample_adjoint_kernel", st);
    ms_upsample_adjoint_kernel<<<dim3((unsigned)((max_threads + 255) / 256), ns), 256, 0, st>>>(sh, g, ad, rescale, g_stats);
    if (int e = check_launch("ms_upsample_adjoint_kernel")) return e;
  }
  return SQLX_OK;
}
