// Per-sample median scaling of the supervised fine-tuning path: finetune/train_ft_SQLdepth.py:236-266.
//
// The reference loops over the first B/2 samples on the HOST: pred[i] and depth[i] are copied to NumPy (one device ->
// host synchronisation per sample), masked by  min_depth_eval < depth < max_depth_eval  and a crop rectangle, and
//     ratio_i = np.median(depth_i[valid]) / np.median(pred_i[valid])        (1 when either median is NaN)
// scales pred[i].  Here: an exact RADIX SELECT of the two middle order statistics over the valid pixels (np.median = their
// mean), no sort, no host round trip.  Four 8-bit digit passes + one "next larger key" pass, each a launch of
// (slices, 2 arrays, count samples) CTAs that histogram their slice in shared memory (warp-aggregated atomics) and add
// the non-empty bins to a global histogram; the digit decisions are re-derived from the global histograms in the
// prologue of the following launch, so there is no grid-wide spin-wait.  (First version: ONE CTA per (array, sample)
// walking the whole frame six times -- a chain of dependent load latencies: 800-870 us for four 320x1024 samples.)
// NaN semantics follow NumPy: a NaN among the selected values, or an empty selection, makes the median NaN.
// The prediction may be given at its own (lower) resolution: it is then read through the align_corners=True bilinear
// resize of train_ft_SQLdepth.py:235, so the resized map is never materialised.
#include "common.cuh"

namespace sqlx {

// monotone map float -> uint32 (total order of IEEE floats; NaNs are filtered before)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct MedParams {
  const float* pred;    // [B, h*w]: the prediction at its own resolution; read through the align_corners=True bilinear
                        // resize to H x W of train_ft_SQLdepth.py:235 (the identity when h == H and w == W)
  const float* depth;   // [B, H*W]
  int B, H, W, count;
  int h, w;             // resolution of pred
  float sy, sx;         // (h - 1) / (H - 1), (w - 1) / (W - 1)
  float lo, hi;         // valid: lo < depth < hi
  int r0, r1, c0, c1;   // crop rectangle [r0, r1) x [c0, c1)
  unsigned int* hist;   // [4 passes][2][B][256]   global digit histograms (zeroed by the call)
  unsigned int* nans;   // [2][B]                  NaNs among the selected values
  unsigned int* le;     // [2][B]                  selected values <= the lower middle key
  unsigned int* nxt;    // [2][B]                  max over (~key) of the selected values above it (0 = none)
  unsigned int* done;   // [B]                     arrival counters of the last launch
  float* med;           // [2][B]
  float* ratio;         // [B]
};

// element (r, c) of the array a CTA works on: the ground truth itself, or the prediction resized to its resolution
// (F.interpolate(mode="bilinear", align_corners=True): src = dst * (in - 1) / (out - 1), as in csrc/silog.cu)
__device__ __forceinline__ float med_value(const float* __restrict__ v, const MedParams& p, bool is_pred, int r, int c) {
  if (!is_pred) return v[r * p.W + c];
  if (p.h == p.H && p.w == p.W) return v[r * p.w + c];
  const float fy = p.sy * (float)r, fx = p.sx * (float)c;
  int y0 = (int)fy, x0 = (int)fx;
  y0 = y0 > p.h - 1 ? p.h - 1 : y0;
  x0 = x0 > p.w - 1 ? p.w - 1 : x0;
  const int y1 = y0 + (y0 < p.h - 1 ? 1 : 0), x1 = x0 + (x0 < p.w - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  return (1.f - ly) * ((1.f - lx) * v[y0 * p.w + x0] + lx * v[y0 * p.w + x1]) +
         ly * ((1.f - lx) * v[y1 * p.w + x0] + lx * v[y1 * p.w + x1]);
}

// Calls f(value) for every selected pixel of this CTA's slice of the sample (valid ground truth inside the crop).
template <typename F>
__device__ __forceinline__ void for_selected(const float* __restrict__ v, bool is_pred, const float* __restrict__ d,
                                             const MedParams& p, F f) {
  // slice blockIdx.x of gridDim.x: whole rows of the crop rectangle, columns by threads
  const int rows = p.r1 - p.r0, per = (rows + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ra = p.r0 + (int)blockIdx.x * per, rb = min(p.r1, ra + per);
  const int cw = p.c1 - p.c0;
  const int items = max(rb - ra, 0) * cw;
  for (int base = threadIdx.x; base < items; base += 4 * blockDim.x) {
    float dv[4];
    int rr[4], cc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = base + j * blockDim.x;
      rr[j] = ra + it / cw;
      cc[j] = p.c0 + it % cw;
      dv[j] = it < items ? d[rr[j] * p.W + cc[j]] : -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (dv[j] > p.lo && dv[j] < p.hi) f(is_pred ? med_value(v, p, true, rr[j], cc[j]) : dv[j]);
  }
}

constexpr int kMedThreads = 256;

// (prefix, k) of the radix select after `passes` digit passes, re-derived by every CTA from the global histograms.
// sh: 258 unsigned ints of shared memory.  Returns false when the median is NaN (empty selection or a NaN inside it).
__device__ bool med_state(const MedParams& p, int which, int b, int passes, bool upper, uint32_t* prefix_out, uint32_t* k_out,
                          uint32_t* sel_out, unsigned int* sh) {
  const unsigned int* h0 = p.hist + ((size_t)(0 * 2 + which) * p.B + b) * 256;
  // selection size = everything pass 0 counted (+ the NaNs it skipped)
  unsigned int part = 0;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) part += __ldcg(h0 + i);
  if (threadIdx.x == 0) sh[256] = 0;
  __syncthreads();
  if (part) atomicAdd(&sh[256], part);
  __syncthreads();
  const uint32_t nn = __ldcg(p.nans + which * p.B + b), sel = sh[256] + nn;
  __syncthreads();
  *sel_out = sel;
  if (sel == 0 || nn != 0) return false;
  uint32_t k = upper ? sel / 2 : (sel - 1) / 2, prefix = 0;
  for (int q = 0; q < passes; ++q) {
    const unsigned int* hq = p.hist + ((size_t)(q * 2 + which) * p.B + b) * 256;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = __ldcg(hq + i);
    __syncthreads();
    if (threadIdx.x == 0) {          // walk the 256 bins: the digit whose cumulative count first exceeds k
      uint32_t acc = 0, digit = 255;
      for (uint32_t i = 0; i < 256; ++i) {
        if (acc + sh[i] > k) { digit = i; break; }
        acc += sh[i];
      }
      sh[256] = digit;
      sh[257] = k - acc;
    }
    __syncthreads();
    prefix |= sh[256] << (24 - 8 * q);
    k = sh[257];
    __syncthreads();
  }
  *prefix_out = prefix;
  *k_out = k;
  return true;
}

// digit pass `pass` (0..3) of the lower middle order statistic.  grid (slices, 2, count)
__global__ void __launch_bounds__(kMedThreads) median_pass_kernel(MedParams p, int pass) {
  __shared__ unsigned int sh[258];
  __shared__ unsigned int hist[256];
  const int which = blockIdx.y, b = blockIdx.z;
  const float* d = p.depth + (size_t)b * p.H * p.W;
  const float* v = which == 0 ? d : p.pred + (size_t)b * p.h * p.w;
  uint32_t prefix = 0, k = 0, sel = 0;
  if (pass > 0 && !med_state(p, which, b, pass, false, &prefix, &k, &sel, sh)) return;     // NaN median: nothing to select
  const uint32_t pmask = pass == 0 ? 0u : 0xffffffffu << (32 - 8 * pass);
  const int shift = 24 - 8 * pass;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned int c_nan = 0;
  for_selected(v, which == 1, d, p, [&](float x) {
    if (x == x) {
      const uint32_t key = float_key(x);
      if ((key & pmask) == prefix) {
        // warp-aggregated: the lanes that hit the same bin elect one to add their count (depth values share their
        // high-order digits, so without this nearly every atomic of a pass lands on one or two shared-memory words)
        const uint32_t digit = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(__activemask(), digit);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned)__popc(peers));
      }
    } else {
      ++c_nan;
    }
  });
  __syncthreads();
  unsigned int* hg = p.hist + ((size_t)(pass * 2 + which) * p.B + b) * 256;
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (hist[i]) atomicAdd(hg + i, hist[i]);
  if (pass == 0 && c_nan) atomicAdd(p.nans + which * p.B + b, c_nan);
}

// last launch: the next order statistic (equal to the lower one when more than khi elements are <= it, else the smallest
// element above it), the medians and -- by the last CTA of a sample -- the ratio.  grid (slices, 2, count)
__global__ void __launch_bounds__(kMedThreads) median_final_kernel(MedParams p) {
  __shared__ unsigned int sh[258];
  __shared__ unsigned int red[2];
  __shared__ int last;
  const int which = blockIdx.y, b = blockIdx.z;
  const float* d = p.depth + (size_t)b * p.H * p.W;
  const float* v = which == 0 ? d : p.pred + (size_t)b * p.h * p.w;
  uint32_t key_lo = 0, k = 0, sel = 0;
  const bool ok = med_state(p, which, b, 4, false, &key_lo, &k, &sel, sh);
  const bool need_next = ok && (sel / 2 != (sel - 1) / 2);
  if (threadIdx.x < 2) red[threadIdx.x] = 0;
  __syncthreads();
  if (need_next) {
    unsigned int c_le = 0;
    uint32_t inv = 0;                         // max over ~key of the keys above key_lo (0 = none seen)
    for_selected(v, which == 1, d, p, [&](float x) {
      const uint32_t key = float_key(x);
      if (key <= key_lo) ++c_le;
      else inv = max(inv, ~key);
    });
    if (c_le) atomicAdd(&red[0], c_le);
    if (inv) atomicMax(&red[1], inv);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (red[0]) atomicAdd(p.le + which * p.B + b, red[0]);
      if (red[1]) atomicMax(p.nxt + which * p.B + b, red[1]);
    }
  }
  // arrival: the last CTA of the sample (all slices of both arrays) finishes the job
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&p.done[b], 1u) == gridDim.x * 2u - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float med[2];
  for (int wch = 0; wch < 2; ++wch) {
    uint32_t klo = 0, kk = 0, ss = 0;
    float m = __int_as_float(0x7fc00000);      // NaN: empty selection or a NaN inside it (numpy.median)
    if (med_state(p, wch, b, 4, false, &klo, &kk, &ss, sh)) {
      const float vlo = key_float(klo);
      float vhi = vlo;
      if (ss / 2 != (ss - 1) / 2) {
        const uint32_t c_le = __ldcg(p.le + wch * p.B + b), inv = __ldcg(p.nxt + wch * p.B + b);
        vhi = c_le > ss / 2 ? vlo : key_float(~inv);
      }
      m = 0.5f * (vlo + vhi);
    }
    med[wch] = m;
  }
  if (threadIdx.x == 0) {
    // train_ft_SQLdepth.py:261-264: ratio = 1 if either median is NaN else median(depth) / median(pred)
    p.ratio[b] = (med[0] != med[0] || med[1] != med[1]) ? 1.f : med[0] / med[1];
    p.med[b] = med[0];
    p.med[p.B + b] = med[1];
  }
}

__global__ void median_ones_kernel(float* ratio, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) ratio[i] = 1.f;
}

}  // namespace sqlx

using namespace sqlx;

static size_t med_ws_uints(int B) { return (size_t)4 * 2 * B * 256 + (size_t)2 * B * 3 + B + (size_t)2 * B; }

extern "C" size_t sqlx_median_ratio_workspace_bytes(int B) { return B > 0 ? sizeof(unsigned int) * med_ws_uints(B) : 0; }

/* ratio[i] = median(depth_i[valid]) / median(pred_i[valid]) for i < count, 1 for count <= i < B
 * (finetune/train_ft_SQLdepth.py:236-266).  pred, depth [B,H,W] fp32 at the ground-truth resolution; valid =
 * min_depth_eval < depth < max_depth_eval inside the crop rectangle rows [r0,r1) x columns [c0,c1).
 * The workspace is scratch: the call clears it itself. */
extern "C" int sqlx_median_ratio_resized(const float* pred, int h, int w, const float* depth, int B, int H, int W, int count,
                                         float min_depth_eval, float max_depth_eval, int r0, int r1, int c0, int c1,
                                         float* ratio, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int sqlx_median_ratio(const float* pred, const float* depth, int B, int H, int W, int count, float min_depth_eval,
                                 float max_depth_eval, int r0, int r1, int c0, int c1, float* ratio, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  return sqlx_median_ratio_resized(pred, H, W, depth, B, H, W, count, min_depth_eval, max_depth_eval, r0, r1, c0, c1, ratio,
                                   workspace, workspace_bytes, stream);
}

/* The same with pred [B,h,w] at its own resolution: every value is read through the align_corners=True bilinear resize to
 * H x W (train_ft_SQLdepth.py:235), so the resized map is never materialised. */
extern "C" int sqlx_median_ratio_resized(const float* pred, int h, int w, const float* depth, int B, int H, int W, int count,
                                         float min_depth_eval, float max_depth_eval, int r0, int r1, int c0, int c1,
                                         float* ratio, void* workspace, size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(pred && depth && ratio && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 0 && W > 0 && h > 0 && w > 0 && count >= 0 && count <= B,
               "bad shape B=%d H=%d W=%d h=%d w=%d count=%d", B, H, W, h, w, count);
  SQLX_REQUIRE((long long)H * W < (1ll << 31), "sample too large");
  SQLX_REQUIRE(workspace_bytes >= sqlx_median_ratio_workspace_bytes(B), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MedParams p;
  p.pred = pred; p.depth = depth; p.B = B; p.H = H; p.W = W; p.count = count;
  p.h = h; p.w = w;
  p.sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  p.sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  p.lo = min_depth_eval; p.hi = max_depth_eval;
  p.r0 = r0 < 0 ? 0 : r0; p.r1 = r1 > H ? H : r1; p.c0 = c0 < 0 ? 0 : c0; p.c1 = c1 > W ? W : c1;
  unsigned int* u = reinterpret_cast<unsigned int*>(workspace);
  p.hist = u; u += (size_t)4 * 2 * B * 256;
  p.nans = u; u += 2 * B;
  p.le = u; u += 2 * B;
  p.nxt = u; u += 2 * B;
  p.done = u; u += B;
  p.med = reinterpret_cast<float*>(u);
  p.ratio = ratio;
  ProfScope prof("median_ratio_kernel", st);
  median_ones_kernel<<<ceil_div(B, 256), 256, 0, st>>>(ratio, B);
  if (int e = check_launch("median_ones_kernel")) return e;
  if (count == 0 || p.r1 <= p.r0 || p.c1 <= p.c0) {
    // (an empty crop makes every median NaN: ratio 1, what the ones kernel wrote)
    return SQLX_OK;
  }
  if (cudaMemsetAsync(workspace, 0, sizeof(unsigned int) * med_ws_uints(B), st) != cudaSuccess)
    return check_launch("cudaMemsetAsync(median workspace)");
  // slices per (array, sample): enough CTAs to fill the machine, at least ~2 rows of the crop each
  int slices = (4 * kNumSMs) / (2 * count);
  const int rows = p.r1 - p.r0;
  slices = slices > rows / 2 ? rows / 2 : slices;
  slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
  const dim3 grid(slices, 2, count);
  for (int pass = 0; pass < 4; ++pass) {
    median_pass_kernel<<<grid, kMedThreads, 0, st>>>(p, pass);
    if (int e = check_launch("median_pass_kernel")) return e;
  }
  median_final_kernel<<<grid, kMedThreads, 0, st>>>(p);
  return check_launch("median_final_kernel");
}
