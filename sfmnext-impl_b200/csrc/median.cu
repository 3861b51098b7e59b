// Per-sample median scaling of the supervised fine-tuning path: finetune/train_ft_SQLdepth.py:236-266.
//
// The reference loops over the first B/2 samples on the HOST: pred[i] and depth[i] are copied to NumPy (one device ->
// host synchronisation per sample), masked by  min_depth_eval < depth < max_depth_eval  and a crop rectangle, and
//     ratio_i = np.median(depth_i[valid]) / np.median(pred_i[valid])        (1 when either median is NaN)
// scales pred[i].  Here: one launch, two CTAs per sample (one per array), an exact RADIX SELECT of the two middle order
// statistics over the valid pixels (np.median = their mean) -- four 8-bit digit passes over the sample with a 256-bin
// shared-memory histogram, no sort, no host round trip -- and the CTA that finishes last for a sample writes its ratio.
// NaN semantics follow NumPy: a NaN among the selected values, or an empty selection, makes the median NaN.
#include "common.cuh"

namespace sqlx {

constexpr int kMedThreads = 1024;

// monotone map float -> uint32 (total order of IEEE floats; NaNs are filtered before)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct MedParams {
  const float* pred;    // [B, H*W]
  const float* depth;   // [B, H*W]
  int B, H, W, count;
  float lo, hi;         // valid: lo < depth < hi
  int r0, r1, c0, c1;   // crop rectangle [r0, r1) x [c0, c1)
  float* med;           // [2][B] scratch: medians of depth / pred
  unsigned int* done;   // [B] arrival counters (zero on entry, left zero)
  float* ratio;         // [B]
};

// k-th smallest (0-based) key among the selected elements of `v`; selection = valid(depth) and not NaN(v).
// Block-wide; `hist` is 256 + 2 unsigned ints of shared memory.
__device__ uint32_t radix_select(const float* __restrict__ v, const float* __restrict__ d, const MedParams& p, uint32_t k,
                                 unsigned int* hist) {
  uint32_t prefix = 0, pmask = 0;
  const int n = p.H * p.W;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 258; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int r = i / p.W, c = i - r * p.W;
      const float dv = d[i], x = v[i];
      if (dv > p.lo && dv < p.hi && r >= p.r0 && r < p.r1 && c >= p.c0 && c < p.c1 && x == x) {
        const uint32_t key = float_key(x);
        if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {          // walk the 256 bins: the digit whose cumulative count first exceeds k
      uint32_t acc = 0, digit = 255;
      for (uint32_t b = 0; b < 256; ++b) {
        if (acc + hist[b] > k) { digit = b; break; }
        acc += hist[b];
      }
      hist[256] = digit;
      hist[257] = k - acc;
    }
    __syncthreads();
    prefix |= hist[256] << shift;
    pmask |= 255u << shift;
    k = hist[257];
    __syncthreads();
  }
  return prefix;
}

// grid (2, B): blockIdx.x = 0 depth, 1 pred; blockIdx.y = sample (samples >= count keep ratio 1)
__global__ void __launch_bounds__(kMedThreads) median_ratio_kernel(MedParams p) {
  __shared__ unsigned int hist[258];
  __shared__ unsigned int cnt_sh[3];
  const int which = blockIdx.x, b = blockIdx.y;
  if (b >= p.count) {                 // train_ft_SQLdepth.py:236 loops over the first half of the batch only
    if (which == 0 && threadIdx.x == 0) p.ratio[b] = 1.f;
    return;
  }
  const int n = p.H * p.W;
  const float* d = p.depth + (size_t)b * n;
  const float* v = which == 0 ? d : p.pred + (size_t)b * n;
  if (threadIdx.x < 3) cnt_sh[threadIdx.x] = 0;
  __syncthreads();
  // selection size and NaN count
  unsigned int c_sel = 0, c_nan = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / p.W, c = i - r * p.W;
    const float dv = d[i], x = v[i];
    if (dv > p.lo && dv < p.hi && r >= p.r0 && r < p.r1 && c >= p.c0 && c < p.c1) {
      ++c_sel;
      if (x != x) ++c_nan;
    }
  }
  atomicAdd(&cnt_sh[0], c_sel);
  atomicAdd(&cnt_sh[1], c_nan);
  __syncthreads();
  const unsigned int sel = cnt_sh[0], nans = cnt_sh[1];
  float med = __int_as_float(0x7fc00000);      // NaN: empty selection or a NaN inside it (numpy.median)
  if (sel > 0 && nans == 0) {
    const uint32_t klo = (sel - 1) / 2, khi = sel / 2;
    const uint32_t key_lo = radix_select(v, d, p, klo, hist);
    float vlo = key_float(key_lo), vhi = vlo;
    if (khi != klo) {
      // the next order statistic: equal to vlo when more than khi elements are <= vlo, else the smallest element > vlo
      unsigned int c_le = 0;
      uint32_t next = 0xffffffffu;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = i / p.W, c = i - r * p.W;
        const float dv = d[i], x = v[i];
        if (dv > p.lo && dv < p.hi && r >= p.r0 && r < p.r1 && c >= p.c0 && c < p.c1) {
          const uint32_t key = float_key(x);
          if (key <= key_lo) ++c_le;
          else next = min(next, key);
        }
      }
      if (threadIdx.x == 0) { cnt_sh[1] = 0; cnt_sh[2] = 0xffffffffu; }
      __syncthreads();
      atomicMin(&cnt_sh[2], next);
      atomicAdd(&cnt_sh[1], c_le);
      __syncthreads();
      vhi = cnt_sh[1] > khi ? vlo : key_float(cnt_sh[2]);
    }
    med = 0.5f * (vlo + vhi);
  }
  if (threadIdx.x == 0) {
    p.med[which * p.B + b] = med;
    __threadfence();
    if (atomicAdd(&p.done[b], 1u) == 1u) {      // second CTA of the sample: both medians are in
      __threadfence();
      const float md = *(volatile float*)&p.med[b], mp = *(volatile float*)&p.med[p.B + b];
      // train_ft_SQLdepth.py:261-264: ratio = 1 if either median is NaN else median(depth) / median(pred)
      p.ratio[b] = (md != md || mp != mp) ? 1.f : md / mp;
      p.done[b] = 0u;
    }
  }
}

}  // namespace sqlx

using namespace sqlx;

extern "C" size_t sqlx_median_ratio_workspace_bytes(int B) { return B > 0 ? 256 + sizeof(float) * 2 * (size_t)B : 0; }

/* ratio[i] = median(depth_i[valid]) / median(pred_i[valid]) for i < count, 1 for count <= i < B
 * (finetune/train_ft_SQLdepth.py:236-266).  pred, depth [B,H,W] fp32 at the ground-truth resolution; valid =
 * min_depth_eval < depth < max_depth_eval inside the crop rectangle rows [r0,r1) x columns [c0,c1).
 * workspace must be zero-initialised once (the kernel leaves it zero). */
extern "C" int sqlx_median_ratio(const float* pred, const float* depth, int B, int H, int W, int count, float min_depth_eval,
                                 float max_depth_eval, int r0, int r1, int c0, int c1, float* ratio, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(pred && depth && ratio && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 0 && W > 0 && count >= 0 && count <= B, "bad shape B=%d H=%d W=%d count=%d", B, H, W, count);
  SQLX_REQUIRE((long long)H * W < (1ll << 31), "sample too large");
  SQLX_REQUIRE(workspace_bytes >= sqlx_median_ratio_workspace_bytes(B), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MedParams p;
  p.pred = pred; p.depth = depth; p.B = B; p.H = H; p.W = W; p.count = count;
  p.lo = min_depth_eval; p.hi = max_depth_eval;
  p.r0 = r0 < 0 ? 0 : r0; p.r1 = r1 > H ? H : r1; p.c0 = c0 < 0 ? 0 : c0; p.c1 = c1 > W ? W : c1;
  p.done = reinterpret_cast<unsigned int*>(workspace);
  p.med = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256) ;
  SQLX_REQUIRE(B <= 64, "batch %d exceeds the arrival-counter block (64)", B);
  p.ratio = ratio;
  ProfScope prof("median_ratio_kernel", st);
  median_ratio_kernel<<<dim3(2, B), kMedThreads, 0, st>>>(p);
  return check_launch("median_ratio_kernel");
}
