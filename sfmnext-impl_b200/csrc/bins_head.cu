// Bins head of the SQLdepth decoder: the three nn.Linear of `bins_regressor` and the bin-centre
// arithmetic (networks/depth_decoder_QTR.py:48-66) as weight-streaming kernels for a handful of samples per GPU.
//   y_b     = bins_regressor(summary.view(B, Q*E))            Linear(QE -> 16Q) LeakyReLU Linear(16Q -> 256) LeakyReLU Linear(256 -> D)
//   y       = relu(y_b) + 0.1 ; y /= sum_d y                  (norm == 'linear')
//   widths  = (max - min) * y ; edges = cumsum(pad(widths, min)) ; centers = (edges[:-1] + edges[1:]) / 2
// With M = batch <= 16 these layers are bound by reading (forward, d_input) and writing (d_weight) the weight
// matrices once -- 8.4 MB for the first layer at Q = 64 -- which cuBLAS does through ~12 small-M GEMM / split-K /
// elementwise launches per direction.  Here: one CTA per output row streams the row with 128-bit loads against all
// samples at once; 4 launches forward, 4 backward, exact fp32, no atomics.
#include "common.cuh"

namespace sqlx {

constexpr int kHeadMaxB = 16;
constexpr float kLeakySlope = 0.01f;   // nn.LeakyReLU() default negative_slope

// Sum 16 per-lane values over the warp (16 shuffles): lanes 2i and 2i+1 end with the total of v[i] in v[0].
__device__ __forceinline__ void head_reduce16(float (&v)[16]) {
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

// y[b,n] = act(sum_k W[n,k] x[b,k] + bias[n]); one 4-warp CTA per output row n, K split across the 128 threads.
// (A variant that stages x once per CTA in shared memory and streams one row per warp measured SLOWER under ncu --
// 18.6 / 11.2 / 6.1 us against 14.4 / 8.6 / 5.6 us for the three layers: x is L2-resident and 1024 small CTAs hide the
// latency better than 128 large ones.  The same staging does pay in the weight-gradient half of the backward.)
__global__ void __launch_bounds__(128) head_linear_fwd_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                              const float* __restrict__ x, int B, int N, int K, int leaky,
                                                              float* __restrict__ y) {
  __shared__ float part[4][kHeadMaxB];
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[kHeadMaxB];
#pragma unroll
  for (int b = 0; b < kHeadMaxB; ++b) acc[b] = 0.f;
  const float4* wr = reinterpret_cast<const float4*>(W + (size_t)n * K);
  for (int k4 = threadIdx.x; k4 < K / 4; k4 += 128) {
    const float4 wv = __ldg(wr + k4);
#pragma unroll
    for (int b = 0; b < kHeadMaxB; ++b) {
      if (b < B) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * K) + k4);
        acc[b] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[b]))));
      }
    }
  }
  head_reduce16(acc);
  if (!(lane & 1)) part[warp][lane >> 1] = acc[0];
  __syncthreads();
  const int b = threadIdx.x;
  if (b < B) {
    float v = ((part[0][b] + part[1][b]) + (part[2][b] + part[3][b])) + __ldg(bias + n);
    if (leaky) v = v > 0.f ? v : kLeakySlope * v;
    y[(size_t)b * N + n] = v;
  }
}

// Backward of one Linear layer in ONE launch (both halves only need dy, y, x, W):
//   dz[b,n] = dy[b,n] * act'(y[b,n]);  dW[n,k] = sum_b dz[b,n] x[b,k];  db[n] = sum_b dz[b,n];  dx[b,k] = sum_n dz[b,n] W[n,k]
// The first `nx` CTAs compute dx (the critical path of the backward chain: the next layer waits for it), the others dW;
// both halves fit two CTAs per SM so that the whole grid is ONE wave and the halves overlap on the machine.
//   dx CTA: 32 input features k (lane = k: W rows are read as coalesced 128-byte segments); the 16 warps split the
//           output rows n (16 rows in flight each), dz is derived from (dy, y) and staged once per CTA as [n][Bp samples]
//           (128-bit broadcast loads), partial sums meet in shared memory: no atomics, fixed summation order.
//   dW CTA: a (1024-column slab, 16..64-row chunk) block of dW (chunk sized so that the grid fills the machine once): the slab of x is staged once in shared memory (every sample's
//           loads in flight together), then one row per warp at a time: 128-bit stores against all samples at once.
// (Round 2, first version: the dW CTAs staged ALL of x -- 128 KB, one load in flight per thread and trip -- and the 128 KB
// of either half allowed one CTA per SM, i.e. two waves: 75 us for the 2048 x 4096 layer under ncu.)
constexpr int kHeadXWarps = 16;
constexpr int kHeadBwdThreads = 32 * kHeadXWarps;
constexpr int kHeadSlab = 1024;      // dW: columns per CTA
constexpr size_t kHeadPartBytes = sizeof(float) * kHeadXWarps * kHeadMaxB * 33;

__global__ void __launch_bounds__(kHeadBwdThreads) head_linear_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, const float* __restrict__ W,
    int B, int Bp, int N, int K, int leaky, int nx, int nslab, int rows_per_cta, float* __restrict__ dz, float* __restrict__ dW,
    float* __restrict__ db, float* __restrict__ dx) {
  extern __shared__ __align__(16) float smem[];           // dx CTAs: dz [N][Bp], then the partial sums; dW CTAs: x slab
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x >= nx) {
    // ---- a (slab, chunk) block of dW (and, from the CTAs of slab 0, db and dz)
    const int id = blockIdx.x - nx, slab = id % nslab, chunk = id / nslab;
    const int k0 = slab * kHeadSlab, kw4 = min(kHeadSlab, K - k0) >> 2;
    float4* sx4 = reinterpret_cast<float4*>(smem);          // [B][kw4]
    for (int c = threadIdx.x; c < kw4; c += kHeadBwdThreads) {
      float4 v[kHeadMaxB];
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b)
        if (b < B) v[b] = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * K + k0) + c);
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b)
        if (b < B) sx4[b * kw4 + c] = v[b];
    }
    __syncthreads();
    const int n_end = min(N, (chunk + 1) * rows_per_cta);
    for (int n = chunk * rows_per_cta + warp; n < n_end; n += kHeadXWarps) {
      float g[kHeadMaxB];
      float bsum = 0.f;
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b) {
        g[b] = 0.f;
        if (b < B) {
          float v = __ldg(dy + (size_t)b * N + n);
          if (leaky && !(__ldg(y + (size_t)b * N + n) > 0.f)) v *= kLeakySlope;
          g[b] = v;
          bsum += v;
        }
      }
      if (slab == 0) {
        if (lane < B) {
          float mine = 0.f;
#pragma unroll
          for (int b = 0; b < kHeadMaxB; ++b) mine = (b == lane) ? g[b] : mine;
          dz[(size_t)lane * N + n] = mine;
        }
        if (lane == 0) db[n] = bsum;
      }
      float4* out = reinterpret_cast<float4*>(dW + (size_t)n * K + k0);
      for (int k4 = lane; k4 < kw4; k4 += 32) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < kHeadMaxB; ++b) {
          if (b < B) {
            const float4 xv = sx4[b * kw4 + k4];
            a.x = fmaf(g[b], xv.x, a.x); a.y = fmaf(g[b], xv.y, a.y); a.z = fmaf(g[b], xv.z, a.z); a.w = fmaf(g[b], xv.w, a.w);
          }
        }
        out[k4] = a;
      }
    }
    return;
  }
  // ---- 32 columns of dx
  float* sdz = smem;
  const int k = blockIdx.x * 32 + lane;
  // dz staging: one output row n per thread and trip, all samples' dy (and y) loads in flight together; every load is
  // coalesced across the threads (consecutive n), the Bp values of a row leave as 128-bit shared stores
  for (int n = threadIdx.x; n < N; n += kHeadBwdThreads) {
    float v[kHeadMaxB], a[kHeadMaxB];
#pragma unroll
    for (int b = 0; b < kHeadMaxB; ++b) {
      v[b] = b < B ? __ldg(dy + (size_t)b * N + n) : 0.f;
      a[b] = (b < B && leaky) ? __ldg(y + (size_t)b * N + n) : 1.f;
    }
#pragma unroll
    for (int b = 0; b < kHeadMaxB; ++b) v[b] = (a[b] > 0.f) ? v[b] : v[b] * kLeakySlope;
    float4* dst = reinterpret_cast<float4*>(sdz + (size_t)n * Bp);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * q < Bp) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  __syncthreads();
  float acc[kHeadMaxB];
#pragma unroll
  for (int b = 0; b < kHeadMaxB; ++b) acc[b] = 0.f;
  const int per = (N + kHeadXWarps - 1) / kHeadXWarps;
  const int n0 = warp * per, n1 = min(N, n0 + per);
  const bool kin = k < K;
  constexpr int kRows = 16;                     // weight rows in flight per warp and trip
  for (int nb = n0; nb < n1; nb += kRows) {
    float wv[kRows];
#pragma unroll
    for (int j = 0; j < kRows; ++j) wv[j] = (kin && nb + j < n1) ? __ldg(W + (size_t)(nb + j) * K + k) : 0.f;
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      if (nb + j >= n1) break;
      const float4* zr = reinterpret_cast<const float4*>(sdz + (size_t)(nb + j) * Bp);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (4 * q >= Bp) break;
        const float4 z = zr[q];
        acc[4 * q] = fmaf(z.x, wv[j], acc[4 * q]); acc[4 * q + 1] = fmaf(z.y, wv[j], acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(z.z, wv[j], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(z.w, wv[j], acc[4 * q + 3]);
      }
    }
  }
  __syncthreads();                              // every warp is done with the staged dz: its memory holds the partials now
  float* part = smem;                           // [warps][kHeadMaxB][33]
#pragma unroll
  for (int b = 0; b < kHeadMaxB; ++b) part[(warp * kHeadMaxB + b) * 33 + lane] = acc[b];
  __syncthreads();
  for (int i = threadIdx.x; i < B * 32; i += kHeadBwdThreads) {
    const int b = i >> 5, l = i & 31;
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kHeadXWarps; ++w2) t += part[(w2 * kHeadMaxB + b) * 33 + l];
    if (blockIdx.x * 32 + l < K) dx[(size_t)b * K + blockIdx.x * 32 + l] = t;
  }
}

// ---- bin centres (depth_decoder_QTR.py:51-66, norm == 'linear'); one block of 256 threads per sample, D <= 256
__global__ void __launch_bounds__(256) head_centers_fwd_kernel(const float* __restrict__ raw, int D, float min_val,
                                                               float max_val, float* __restrict__ centers) {
  __shared__ float sh[256];
  const int b = blockIdx.x, d = threadIdx.x;
  const float y = d < D ? fmaxf(__ldg(raw + (size_t)b * D + d), 0.f) + 0.1f : 0.f;
  sh[d] = y;
  __syncthreads();
  // inclusive scan (Hillis-Steele, D <= 256)
  for (int o = 1; o < 256; o <<= 1) {
    const float t = d >= o ? sh[d - o] : 0.f;
    __syncthreads();
    sh[d] += t;
    __syncthreads();
  }
  const float total = sh[255];
  const float c = (max_val - min_val) / total;
  if (d < D) {
    const float incl = sh[d], excl = incl - y;                  // cumulative y up to and excluding / including bin d
    centers[(size_t)b * D + d] = min_val + c * (excl + 0.5f * y);   // 0.5 (edge_d + edge_{d+1})
  }
}

// centers_d = min + c/s (sum_{j<d} y_j + y_d / 2), s = sum y, c = max - min
//   d centers_d / d y_j = c/s ([j<d] + [j==d]/2) - (centers_d - min) / s
__global__ void __launch_bounds__(256) head_centers_bwd_kernel(const float* __restrict__ raw,
                                                               const float* __restrict__ centers,
                                                               const float* __restrict__ g_centers, int D, float min_val,
                                                               float max_val, float* __restrict__ d_raw) {
  __shared__ float sy[256], sg[256], sgc[256];
  const int b = blockIdx.x, d = threadIdx.x;
  const float r = d < D ? __ldg(raw + (size_t)b * D + d) : 0.f;
  const float y = d < D ? fmaxf(r, 0.f) + 0.1f : 0.f;
  const float g = d < D ? __ldg(g_centers + (size_t)b * D + d) : 0.f;
  const float cen = d < D ? __ldg(centers + (size_t)b * D + d) : min_val;
  sy[d] = y;
  sg[d] = g;                        // suffix sums of g
  sgc[d] = g * (cen - min_val);     // sum_d g_d (centers_d - min)
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {            // totals of y and g (centers - min)
    if (d < o) { sy[d] += sy[d + o]; sgc[d] += sgc[d + o]; }
    __syncthreads();
  }
  const float s = sy[0], dot = sgc[0];
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {            // inclusive suffix scan of g
    const float t = d + o < 256 ? sg[d + o] : 0.f;
    __syncthreads();
    sg[d] += t;
    __syncthreads();
  }
  if (d < D) {
    const float c = (max_val - min_val) / s;
    const float suffix_excl = sg[d] - g;          // sum_{d' > d} g_d'
    const float dy = c * (suffix_excl + 0.5f * g) - dot / s;
    d_raw[(size_t)b * D + d] = r > 0.f ? dy : 0.f;
  }
}

}  // namespace sqlx

using namespace sqlx;

extern "C" int sqlx_head_linear_fwd(const float* W, const float* bias, const float* x, int B, int N, int K, int leaky,
                                    float* y, void* stream) {
  SQLX_REQUIRE(W && bias && x && y, "NULL pointer argument");
  SQLX_REQUIRE(B >= 1 && B <= kHeadMaxB, "batch %d outside 1..%d", B, kHeadMaxB);
  SQLX_REQUIRE(N >= 1 && K >= 4 && K % 4 == 0, "in_features must be a positive multiple of 4 (got %d)", K);
  SQLX_REQUIRE(((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, "W and x must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope prof("head_linear_fwd_kernel", st);
  head_linear_fwd_kernel<<<N, 128, 0, st>>>(W, bias, x, B, N, K, leaky, y);
  return check_launch("head_linear_fwd_kernel");
}

/* dy: gradient wrt the layer output y (after the activation when leaky); dz [B,N] scratch; dW [N,K], db [N] overwritten;
 * dx [B,K] overwritten (may be NULL: first layer input without gradient). */
extern "C" int sqlx_head_linear_bwd(const float* W, const float* x, const float* y, const float* dy, int B, int N, int K,
                                    int leaky, float* dz, float* dW, float* db, float* dx, void* stream) {
  SQLX_REQUIRE(W && x && dy && dz && dW && db && (!leaky || y), "NULL pointer argument");
  SQLX_REQUIRE(B >= 1 && B <= kHeadMaxB, "batch %d outside 1..%d", B, kHeadMaxB);
  SQLX_REQUIRE(N >= 1 && K >= 4 && K % 4 == 0, "in_features must be a positive multiple of 4 (got %d)", K);
  SQLX_REQUIRE(((reinterpret_cast<uintptr_t>(dW) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, "dW and x must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nx = dx ? ceil_div(K, 32) : 0;
  const int Bp = (B + 3) / 4 * 4;
  const int nslab = ceil_div(K, kHeadSlab);
  // rows per dW CTA: a multiple of the 16 warps, as small as keeps the whole grid within one wave of two CTAs per SM
  int target = 2 * kNumSMs - nx;
  target = target < 64 ? 64 : target;
  int rows_per_cta = kHeadXWarps * ceil_div(N * nslab, kHeadXWarps * target);
  rows_per_cta = rows_per_cta > 64 ? 64 : rows_per_cta;
  const int nchunk = ceil_div(N, rows_per_cta);
  const size_t smem_dx = dx ? sizeof(float) * (size_t)N * Bp : 0;
  const size_t smem_dw = sizeof(float) * (size_t)B * (K < kHeadSlab ? K : kHeadSlab);
  size_t smem = smem_dx > smem_dw ? smem_dx : smem_dw;
  smem = smem > kHeadPartBytes ? smem : kHeadPartBytes;
  SQLX_REQUIRE(smem <= 200 * 1024, "layer %d x %d too large for the shared-memory staging", N, K);
  if (int e = ensure_dyn_smem(head_linear_bwd_kernel, 200 * 1024)) return e;   // the cap checked above, once per device
  ProfScope prof("head_linear_bwd_kernel", st);
  head_linear_bwd_kernel<<<nx + nslab * nchunk, kHeadBwdThreads, smem, st>>>(dy, y, x, W, B, Bp, N, K, leaky, nx, nslab, rows_per_cta, dz,
                                                                           dW, db, dx);
  return check_launch("head_linear_bwd_kernel");
}

extern "C" int sqlx_head_centers_fwd(const float* raw, int B, int D, float min_val, float max_val, float* centers,
                                     void* stream) {
  SQLX_REQUIRE(raw && centers, "NULL pointer argument");
  SQLX_REQUIRE(B >= 1 && D >= 1 && D <= 256, "dim_out %d outside 1..256", D);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  head_centers_fwd_kernel<<<B, 256, 0, st>>>(raw, D, min_val, max_val, centers);
  return check_launch("head_centers_fwd_kernel");
}

extern "C" int sqlx_head_centers_bwd(const float* raw, const float* centers, const float* g_centers, int B, int D,
                                     float min_val, float max_val, float* d_raw, void* stream) {
  SQLX_REQUIRE(raw && centers && g_centers && d_raw, "NULL pointer argument");
  SQLX_REQUIRE(B >= 1 && D >= 1 && D <= 256, "dim_out %d outside 1..256", D);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  head_centers_bwd_kernel<<<B, 256, 0, st>>>(raw, centers, g_centers, D, min_val, max_val, d_raw);
  return check_launch("head_centers_bwd_kernel");
}
