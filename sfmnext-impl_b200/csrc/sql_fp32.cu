// SQL block (Self Query Layer tail of the depth decoder), fp32 CUDA-core kernels.
//
// These are the exact-fp32 kernels: the parity ground truth on the GPU and the path taken for shapes the
// tensor-core kernels (sql_tc.cu) do not cover.  No [pixels x queries] tensor is written to HBM: every kernel
// recomputes the self-cost volume y = x^T K tile by tile (64 pixels) in shared memory.
//
// Reference lines (relative to /root/reference):
//   y, pixel-softmax, summary     networks/layers.py:17-19
//   1x1 conv + Softmax(dim=1)     networks/depth_decoder_QTR.py:28-29,61
//   pred = sum_d prob * centers   networks/depth_decoder_QTR.py:70
// Backward formulas: SURVEY.md Appendix A.1.
#include "common.cuh"

#include <atomic>

#include <math.h>

namespace sqlx {

constexpr int kTP = 64;        // pixels per tile
constexpr int kNT = 256;       // threads per CTA
constexpr int kMaxQ = 128;     // query_nums limit of the fp32 path
constexpr int kMaxD = 128;     // dim_out limit of the fp32 path
constexpr int kLD = 132;       // leading dimension of the [pixel][q or d] shared tiles (128 + 4: conflict-free float4 rows)
constexpr int kNJ = kMaxD / 4; // logits per thread (thread = (pixel, d mod 4))

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct ChunkPlan {
  int tiles, chunks, tiles_per_chunk;
};
// `target` CTAs in total over B samples
static inline ChunkPlan plan_chunks(int B, int n, int target) {
  ChunkPlan c;
  c.tiles = ceil_div(n, kTP);
  int chunks = target / B;
  if (chunks < 1) chunks = 1;
  if (chunks > c.tiles) chunks = c.tiles;
  c.tiles_per_chunk = ceil_div(c.tiles, chunks);
  c.chunks = ceil_div(c.tiles, c.tiles_per_chunk);
  return c;
}
constexpr int kSummaryTarget = 3 * kNumSMs;  // small shared-memory footprint: 3 CTAs / SM
constexpr int kTileTarget = kNumSMs;         // ~190 KB shared memory: 1 CTA / SM

// ------------------------------------------------------------------------------------------------
// K1  pixel-softmax summaries (flash style: running max / sum per (query, pixel slice))
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kNT) sql_summary_kernel(const float* __restrict__ x, const float* __restrict__ queries,
                                                          int Q, int n, int tiles_per_chunk,
                                                          float* __restrict__ partial /*[B][chunks][Q][E+2]*/) {
  constexpr int LDX = E + 4;
  constexpr int REC = E + 2;
  __shared__ __align__(16) float xt[kTP * LDX];
  extern __shared__ __align__(16) float merge[];  // [nsl][QB][REC]
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int QB = round_up(Q, 32);
  const int nsl = kNT / QB;
  const int q = threadIdx.x % QB, slice = threadIdx.x / QB;
  const bool active = slice < nsl && q < Q;
  float kr[E], acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    kr[e] = active ? __ldg(queries + ((size_t)b * Q + q) * E + e) : 0.f;
    acc[e] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const float* xb = x + (size_t)b * E * n;
  const int p_begin = chunk * tiles_per_chunk * kTP;
  const int p_end = min(n, p_begin + tiles_per_chunk * kTP);
  for (int p0 = p_begin; p0 < p_end; p0 += kTP) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < E * kTP; idx += kNT) {
      const int e = idx / kTP, p = idx - e * kTP;
      xt[p * LDX + e] = (p0 + p < n) ? __ldg(xb + (size_t)e * n + p0 + p) : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int np = min(kTP, p_end - p0);
    for (int p = slice; p < np; p += nsl) {
      float xv[E];
      const float4* row = reinterpret_cast<const float4*>(xt + p * LDX);
      float y = 0.f;
#pragma unroll
      for (int e4 = 0; e4 < E / 4; ++e4) {
        const float4 v = row[e4];
        xv[4 * e4] = v.x; xv[4 * e4 + 1] = v.y; xv[4 * e4 + 2] = v.z; xv[4 * e4 + 3] = v.w;
        y = fmaf(v.x, kr[4 * e4], y); y = fmaf(v.y, kr[4 * e4 + 1], y);
        y = fmaf(v.z, kr[4 * e4 + 2], y); y = fmaf(v.w, kr[4 * e4 + 3], y);
      }
      // lazy rescale: the running reference only moves when it is exceeded by more than 8 (exp args stay <= 8)
      if (y > m + 8.f) {
        const float s = expf(m - y);  // m = -inf on the first pixel -> 0
        l *= s;
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] *= s;
        m = y;
      }
      const float pe = expf(y - m);
      l += pe;
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(pe, xv[e], acc[e]);
    }
  }
  __syncthreads();
  if (active) {
    float* rec = merge + ((size_t)slice * QB + q) * REC;
    rec[0] = m; rec[1] = l;
#pragma unroll
    for (int e = 0; e < E; ++e) rec[2 + e] = acc[e];
  }
  __syncthreads();
  if (active && slice == 0) {
    float M = -INFINITY;
    for (int s = 0; s < nsl; ++s) M = fmaxf(M, merge[((size_t)s * QB + q) * REC]);
    float L = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    for (int s = 0; s < nsl; ++s) {
      const float* rec = merge + ((size_t)s * QB + q) * REC;
      const float wgt = rec[0] == -INFINITY ? 0.f : expf(rec[0] - M);
      L = fmaf(rec[1], wgt, L);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(rec[2 + e], wgt, acc[e]);
    }
    float* out = partial + (((size_t)b * chunks + chunk) * Q + q) * REC;
    out[0] = M; out[1] = L;
#pragma unroll
    for (int e = 0; e < E; ++e) out[2 + e] = acc[e];
  }
}

// summary[b,q,:] = sum_c acc_c e^{m_c - M} / L   (split-softmax combine);  also emits M and L
template <int E>
__global__ void sql_summary_combine_kernel(const float* __restrict__ partial, int Q, int chunks,
                                           float* __restrict__ summary, float* __restrict__ row_max,
                                           float* __restrict__ row_sum) {
  // one warp per (sample, query): lanes own the embedding channels, chunk records are read as coalesced rows,
  // eight chunks in flight per trip (the first version ran one THREAD per query: 12 x 64 threads walking 800
  // dependent loads each)
  constexpr int REC = E + 2;
  constexpr int EPL = (E + 31) / 32;   // channels per lane
  const int b = blockIdx.x, lane = threadIdx.x & 31;
  const int q = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const float* base = partial + ((size_t)b * chunks * Q + q) * REC;
  const size_t cstride = (size_t)Q * REC;
  float M = -INFINITY;
  for (int c = lane; c < chunks; c += 32) M = fmaxf(M, base[c * cstride]);
  M = warp_max(M);
  float L = 0.f, acc[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) acc[k] = 0.f;
  for (int c0 = 0; c0 < chunks; c0 += 8) {
    float m8[8], l8[8], a8[8][EPL];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool on = c0 + j < chunks;
      const float* rec = base + (size_t)(on ? c0 + j : 0) * cstride;
      m8[j] = on ? rec[0] : -INFINITY;
      l8[j] = on ? rec[1] : 0.f;
#pragma unroll
      for (int k = 0; k < EPL; ++k) a8[j][k] = (on && lane + 32 * k < E) ? rec[2 + lane + 32 * k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float wgt = m8[j] == -INFINITY ? 0.f : expf(m8[j] - M);
      L = fmaf(l8[j], wgt, L);
#pragma unroll
      for (int k = 0; k < EPL; ++k) acc[k] = fmaf(a8[j][k], wgt, acc[k]);
    }
  }
  const float inv = 1.f / L;
#pragma unroll
  for (int k = 0; k < EPL; ++k)
    if (lane + 32 * k < E) summary[((size_t)b * Q + q) * E + lane + 32 * k] = acc[k] * inv;
  if (lane == 0) {
    if (row_max) row_max[b * Q + q] = M;
    if (row_sum) row_sum[b * Q + q] = L;
  }
}

// ------------------------------------------------------------------------------------------------
// shared tile machinery of the pred / backward kernels
// ------------------------------------------------------------------------------------------------
template <int E>
struct TileSmem {
  float* xs;    // [E][kTP]
  float* Ks;    // [Qp][E]      rows >= Q are zero
  float* Wps;   // [D][Qp]      cols >= Q are zero
  float* bps;   // [kMaxD]
  float* cs;    // [kMaxD]      bin centres of this sample
  float* ys;    // [kTP][kLD]   y (later dy) ; cols >= Q are zero
  float* red;   // [3][4][kTP]
  static __host__ __device__ size_t floats(int Q, int D) {
    const int Qp = round_up(Q, 8);
    return (size_t)E * kTP + (size_t)Qp * E + (size_t)D * Qp + 2 * kMaxD + (size_t)kTP * kLD + 3 * 4 * kTP;
  }
  __device__ float* carve(float* base, int Q, int D) {
    const int Qp = round_up(Q, 8);
    xs = base; base += E * kTP;
    Ks = base; base += Qp * E;
    Wps = base; base += D * Qp;
    bps = base; base += kMaxD;
    cs = base; base += kMaxD;
    ys = base; base += kTP * kLD;
    red = base; base += 3 * 4 * kTP;
    return base;
  }
};

template <int E>
__device__ __forceinline__ void load_weights(const TileSmem<E>& s, const float* __restrict__ queries_b,
                                             const float* __restrict__ Wp, const float* __restrict__ bp,
                                             const float* __restrict__ centers_b, int Q, int D) {
  const int Qp = round_up(Q, 8);
  for (int idx = threadIdx.x; idx < Qp * E; idx += kNT) s.Ks[idx] = idx < Q * E ? __ldg(queries_b + idx) : 0.f;
  if (Wp) {
    for (int idx = threadIdx.x; idx < D * Qp; idx += kNT) {
      const int d = idx / Qp, q = idx - d * Qp;
      s.Wps[idx] = q < Q ? __ldg(Wp + (size_t)d * Q + q) : 0.f;
    }
    for (int d = threadIdx.x; d < kMaxD; d += kNT) {
      s.bps[d] = d < D ? __ldg(bp + d) : 0.f;
      s.cs[d] = d < D ? __ldg(centers_b + d) : 0.f;
    }
  }
  for (int idx = threadIdx.x; idx < kTP * kLD; idx += kNT) s.ys[idx] = 0.f;
}

template <int E>
__device__ __forceinline__ void load_x_tile(const TileSmem<E>& s, const float* __restrict__ xb, int n, int p0) {
  for (int idx = threadIdx.x; idx < E * kTP; idx += kNT) {
    const int e = idx / kTP, p = idx - e * kTP;
    s.xs[idx] = (p0 + p < n) ? __ldg(xb + (size_t)e * n + p0 + p) : 0.f;
  }
}

// ys[p][q] = sum_e xs[e][p] * Ks[q][e]     thread = (pixel p, q mod 4 == sub)
template <int E>
__device__ __forceinline__ void compute_y(const TileSmem<E>& s, int Q, int p, int sub) {
  const int Qp = round_up(Q, 8);
  float xr[E];
#pragma unroll
  for (int e = 0; e < E; ++e) xr[e] = s.xs[e * kTP + p];
  for (int q = sub; q < Qp; q += 4) {
    const float4* kr = reinterpret_cast<const float4*>(s.Ks + q * E);
    float acc = 0.f;
#pragma unroll
    for (int e4 = 0; e4 < E / 4; ++e4) {
      const float4 k = kr[e4];
      acc = fmaf(xr[4 * e4], k.x, acc); acc = fmaf(xr[4 * e4 + 1], k.y, acc);
      acc = fmaf(xr[4 * e4 + 2], k.z, acc); acc = fmaf(xr[4 * e4 + 3], k.w, acc);
    }
    s.ys[p * kLD + q] = acc;
  }
}

// logits -> softmax over d -> expected centre.  On return prob[j] (d = sub + 4j) holds softmax_d, and the
// return value is pred.  Contains __syncthreads (all threads must call).
template <int E>
__device__ __forceinline__ float compute_prob(const TileSmem<E>& s, int Q, int D, int p, int sub, float (&prob)[kNJ]) {
  const int Qp = round_up(Q, 8);
#pragma unroll
  for (int j = 0; j < kNJ; ++j) prob[j] = (sub + 4 * j < D) ? s.bps[sub + 4 * j] : 0.f;
  for (int q4 = 0; q4 < Qp; q4 += 4) {
    const float4 yv = *reinterpret_cast<const float4*>(s.ys + p * kLD + q4);
#pragma unroll
    for (int j = 0; j < kNJ; ++j) {
      if (sub + 4 * j < D) {
        const float4 w = *reinterpret_cast<const float4*>(s.Wps + (sub + 4 * j) * Qp + q4);
        prob[j] = fmaf(yv.x, w.x, prob[j]); prob[j] = fmaf(yv.y, w.y, prob[j]);
        prob[j] = fmaf(yv.z, w.z, prob[j]); prob[j] = fmaf(yv.w, w.w, prob[j]);
      }
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kNJ; ++j)
    if (sub + 4 * j < D) mx = fmaxf(mx, prob[j]);
  s.red[sub * kTP + p] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s.red[p], s.red[kTP + p]), fmaxf(s.red[2 * kTP + p], s.red[3 * kTP + p]));
  float se = 0.f, sc = 0.f;
#pragma unroll
  for (int j = 0; j < kNJ; ++j) {
    if (sub + 4 * j < D) {
      const float ev = expf(prob[j] - mx);
      prob[j] = ev;
      se += ev;
      sc = fmaf(ev, s.cs[sub + 4 * j], sc);
    } else {
      prob[j] = 0.f;
    }
  }
  s.red[(4 + sub) * kTP + p] = se;
  s.red[(8 + sub) * kTP + p] = sc;
  __syncthreads();
  se = (s.red[4 * kTP + p] + s.red[5 * kTP + p]) + (s.red[6 * kTP + p] + s.red[7 * kTP + p]);
  sc = (s.red[8 * kTP + p] + s.red[9 * kTP + p]) + (s.red[10 * kTP + p] + s.red[11 * kTP + p]);
  const float inv = 1.f / se;
#pragma unroll
  for (int j = 0; j < kNJ; ++j) prob[j] *= inv;
  return sc * inv;
}

// ------------------------------------------------------------------------------------------------
// K2  depth regression forward (+ optional energy map for the module-level FullQueryLayer)
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kNT) sql_pred_kernel(const float* __restrict__ x, const float* __restrict__ queries,
                                                       const float* __restrict__ Wp, const float* __restrict__ bp,
                                                       const float* __restrict__ centers, int Q, int D, int n,
                                                       int tiles_per_chunk, float* __restrict__ pred,
                                                       float* __restrict__ energy) {
  extern __shared__ __align__(16) float smem[];
  TileSmem<E> s;
  s.carve(smem, Q, D);
  const int b = blockIdx.y;
  const int p = threadIdx.x % kTP, sub = threadIdx.x / kTP;
  load_weights<E>(s, queries + (size_t)b * Q * E, Wp, bp, Wp ? centers + (size_t)b * D : nullptr, Q, D);
  const float* xb = x + (size_t)b * E * n;
  const int p_begin = blockIdx.x * tiles_per_chunk * kTP;
  const int p_end = min(n, p_begin + tiles_per_chunk * kTP);
  for (int p0 = p_begin; p0 < p_end; p0 += kTP) {
    __syncthreads();
    load_x_tile<E>(s, xb, n, p0);
    __syncthreads();
    compute_y<E>(s, Q, p, sub);
    __syncthreads();
    if (energy) {
      for (int q = sub; q < Q; q += 4)
        if (p0 + p < n) energy[((size_t)b * Q + q) * n + p0 + p] = s.ys[p * kLD + q];
    }
    if (Wp) {
      float prob[kNJ];
      const float pr = compute_prob<E>(s, Q, D, p, sub, prob);
      if (sub == 0 && p0 + p < n) pred[(size_t)b * n + p0 + p] = pr;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3  backward pass 1: pixel reductions  d_centers, d_Wp, d_bp   (per-CTA partials -> reduce kernel)
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kNT) sql_bwd_reduce_kernel(
    const float* __restrict__ x, const float* __restrict__ queries, const float* __restrict__ Wp,
    const float* __restrict__ bp, const float* __restrict__ centers, const float* __restrict__ g_pred, int Q, int D,
    int n, int tiles_per_chunk, float* __restrict__ part_dW /*[cta][D][Q]*/, float* __restrict__ part_dc /*[cta][D]*/,
    float* __restrict__ part_db /*[cta][D]*/) {
  extern __shared__ __align__(16) float smem[];
  TileSmem<E> s;
  float* dzs = s.carve(smem, Q, D);   // [kTP][kLD]
  float* wred = dzs + kTP * kLD;      // [8 warps][2][kNJ]
  const int b = blockIdx.y;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int p = threadIdx.x % kTP, sub = threadIdx.x / kTP;
  load_weights<E>(s, queries + (size_t)b * Q * E, Wp, bp, centers + (size_t)b * D, Q, D);
  for (int idx = threadIdx.x; idx < kTP * kLD; idx += kNT) dzs[idx] = 0.f;
  const float* xb = x + (size_t)b * E * n;
  const int p_begin = blockIdx.x * tiles_per_chunk * kTP;
  const int p_end = min(n, p_begin + tiles_per_chunk * kTP);
  float dc[kNJ], db[kNJ];
#pragma unroll
  for (int j = 0; j < kNJ; ++j) { dc[j] = 0.f; db[j] = 0.f; }
  // dW micro-tile: d0 = 8*(tid/16), q0 = 8*(tid%16)
  const int d0 = 8 * (threadIdx.x >> 4), q0 = 8 * (threadIdx.x & 15);
  float dW[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dW[i][j] = 0.f;
  for (int p0 = p_begin; p0 < p_end; p0 += kTP) {
    __syncthreads();
    load_x_tile<E>(s, xb, n, p0);
    __syncthreads();
    compute_y<E>(s, Q, p, sub);
    __syncthreads();
    float prob[kNJ];
    const float pr = compute_prob<E>(s, Q, D, p, sub, prob);
    const float g = (p0 + p < n) ? __ldg(g_pred + (size_t)b * n + p0 + p) : 0.f;
#pragma unroll
    for (int j = 0; j < kNJ; ++j) {
      if (sub + 4 * j < D) {
        const float pg = prob[j] * g;
        const float dz = pg * (s.cs[sub + 4 * j] - pr);
        dc[j] += pg;
        db[j] += dz;
        dzs[p * kLD + sub + 4 * j] = dz;
      }
    }
    __syncthreads();
    if (d0 < D && q0 < Q) {
      for (int pp = 0; pp < kTP; ++pp) {
        const float4 a0 = *reinterpret_cast<const float4*>(dzs + pp * kLD + d0);
        const float4 a1 = *reinterpret_cast<const float4*>(dzs + pp * kLD + d0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s.ys + pp * kLD + q0);
        const float4 b1 = *reinterpret_cast<const float4*>(s.ys + pp * kLD + q0 + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dW[i][j] = fmaf(av[i], bv[j], dW[i][j]);
      }
    }
  }
  // per-CTA partials
  if (d0 < D && q0 < Q) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (d0 + i < D && q0 + j < Q) part_dW[((size_t)cta * D + d0 + i) * Q + q0 + j] = dW[i][j];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kNJ; ++j) {
    const float a = warp_sum(dc[j]), c = warp_sum(db[j]);
    if (lane == 0) { wred[(warp * 2 + 0) * kNJ + j] = a; wred[(warp * 2 + 1) * kNJ + j] = c; }
  }
  __syncthreads();
  // warps 2*sub and 2*sub+1 hold the two pixel halves of d = sub + 4j
  for (int d = threadIdx.x; d < D; d += kNT) {
    const int sb = d & 3, j = d >> 2;
    part_dc[(size_t)cta * D + d] = wred[((2 * sb) * 2 + 0) * kNJ + j] + wred[((2 * sb + 1) * 2 + 0) * kNJ + j];
    part_db[(size_t)cta * D + d] = wred[((2 * sb) * 2 + 1) * kNJ + j] + wred[((2 * sb + 1) * 2 + 1) * kNJ + j];
  }
}

// The three partial-sum reductions after the regression backward in ONE launch (grid (ceil(D*32/256), B, 3)):
//   z = 0: d_M[b][i]  = sum_{c < chunks} part_dM[(b*chunks + c) * D*32 + i]
//   z = 1: d_c[b][d]  = sum_{c < chunks} part_dc[(b*chunks + c) * D + d]
//   z = 2: d_b[d]     = sum_{c < B*chunks} part_db[c * D + d]   (one block; two halves of the CTAs, then a fixed-order add)
__global__ void __launch_bounds__(256) sum_partials3_kernel(const float* __restrict__ part_dM, const float* __restrict__ part_dc,
                                                            const float* __restrict__ part_db, int chunks, int B, int D,
                                                            float* __restrict__ d_M, float* __restrict__ d_c,
                                                            float* __restrict__ d_b) {
  __shared__ float half[256];
  const int b = blockIdx.y, t = threadIdx.x;
  if (blockIdx.z == 0) {
    const int stride = D * 32, i = blockIdx.x * 256 + t;
    if (i >= stride) return;
    const float* src = part_dM + (size_t)b * chunks * stride + i;
    float acc = 0.f;
    int c = 0;
    for (; c + 8 <= chunks; c += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = src[(size_t)(c + j) * stride];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    for (; c < chunks; ++c) acc += src[(size_t)c * stride];
    d_M[(size_t)b * stride + i] = acc;
  } else if (blockIdx.z == 1) {
    if (blockIdx.x != 0 || t >= D) return;
    const float* src = part_dc + (size_t)b * chunks * D + t;
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += src[(size_t)c * D];
    d_c[(size_t)b * D + t] = acc;
  } else {
    if (blockIdx.x != 0 || b != 0) return;
    const int d = t & 127, g = t >> 7, ctas = B * chunks;      // D <= 128: two groups of 128 threads
    const int c0 = g * (ctas / 2), c1 = g ? ctas : ctas / 2;
    float acc = 0.f;
    if (d < D) {
      int c = c0;
      for (; c + 8 <= c1; c += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = part_db[(size_t)(c + j) * D + d];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
      }
      for (; c < c1; ++c) acc += part_db[(size_t)c * D + d];
    }
    half[t] = acc;
    __syncthreads();
    if (g == 0 && d < D) d_b[d] = half[d] + half[128 + d];
  }
}

// out[i] = sum_{c < count} part[(c0 + c) * stride + i]     grid.y selects the group (c0 = group * count)
__global__ void sum_partials_kernel(const float* __restrict__ part, int count, int stride, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride) return;
  const float* src = part + (size_t)blockIdx.y * count * stride + i;
  float acc = 0.f;
  int c = 0;
  for (; c + 8 <= count; c += 8) {      // eight independent loads per trip, summed in index order
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = src[(size_t)(c + j) * stride];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  for (; c < count; ++c) acc += src[(size_t)c * stride];
  out[(size_t)blockIdx.y * stride + i] = acc;
}

// ------------------------------------------------------------------------------------------------
// K4  backward pass 2:  d_x, d_queries
// ------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(kNT) sql_bwd_dx_kernel(
    const float* __restrict__ x, const float* __restrict__ queries, const float* __restrict__ Wp,
    const float* __restrict__ bp, const float* __restrict__ centers, const float* __restrict__ g_pred,
    const float* __restrict__ summary, const float* __restrict__ row_max, const float* __restrict__ row_sum,
    const float* __restrict__ d_summary, const float* __restrict__ g_energy, int Q, int D, int n, int tiles_per_chunk,
    float* __restrict__ d_x, float* __restrict__ part_dK /*[cta][Q][E]*/) {
  extern __shared__ __align__(16) float smem[];
  TileSmem<E> s;
  const int Dw = Wp ? D : 0;
  float* aux = s.carve(smem, Q, Dw);
  const int Qp = round_up(Q, 8);
  float* dzs = aux;                   // [kTP][kLD]  dz, later a = pixel-softmax weights
  float* dsS = dzs + kTP * kLD;       // [Qp][E]     d_summary
  float* dsT = dsS + Qp * E;          // [E][Qp]     d_summary transposed
  float* mq = dsT + E * Qp;           // [kMaxQ] row max
  float* il = mq + kMaxQ;             // [kMaxQ] 1 / row sum
  float* dl = il + kMaxQ;             // [kMaxQ] delta_q = sum_e d_summary * summary
  const int b = blockIdx.y;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int p = threadIdx.x % kTP, sub = threadIdx.x / kTP;
  const bool has_pred = Wp != nullptr, has_sum = d_summary != nullptr;
  load_weights<E>(s, queries + (size_t)b * Q * E, Wp, bp, has_pred ? centers + (size_t)b * D : nullptr, Q, D);
  for (int idx = threadIdx.x; idx < kTP * kLD; idx += kNT) dzs[idx] = 0.f;
  for (int idx = threadIdx.x; idx < Qp * E; idx += kNT) {
    const int q = idx / E, e = idx - q * E;
    const float v = (has_sum && q < Q) ? __ldg(d_summary + ((size_t)b * Q + q) * E + e) : 0.f;
    dsS[idx] = v;
    dsT[e * Qp + q] = v;
  }
  for (int q = threadIdx.x; q < kMaxQ; q += kNT) {
    float m = 0.f, inv = 0.f, delta = 0.f;
    if (has_sum && q < Q) {
      m = __ldg(row_max + b * Q + q);
      inv = 1.f / __ldg(row_sum + b * Q + q);
      for (int e = 0; e < E; ++e)
        delta = fmaf(__ldg(d_summary + ((size_t)b * Q + q) * E + e), __ldg(summary + ((size_t)b * Q + q) * E + e), delta);
    }
    mq[q] = m; il[q] = inv; dl[q] = delta;
  }
  const float* xb = x + (size_t)b * E * n;
  float* dxb = d_x + (size_t)b * E * n;
  const int p_begin = blockIdx.x * tiles_per_chunk * kTP;
  const int p_end = min(n, p_begin + tiles_per_chunk * kTP);
  // dy micro-tile: 4 pixels x 8 queries
  const int mp0 = 4 * (threadIdx.x >> 4), mq0 = 8 * (threadIdx.x & 15);
  // dK micro-tile: 4 queries x (E/8) channels
  constexpr int EK = E / 8;
  const int kq0 = 4 * (threadIdx.x >> 3), ke0 = EK * (threadIdx.x & 7);
  float dK[4][EK];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < EK; ++j) dK[i][j] = 0.f;
  // dx: thread = (pixel p, channels sub*E/4 .. +E/4)
  constexpr int EX = E / 4;
  const int xe0 = EX * sub;

  for (int p0 = p_begin; p0 < p_end; p0 += kTP) {
    __syncthreads();
    load_x_tile<E>(s, xb, n, p0);
    __syncthreads();
    compute_y<E>(s, Q, p, sub);
    __syncthreads();
    if (has_pred) {
      float prob[kNJ];
      const float pr = compute_prob<E>(s, Q, D, p, sub, prob);
      const float g = (p0 + p < n) ? __ldg(g_pred + (size_t)b * n + p0 + p) : 0.f;
#pragma unroll
      for (int j = 0; j < kNJ; ++j)
        if (sub + 4 * j < D) dzs[p * kLD + sub + 4 * j] = prob[j] * g * (s.cs[sub + 4 * j] - pr);
      __syncthreads();
    }
    // ---- dy[p][q] = sum_d dz[p][d] Wp[d][q] + a[p][q] (sum_e ds[q][e] x[e][p] - delta_q) + g_energy
    float dy[4][8], av[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { dy[i][j] = 0.f; av[i][j] = 0.f; }
    const bool qa = mq0 < Qp;
    if (has_pred && qa) {
      for (int d = 0; d < D; ++d) {
        const float4 w0 = *reinterpret_cast<const float4*>(s.Wps + d * Qp + mq0);
        const float4 w1 = *reinterpret_cast<const float4*>(s.Wps + d * Qp + mq0 + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = dzs[(mp0 + i) * kLD + d];
#pragma unroll
          for (int j = 0; j < 8; ++j) dy[i][j] = fmaf(a, wv[j], dy[i][j]);
        }
      }
    }
    if (has_sum && qa) {
      float t[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) t[i][j] = 0.f;
#pragma unroll 4
      for (int e = 0; e < E; ++e) {
        const float4 xv = *reinterpret_cast<const float4*>(s.xs + e * kTP + mp0);
        const float4 w0 = *reinterpret_cast<const float4*>(dsT + e * Qp + mq0);
        const float4 w1 = *reinterpret_cast<const float4*>(dsT + e * Qp + mq0 + 4);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) t[i][j] = fmaf(xa[i], wv[j], t[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int q = mq0 + j;
          const float a = (q < Q) ? expf(s.ys[(mp0 + i) * kLD + q] - mq[q]) * il[q] : 0.f;
          av[i][j] = a;
          dy[i][j] = fmaf(a, t[i][j] - dl[q < kMaxQ ? q : 0], dy[i][j]);
        }
    }
    if (g_energy && qa) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (mq0 + j < Q && p0 + mp0 + i < n) dy[i][j] += __ldg(g_energy + ((size_t)b * Q + mq0 + j) * n + p0 + mp0 + i);
    }
    __syncthreads();  // everyone is done reading dzs (dz) and ys (y)
    if (qa) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool live = p0 + mp0 + i < n;  // padded pixels must not leak into dK
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s.ys[(mp0 + i) * kLD + mq0 + j] = (live && mq0 + j < Q) ? dy[i][j] : 0.f;
          dzs[(mp0 + i) * kLD + mq0 + j] = (live && mq0 + j < Q) ? av[i][j] : 0.f;
        }
      }
    }
    __syncthreads();
    // ---- dx[e][p] = sum_q dy[p][q] K[q][e] + a[p][q] ds[q][e]
    {
      float dx[EX];
#pragma unroll
      for (int k = 0; k < EX; ++k) dx[k] = 0.f;
      for (int q = 0; q < Q; ++q) {
        const float dyv = s.ys[p * kLD + q];
        const float aq = dzs[p * kLD + q];
#pragma unroll
        for (int k4 = 0; k4 < EX / 4; ++k4) {
          const float4 kv = *reinterpret_cast<const float4*>(s.Ks + q * E + xe0 + 4 * k4);
          const float4 sv = *reinterpret_cast<const float4*>(dsS + q * E + xe0 + 4 * k4);
          dx[4 * k4] = fmaf(dyv, kv.x, fmaf(aq, sv.x, dx[4 * k4]));
          dx[4 * k4 + 1] = fmaf(dyv, kv.y, fmaf(aq, sv.y, dx[4 * k4 + 1]));
          dx[4 * k4 + 2] = fmaf(dyv, kv.z, fmaf(aq, sv.z, dx[4 * k4 + 2]));
          dx[4 * k4 + 3] = fmaf(dyv, kv.w, fmaf(aq, sv.w, dx[4 * k4 + 3]));
        }
      }
      if (p0 + p < n) {
#pragma unroll
        for (int k = 0; k < EX; ++k) dxb[(size_t)(xe0 + k) * n + p0 + p] = dx[k];
      }
    }
    // ---- dK[q][e] += sum_p dy[p][q] x[e][p]
    if (kq0 < Qp) {
      for (int pp = 0; pp < kTP; ++pp) {
        const float4 dv = *reinterpret_cast<const float4*>(s.ys + pp * kLD + kq0);
        const float dq[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < EK; ++j) {
          const float xv = s.xs[(ke0 + j) * kTP + pp];
#pragma unroll
          for (int i = 0; i < 4; ++i) dK[i][j] = fmaf(dq[i], xv, dK[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < EK; ++j)
      if (kq0 + i < Q) part_dK[((size_t)cta * Q + kq0 + i) * E + ke0 + j] = dK[i][j];
}

}  // namespace sqlx

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace sqlx;

namespace sqlx {
void tc_summary_plan(int B, int n, int* chunks, int* tiles_per_chunk);
int tc_summary_partials(const float* x, const float* queries, int B, int Q, int n, float* partial, int* chunks_out,
                        cudaStream_t st);
int tc_pred_mix_fwd(const float* x, const float* Mx, const float* bp, const float* centers, int B, int D, int n,
                    float* pred, cudaStream_t st);
int tc_bwd_pred_mix(const float* x, const float* Mx, const float* bp, const float* centers, const float* g_pred, int B,
                    int D, int n, float* d_x, float* part_dM, float* part_db, float* part_dc, int chunks, int tpc,
                    cudaStream_t st);
int tc_bwd_sum(const float* x, const float* queries, const float* summary, const float* row_max, const float* row_sum,
               const float* d_summary, int B, int Q, int n, int accumulate, float* d_x, float* part_dK, int chunks, int tpc,
               cudaStream_t st);
void tc_bwd_plan(int B, int n, int* chunks, int* tiles_per_chunk);
// warp-specialised generation (sql_ws.cu)
void ws_plan(int B, int n, int* chunks, int* tiles_per_chunk);
void ws_summary_plan(int B, int n, int* chunks, int* steps_per_chunk);
int ws_summary_partials(const float* x, const float* queries, int B, int Q, int n, float* partial, int* chunks_out,
                        cudaStream_t st);
int ws_pred_fwd(const float* x, const float* Mx, const float* bp, const float* centers, int B, int D, int n, float* pred,
                float* stat_m, float* stat_inv, cudaStream_t st);
int ws_bwd_sum(const float* x, const float* queries, const float* summary, const float* row_max, const float* row_sum,
               const float* d_summary, int B, int Q, int n, int accumulate, float* d_x, float* part_dK, int chunks, int tpc,
               cudaStream_t st);
int ws_bwd_pred(const float* x, const float* Mx, const float* bp, const float* centers, const float* g_pred,
                const float* pred, const float* stat_m, const float* stat_inv, int B, int D, int n, float* d_x,
                float* part_dM, float* part_db, float* part_dc, int chunks, int tpc, cudaStream_t st);
}
extern "C" int sqlx_sql_tc_supported(int E, int Q, int D, int n);
extern "C" int sqlx_sql_set_tensor_cores(int on);
extern "C" int sqlx_sql_energy_tc(const float* x, const float* queries, int B, int E, int Q, int n, float* energy,
                                  void* stream);

namespace {

std::atomic<int> g_use_tc{1};   // sqlx_sql_set_tensor_cores(0) forces the exact-fp32 CUDA-core kernels (tests, A/B timing)
bool use_tensor_cores(int E, int Q, int D, int n) { return g_use_tc && sqlx_sql_tc_supported(E, Q, D, n); }

int check_sql_shape(int B, int E, int Q, int D, int n, bool need_d) {
  SQLX_REQUIRE(B > 0 && n > 0, "non-positive shape (B=%d, n=%d)", B, n);
  SQLX_REQUIRE(B <= 65535, "B=%d too large for one launch", B);
  SQLX_REQUIRE(E == 16 || E == 32 || E == 64, "embedding dim E=%d unsupported (16, 32 or 64)", E);
  SQLX_REQUIRE(Q >= 1 && Q <= kMaxQ, "query_nums Q=%d outside 1..%d", Q, kMaxQ);
  SQLX_REQUIRE(!need_d || (D >= 1 && D <= kMaxD), "dim_out D=%d outside 1..%d", D, kMaxD);
  return SQLX_OK;
}

template <typename F>
int set_smem(F kern, size_t bytes) {
  SQLX_REQUIRE(bytes <= 227 * 1024, "shape needs %zu bytes of shared memory per CTA (limit 232448): reduce E, Q or D",
               bytes);
  if (bytes > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return check_launch("cudaFuncSetAttribute");
  }
  return SQLX_OK;
}

size_t summary_ws_floats(int B, int E, int Q, int n) {
  const ChunkPlan c = plan_chunks(B, n, kSummaryTarget);
  int tc_chunks = 0, tpc = 0;
  tc_summary_plan(B, n, &tc_chunks, &tpc);
  int ws_chunks = 0;
  ws_summary_plan(B, n, &ws_chunks, &tpc);
  int chunks = c.chunks > tc_chunks ? c.chunks : tc_chunks;
  chunks = chunks > ws_chunks ? chunks : ws_chunks;
  return (size_t)B * chunks * Q * (E + 2);
}
int max_bwd_chunks(int B, int n) {
  const ChunkPlan c = plan_chunks(B, n, kTileTarget);
  int tc_chunks = 0, tpc = 0;
  tc_bwd_plan(B, n, &tc_chunks, &tpc);
  return c.chunks > tc_chunks ? c.chunks : tc_chunks;
}
size_t reduce_ws_floats(int B, int Q, int D, int n) {
  return (size_t)B * max_bwd_chunks(B, n) * ((size_t)D * Q + 2 * D) + (size_t)2 * D;
}
size_t dx_ws_floats(int B, int E, int Q, int n) { return (size_t)B * max_bwd_chunks(B, n) * Q * E; }

template <int E>
int run_summary(const float* x, const float* queries, int B, int Q, int n, float* summary, float* row_max,
                float* row_sum, float* energy, float* ws, cudaStream_t st) {
  const ChunkPlan c = plan_chunks(B, n, kSummaryTarget);
  const int QB = round_up(Q, 32), nsl = kNT / QB;
  const size_t smem = sizeof(float) * (size_t)nsl * QB * (E + 2);
  if (int e = set_smem(sql_summary_kernel<E>, smem)) return e;
  {
    ProfScope prof("sql_summary_kernel", st);
    sql_summary_kernel<E><<<dim3(c.chunks, B), kNT, smem, st>>>(x, queries, Q, n, c.tiles_per_chunk, ws);
  }
  if (int e = check_launch("sql_summary_kernel")) return e;
  sql_summary_combine_kernel<E><<<dim3(B, (Q + 3) / 4), 128, 0, st>>>(ws, Q, c.chunks, summary, row_max, row_sum);
  if (int e = check_launch("sql_summary_combine_kernel")) return e;
  if (energy) {
    const ChunkPlan t = plan_chunks(B, n, 2 * kTileTarget);
    const size_t sm2 = sizeof(float) * TileSmem<E>::floats(Q, 0);
    if (int e = set_smem(sql_pred_kernel<E>, sm2)) return e;
    sql_pred_kernel<E><<<dim3(t.chunks, B), kNT, sm2, st>>>(x, queries, nullptr, nullptr, nullptr, Q, 0, n,
                                                            t.tiles_per_chunk, nullptr, energy);
    if (int e = check_launch("sql_pred_kernel(energy)")) return e;
  }
  return SQLX_OK;
}

template <int E>
int run_pred(const float* x, const float* queries, const float* Wp, const float* bp, const float* centers, int B,
             int Q, int D, int n, float* pred, cudaStream_t st) {
  const ChunkPlan c = plan_chunks(B, n, 2 * kTileTarget);
  const size_t smem = sizeof(float) * TileSmem<E>::floats(Q, D);
  if (int e = set_smem(sql_pred_kernel<E>, smem)) return e;
  ProfScope prof("sql_pred_kernel", st);
  sql_pred_kernel<E><<<dim3(c.chunks, B), kNT, smem, st>>>(x, queries, Wp, bp, centers, Q, D, n, c.tiles_per_chunk,
                                                           pred, nullptr);
  return check_launch("sql_pred_kernel");
}

template <int E>
int run_bwd_reduce(const float* x, const float* queries, const float* Wp, const float* bp, const float* centers,
                   const float* g_pred, int B, int Q, int D, int n, float* d_centers, float* d_Wp, float* d_bp,
                   float* ws, cudaStream_t st) {
  const ChunkPlan c = plan_chunks(B, n, kTileTarget);
  const int ctas = B * c.chunks;
  float* part_dW = ws;
  float* part_dc = part_dW + (size_t)ctas * D * Q;
  float* part_db = part_dc + (size_t)ctas * D;
  const size_t smem = sizeof(float) * (TileSmem<E>::floats(Q, D) + (size_t)kTP * kLD + 8 * 2 * kNJ);
  if (int e = set_smem(sql_bwd_reduce_kernel<E>, smem)) return e;
  {
    ProfScope prof("sql_bwd_reduce_kernel", st);
    sql_bwd_reduce_kernel<E><<<dim3(c.chunks, B), kNT, smem, st>>>(x, queries, Wp, bp, centers, g_pred, Q, D, n,
                                                                   c.tiles_per_chunk, part_dW, part_dc, part_db);
  }
  if (int e = check_launch("sql_bwd_reduce_kernel")) return e;
  sum_partials_kernel<<<dim3(ceil_div(D * Q, 256), 1), 256, 0, st>>>(part_dW, ctas, D * Q, d_Wp);
  if (int e = check_launch("sum_partials_kernel")) return e;
  sum_partials_kernel<<<dim3(ceil_div(D, 256), 1), 256, 0, st>>>(part_db, ctas, D, d_bp);
  if (int e = check_launch("sum_partials_kernel")) return e;
  sum_partials_kernel<<<dim3(ceil_div(D, 256), B), 256, 0, st>>>(part_dc, c.chunks, D, d_centers);
  return check_launch("sum_partials_kernel");
}

template <int E>
int run_bwd_dx(const float* x, const float* queries, const float* Wp, const float* bp, const float* centers,
               const float* g_pred, const float* summary, const float* row_max, const float* row_sum,
               const float* d_summary, const float* g_energy, int B, int Q, int D, int n, float* d_x, float* d_queries,
               float* ws, cudaStream_t st) {
  const ChunkPlan c = plan_chunks(B, n, kTileTarget);
  const int Qp = round_up(Q, 8);
  const size_t smem = sizeof(float) * (TileSmem<E>::floats(Q, Wp ? D : 0) + (size_t)kTP * kLD + 2 * (size_t)Qp * E + 3 * kMaxQ);
  if (int e = set_smem(sql_bwd_dx_kernel<E>, smem)) return e;
  {
    ProfScope prof("sql_bwd_dx_kernel", st);
    sql_bwd_dx_kernel<E><<<dim3(c.chunks, B), kNT, smem, st>>>(x, queries, Wp, bp, centers, g_pred, summary, row_max,
                                                               row_sum, d_summary, g_energy, Q, D, n,
                                                               c.tiles_per_chunk, d_x, ws);
  }
  if (int e = check_launch("sql_bwd_dx_kernel")) return e;
  sum_partials_kernel<<<dim3(ceil_div(Q * E, 256), B), 256, 0, st>>>(ws, c.chunks, Q * E, d_queries);
  return check_launch("sum_partials_kernel");
}

#define SQLX_DISPATCH_E(E_, CALL)                 \
  switch (E_) {                                   \
    case 16: { constexpr int kE = 16; return CALL; } \
    case 32: { constexpr int kE = 32; return CALL; } \
    default: { constexpr int kE = 64; return CALL; } \
  }

}  // namespace

extern "C" int sqlx_sql_set_tensor_cores(int on) { return g_use_tc.exchange(on ? 1 : 0); }
extern "C" int sqlx_sql_get_tensor_cores(void) { return g_use_tc.load(); }

extern "C" size_t sqlx_sql_workspace_bytes(int B, int E, int Q, int D, int n) {
  if (B <= 0 || E <= 0 || Q <= 0 || D < 0 || n <= 0) return 0;
  size_t f = summary_ws_floats(B, E, Q, n);
  const size_t r = reduce_ws_floats(B, Q, D, n), x = dx_ws_floats(B, E, Q, n);
  if (r > f) f = r;
  if (x > f) f = x;
  return sizeof(float) * (f + 64);
}

extern "C" int sqlx_sql_summary_fwd(const float* x, const float* queries, int B, int E, int Q, int n, float* summary,
                                    float* row_max, float* row_sum, float* energy, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (int e = check_sql_shape(B, E, Q, 0, n, false)) return e;
  SQLX_REQUIRE(x && queries && summary, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sizeof(float) * summary_ws_floats(B, E, Q, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* ws = reinterpret_cast<float*>(workspace);
  if (use_tensor_cores(E, Q, 0, n)) {
    int chunks = 0;
    if (int e = ws_summary_partials(x, queries, B, Q, n, ws, &chunks, st)) return e;
    sql_summary_combine_kernel<32><<<dim3(B, (Q + 3) / 4), 128, 0, st>>>(ws, Q, chunks, summary, row_max, row_sum);
    if (int e = check_launch("sql_summary_combine_kernel")) return e;
    if (energy) return sqlx_sql_energy_tc(x, queries, B, E, Q, n, energy, stream);
    return SQLX_OK;
  }
  SQLX_DISPATCH_E(E, run_summary<kE>(x, queries, B, Q, n, summary, row_max, row_sum, energy, ws, st));
}

/* round-1 generation of the tensor-core summary kernel (one warpgroup, serial phases): A/B and cross-check of the
 * warp-specialised kernel (tests/test_sql_tc_gpu.py::test_ws_kernels_match_v1) */
extern "C" int sqlx_sql_summary_fwd_v1(const float* x, const float* queries, int B, int E, int Q, int n, float* summary,
                                       float* row_max, float* row_sum, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  if (int e = check_sql_shape(B, E, Q, 0, n, false)) return e;
  SQLX_REQUIRE(x && queries && summary, "NULL pointer argument");
  SQLX_REQUIRE(use_tensor_cores(E, Q, 0, n), "shape E=%d Q=%d n=%d is not supported by the tensor-core path", E, Q, n);
  SQLX_REQUIRE(workspace && workspace_bytes >= sizeof(float) * summary_ws_floats(B, E, Q, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* ws = reinterpret_cast<float*>(workspace);
  int chunks = 0;
  if (int e = tc_summary_partials(x, queries, B, Q, n, ws, &chunks, st)) return e;
  sql_summary_combine_kernel<32><<<dim3(B, (Q + 3) / 4), 128, 0, st>>>(ws, Q, chunks, summary, row_max, row_sum);
  return check_launch("sql_summary_combine_kernel");
}

extern "C" int sqlx_sql_pred_fwd(const float* x, const float* queries, const float* Wp, const float* bp,
                                 const float* centers, int B, int E, int Q, int D, int n, float* pred, void* stream) {
  if (int e = check_sql_shape(B, E, Q, D, n, true)) return e;
  SQLX_REQUIRE(x && queries && Wp && bp && centers && pred, "NULL pointer argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // (un-mixed formulation: exact-fp32 CUDA-core kernel for every shape; the tensor-core path is sqlx_sql_pred_mix_fwd)
  SQLX_DISPATCH_E(E, run_pred<kE>(x, queries, Wp, bp, centers, B, Q, D, n, pred, st));
}

extern "C" int sqlx_sql_bwd_reduce(const float* x, const float* queries, const float* Wp, const float* bp,
                                   const float* centers, const float* pred, const float* g_pred, int B, int E, int Q,
                                   int D, int n, float* d_centers, float* d_Wp, float* d_bp, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  (void)pred;  // recomputed on chip: cheaper than reading it back
  if (int e = check_sql_shape(B, E, Q, D, n, true)) return e;
  SQLX_REQUIRE(x && queries && Wp && bp && centers && g_pred && d_centers && d_Wp && d_bp, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sizeof(float) * reduce_ws_floats(B, Q, D, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* ws = reinterpret_cast<float*>(workspace);
  SQLX_DISPATCH_E(E, run_bwd_reduce<kE>(x, queries, Wp, bp, centers, g_pred, B, Q, D, n, d_centers, d_Wp, d_bp, ws, st));
}

extern "C" int sqlx_sql_bwd_dx(const float* x, const float* queries, const float* Wp, const float* bp,
                               const float* centers, const float* pred, const float* g_pred, const float* summary,
                               const float* row_max, const float* row_sum, const float* d_summary,
                               const float* g_energy, int B, int E, int Q, int D, int n, float* d_x, float* d_queries,
                               void* workspace, size_t workspace_bytes, void* stream) {
  (void)pred;
  const bool has_pred = g_pred != nullptr;
  if (int e = check_sql_shape(B, E, Q, D, n, has_pred)) return e;
  SQLX_REQUIRE(x && queries && d_x && d_queries, "NULL pointer argument");
  SQLX_REQUIRE(!has_pred || (Wp && bp && centers), "g_pred given without Wp / bp / centers");
  SQLX_REQUIRE(!d_summary || (summary && row_max && row_sum), "d_summary given without summary / row_max / row_sum");
  SQLX_REQUIRE(workspace && workspace_bytes >= sizeof(float) * dx_ws_floats(B, E, Q, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* ws = reinterpret_cast<float*>(workspace);
  if (!has_pred) { Wp = nullptr; bp = nullptr; centers = nullptr; }
  SQLX_DISPATCH_E(E, run_bwd_dx<kE>(x, queries, Wp, bp, centers, g_pred, summary, row_max, row_sum, d_summary, g_energy,
                                    B, Q, D, n, d_x, d_queries, ws, st));
}

// ------------------------------------------------------------------------------------------------
// mixed-weight decomposition (tensor-core only):  logits = (Wp K) x + b = M x + b   -- see sql_tc.cu
// ------------------------------------------------------------------------------------------------
extern "C" size_t sqlx_sql_mix_workspace_bytes(int B, int Q, int D, int n) {
  if (B <= 0 || n <= 0) return 0;
  int chunks = 0, tpc = 0;
  tc_bwd_plan(B, n, &chunks, &tpc);
  const size_t ctas = (size_t)B * chunks;
  const size_t a = ctas * ((size_t)D * 32 + 2 * D), b = ctas * (size_t)Q * 32;
  return sizeof(float) * ((a > b ? a : b) + 64);
}

extern "C" int sqlx_sql_pred_mix_fwd(const float* x, const float* Mx, const float* bp, const float* centers, int B, int E,
                                     int D, int n, float* pred, float* stats, void* stream) {
  SQLX_REQUIRE(x && Mx && bp && centers && pred && stats, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, 1, D, n), "shape E=%d D=%d n=%d is not supported by the tensor-core path", E, D, n);
  return ws_pred_fwd(x, Mx, bp, centers, B, D, n, pred, stats, stats + (size_t)B * n, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sqlx_sql_bwd_pred_mix(const float* x, const float* Mx, const float* bp, const float* centers,
                                     const float* g_pred, const float* pred, const float* stats, int B, int E, int D, int n,
                                     float* d_M, float* d_bp, float* d_centers, float* d_x, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(x && Mx && bp && centers && g_pred && pred && stats && d_M && d_bp && d_centers && d_x && workspace,
               "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, 1, D, n), "shape E=%d D=%d n=%d is not supported by the tensor-core path", E, D, n);
  SQLX_REQUIRE(workspace_bytes >= sqlx_sql_mix_workspace_bytes(B, 1, D, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int chunks = 0, tpc = 0;
  ws_plan(B, n, &chunks, &tpc);
  const int ctas = B * chunks;   // partial arrays are [ctas][...]
  float* part_dM = reinterpret_cast<float*>(workspace);
  float* part_db = part_dM + (size_t)ctas * D * 32;
  float* part_dc = part_db + (size_t)ctas * D;
  if (int e = ws_bwd_pred(x, Mx, bp, centers, g_pred, pred, stats, stats + (size_t)B * n, B, D, n, d_x, part_dM, part_db, part_dc,
                          chunks, tpc, st))
    return e;
  sum_partials3_kernel<<<dim3(ceil_div(D * 32, 256), B, 3), 256, 0, st>>>(part_dM, part_dc, part_db, chunks, B, D, d_M, d_centers,
                                                                         d_bp);
  return check_launch("sum_partials3_kernel");
}

/* round-1 generation of the two kernels above (single warpgroup, serial phases): kept for one round as the A/B and
 * cross-check of the warp-specialised kernels (tests/test_sql_tc_gpu.py::test_ws_kernels_match_v1) */
extern "C" int sqlx_sql_pred_mix_fwd_v1(const float* x, const float* Mx, const float* bp, const float* centers, int B, int E,
                                        int D, int n, float* pred, void* stream) {
  SQLX_REQUIRE(x && Mx && bp && centers && pred, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, 1, D, n), "shape E=%d D=%d n=%d is not supported by the tensor-core path", E, D, n);
  return tc_pred_mix_fwd(x, Mx, bp, centers, B, D, n, pred, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sqlx_sql_bwd_pred_mix_v1(const float* x, const float* Mx, const float* bp, const float* centers,
                                        const float* g_pred, int B, int E, int D, int n, float* d_M, float* d_bp,
                                        float* d_centers, float* d_x, void* workspace, size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(x && Mx && bp && centers && g_pred && d_M && d_bp && d_centers && d_x && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, 1, D, n), "shape E=%d D=%d n=%d is not supported by the tensor-core path", E, D, n);
  SQLX_REQUIRE(workspace_bytes >= sqlx_sql_mix_workspace_bytes(B, 1, D, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int chunks = 0, tpc = 0;
  tc_bwd_plan(B, n, &chunks, &tpc);
  const int ctas = B * chunks;   // partial arrays are [ctas][...]
  float* part_dM = reinterpret_cast<float*>(workspace);
  float* part_db = part_dM + (size_t)ctas * D * 32;
  float* part_dc = part_db + (size_t)ctas * D;
  if (int e = tc_bwd_pred_mix(x, Mx, bp, centers, g_pred, B, D, n, d_x, part_dM, part_db, part_dc, chunks, tpc, st)) return e;
  SQLX_REQUIRE(D <= 128, "dim_out %d exceeds the tensor-core path's limit", D);
  sum_partials3_kernel<<<dim3(ceil_div(D * 32, 256), B, 3), 256, 0, st>>>(part_dM, part_dc, part_db, chunks, B, D, d_M, d_centers,
                                                                         d_bp);
  return check_launch("sum_partials3_kernel");
}

extern "C" int sqlx_sql_bwd_summary(const float* x, const float* queries, const float* summary, const float* row_max,
                                    const float* row_sum, const float* d_summary, int B, int E, int Q, int n,
                                    int accumulate, float* d_x, float* d_queries, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  SQLX_REQUIRE(x && queries && summary && row_max && row_sum && d_summary && d_x && d_queries && workspace,
               "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, Q, 0, n), "shape E=%d Q=%d n=%d is not supported by the tensor-core path", E, Q, n);
  SQLX_REQUIRE(workspace_bytes >= sqlx_sql_mix_workspace_bytes(B, Q, 0, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int chunks = 0, tpc = 0;
  ws_plan(B, n, &chunks, &tpc);
  float* part_dK = reinterpret_cast<float*>(workspace);
  if (int e = ws_bwd_sum(x, queries, summary, row_max, row_sum, d_summary, B, Q, n, accumulate, d_x, part_dK, chunks, tpc, st))
    return e;
  sum_partials_kernel<<<dim3(ceil_div(Q * 32, 256), B), 256, 0, st>>>(part_dK, chunks, Q * 32, d_queries);
  return check_launch("sum_partials_kernel");
}

/* round-1 generation of sqlx_sql_bwd_summary (A/B and cross-check, see sqlx_sql_bwd_pred_mix_v1) */
extern "C" int sqlx_sql_bwd_summary_v1(const float* x, const float* queries, const float* summary, const float* row_max,
                                    const float* row_sum, const float* d_summary, int B, int E, int Q, int n,
                                    int accumulate, float* d_x, float* d_queries, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  SQLX_REQUIRE(x && queries && summary && row_max && row_sum && d_summary && d_x && d_queries && workspace,
               "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, Q, 0, n), "shape E=%d Q=%d n=%d is not supported by the tensor-core path", E, Q, n);
  SQLX_REQUIRE(workspace_bytes >= sqlx_sql_mix_workspace_bytes(B, Q, 0, n), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int chunks = 0, tpc = 0;
  tc_bwd_plan(B, n, &chunks, &tpc);
  float* part_dK = reinterpret_cast<float*>(workspace);
  if (int e = tc_bwd_sum(x, queries, summary, row_max, row_sum, d_summary, B, Q, n, accumulate, d_x, part_dK, chunks, tpc, st))
    return e;
  sum_partials_kernel<<<dim3(ceil_div(Q * 32, 256), B), 256, 0, st>>>(part_dK, chunks, Q * 32, d_queries);
  return check_launch("sum_partials_kernel");
}
