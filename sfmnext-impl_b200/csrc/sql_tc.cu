// SQL block on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// Contractions (per 128-pixel tile, E = 32), all in the mixed-weight form logits = (Wp K) x + b = M x + b:
//   Y[p,q]  = sum_e x[e,p] K[q,e]         A = x tile (MN-major: pixels contiguous, as NCHW holds it), B = K (K-major)
//   Z[p,d]  = sum_e x[e,p] M[d,e]         same operand layouts with M = Wp K in place of K
// Precision: the 1e-4 depth bar rules out single-pass TF32 (2e-3), so the forward contractions
// run as 3xTF32 (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulation in TMEM), which reproduces fp32.
// (The un-mixed formulation of round 1 -- Z = Y Wp^T with Y handed over in TMEM -- was superseded and removed; the
// exact-fp32 CUDA-core kernels of sql_fp32.cu are the in-repo cross-check of these kernels.)
//
// Reference lines: networks/layers.py:17-20, networks/depth_decoder_QTR.py:61,70.
#include "common.cuh"
#include "tc_common.cuh"

#include <math.h>
#include <stdlib.h>

namespace sqlx {

// ------------------------------------------------------------------------------------------------
// host: tensor map through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map_2d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols, int atom32) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) {
      cudaGetLastError();
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return SQLX_ECUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a [%llu x %llu] fp32 tensor", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
    return SQLX_ECUDA;
  }
  return SQLX_OK;
}

namespace tcsql {

using namespace tc;

constexpr int kE = 32;           // embedding channels (K of the first contraction)
constexpr int kTile = 128;       // pixels per tile = UMMA M
constexpr int kThreads = 128;    // 4 warps: warp w owns TMEM lanes [32w, 32w+32)
constexpr int kXBlock = 32 * kE * 4;         // one [32 e][32 px] SWIZZLE_128B block: 4096 B
constexpr int kXTile = 4 * kXBlock;          // 128 pixels: 16 KB
constexpr float kLog2eS = 1.4426950408889634f;   // softmax arithmetic runs in base 2: exp(a - m) = 2^(a log2e - m log2e)

// shared-memory carve-up (all operand regions 1024-B aligned for SWIZZLE_128B)
struct Smem {
  uint8_t* x_raw;  // TMA landing zone for the NEXT tile (prefetched while the current tile is computed)
  uint8_t* x_hi;   // [4 blocks][32 e][32 px]   TMA destination (SWIZZLE_128B_ATOM_32B), truncated in place
  uint8_t* x_lo;   // same layout, x - hi
  uint8_t* k_hi;   // [Qp rows][32 e]  K-major SW128
  uint8_t* k_lo;
  uint8_t* w_hi;   // [Qp/32 k-atoms][Dp rows][32 q]  K-major SW128 (also MN-major view for the transposed product)
  uint8_t* w_lo;
  float* bias;     // [Dp]
  float* cen;      // [Dp]
  uint64_t* bar_tma;
  uint64_t* bar_mma;
  uint32_t* tmem_slot;
};

__host__ __device__ inline size_t smem_bytes(int Qp, int Dp) {
  return 1024 /*alignment slack*/ + 3 * kXTile + 2 * (size_t)Qp * 128 + 2 * (size_t)(Qp / 32 + (Qp % 32 ? 1 : 0)) * Dp * 128 +
         2 * (size_t)(Dp > 0 ? Dp : 1) * sizeof(float) + 64;
}

__device__ __forceinline__ Smem carve(uint8_t* raw, int Qp, int Dp) {
  Smem s;
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  s.x_raw = p; p += kXTile;
  s.x_hi = p; p += kXTile;
  s.x_lo = p; p += kXTile;
  s.k_hi = p; p += (size_t)Qp * 128;
  s.k_lo = p; p += (size_t)Qp * 128;
  const int katoms = (Qp + 31) / 32;
  s.w_hi = p; p += (size_t)katoms * Dp * 128;
  s.w_lo = p; p += (size_t)katoms * Dp * 128;
  s.bias = reinterpret_cast<float*>(p); p += (size_t)(Dp > 0 ? Dp : 1) * sizeof(float);
  s.cen = reinterpret_cast<float*>(p); p += (size_t)(Dp > 0 ? Dp : 1) * sizeof(float);
  p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 7) & ~(uintptr_t)7);
  s.bar_tma = reinterpret_cast<uint64_t*>(p); p += 8;
  s.bar_mma = reinterpret_cast<uint64_t*>(p); p += 8;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

// queries [Q][32] -> K-major SW128 hi / lo blocks with Qp rows (rows >= Q zero)
__device__ __forceinline__ void stage_queries(const Smem& s, const float* __restrict__ qb, int Q, int Qp) {
  // eight independent loads in flight per trip (an in-order warp otherwise pays the global latency once per element)
  for (int base = threadIdx.x; base < Qp * kE; base += 8 * kThreads) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + j * kThreads;
      v[j] = (idx < Qp * kE && (idx >> 5) < Q) ? __ldg(qb + idx) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + j * kThreads;
      if (idx >= Qp * kE) break;
      const float hi = tf32_hi(v[j]);
      const uint32_t off = sw128_offset(idx >> 5, idx & 31);
      *reinterpret_cast<float*>(s.k_hi + off) = hi;
      *reinterpret_cast<float*>(s.k_lo + off) = v[j] - hi;
    }
  }
}

// split the freshly landed x tile (x_raw) into hi and lo operand tiles (same swizzled layout: elementwise)
__device__ __forceinline__ void split_x_tile(const Smem& s) {
  const float4* raw = reinterpret_cast<const float4*>(s.x_raw);
  float4* hi = reinterpret_cast<float4*>(s.x_hi);
  float4* lo = reinterpret_cast<float4*>(s.x_lo);
#pragma unroll
  for (int i = 0; i < kXTile / 16 / kThreads; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    const float4 v = raw[idx];
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    hi[idx] = h;
    lo[idx] = l;
  }
}

// Y[128 px, Qp] = x^T K as 3xTF32 into TMEM columns [tm_y, tm_y + Qp).  One thread issues.
__device__ __forceinline__ void issue_y(const Smem& s, uint32_t tm_y, int Qp) {
  const uint32_t idesc = make_idesc_tf32(kTile, Qp, /*A MN-major*/ 1, /*B K-major*/ 0);
  const uint32_t xh = smem_u32(s.x_hi), xl = smem_u32(s.x_lo), kh = smem_u32(s.k_hi), kl = smem_u32(s.k_lo);
  uint32_t acc = 0;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t xa = pass == 1 ? xl : xh;
    const uint32_t kb = pass == 2 ? kl : kh;
#pragma unroll
    for (int k = 0; k < kE / 8; ++k) {
      // A: MN-major (SWIZZLE_128B_BASE32B), 8 e-rows = 1024 B per k-step; 32-pixel blocks are 4096 B apart (LBO)
      const uint64_t da = make_desc_mn32(xa + k * 1024, kXBlock);
      // B: K-major, 8-row groups 1024 B apart (SBO); 8 e (32 B) per k-step inside the 128-B row
      const uint64_t db = make_desc_sw128(kb + k * 32, 16, 1024);
      umma_tf32_ss(tm_y, da, db, idesc, acc);
      acc = 1;
    }
  }
}

// one 128-pixel x tile = four [32 e][32 px] boxes (out-of-range pixels are zero-filled by TMA)
__device__ __forceinline__ void issue_x_tma(const Smem& s, const CUtensorMap* xmap, int p0, int row0) {
  mbar_arrive_expect_tx(s.bar_tma, kXTile);
#pragma unroll
  for (int j = 0; j < 4; ++j) tma_load_2d(s.x_raw + j * kXBlock, xmap, p0 + 32 * j, row0, s.bar_tma);
}

// softmax over the Dp logits of this thread's pixel (TMEM lane) and expected bin centre, one pass (online max)
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// s.bias holds bias * log2e: the logits are handled in base 2 (one FFMA + one MUFU.EX2 per element)
__device__ __forceinline__ float softmax_expect(const Smem& s, uint32_t lane_base, uint32_t tm_z, int Dp) {
  // four independent accumulation chains (the kernels run 4-12 warps per SM: a single dependent add / fma chain per
  // thread leaves the issue slots idle), bias and centres fetched as 128-bit shared loads
  float m = -INFINITY, se[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* bias4 = reinterpret_cast<const float4*>(s.bias);
  const float4* cen4 = reinterpret_cast<const float4*>(s.cen);
  for (int c = 0; c < Dp; c += 16) {
    float v[16];
    tmem_ld16(lane_base + tm_z + c, v);
    tmem_wait_ld();
    float cmx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bq = bias4[(c >> 2) + j];
      v[4 * j] = fmaf(v[4 * j], kLog2eS, bq.x); v[4 * j + 1] = fmaf(v[4 * j + 1], kLog2eS, bq.y);
      v[4 * j + 2] = fmaf(v[4 * j + 2], kLog2eS, bq.z); v[4 * j + 3] = fmaf(v[4 * j + 3], kLog2eS, bq.w);
      cmx[j] = fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3]));
    }
    const float cm = fmaxf(fmaxf(cmx[0], cmx[1]), fmaxf(cmx[2], cmx[3]));
    if (cm > m) {
      const float r = ex2_approx(m - cm);
#pragma unroll
      for (int j = 0; j < 4; ++j) { se[j] *= r; sc[j] *= r; }
      m = cm;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 cq = cen4[(c >> 2) + j];
      const float e0 = ex2_approx(v[4 * j] - m), e1 = ex2_approx(v[4 * j + 1] - m);
      const float e2 = ex2_approx(v[4 * j + 2] - m), e3 = ex2_approx(v[4 * j + 3] - m);
      se[j] += (e0 + e1) + (e2 + e3);
      sc[j] += fmaf(e0, cq.x, e1 * cq.y) + fmaf(e2, cq.z, e3 * cq.w);
    }
  }
  return ((sc[0] + sc[1]) + (sc[2] + sc[3])) / ((se[0] + se[1]) + (se[2] + se[3]));
}

// ------------------------------------------------------------------------------------------------
// pixel-softmax summaries, flash style with queries on the TMEM lanes:
//   Y^T[q, px] = K x          A = K (K-major, M = 128 padded queries), B = x tile (MN-major, N = 64 pixels)
//   S[q, e]    = P x^T        A = P = exp(Y^T - m_q) written back to TMEM (hi / lo), B = x tile (K-major over pixels)
// thread = query keeps the running (max, sum, acc[32]) in registers; per-CTA partials -> split-softmax combine.
// TMEM columns: [0,64) Y^T / P_hi   [64,128) P_lo   [128,160) S of the current tile
// ------------------------------------------------------------------------------------------------
constexpr int kTileS = 64;                    // pixels per summary tile
constexpr int kXTileS = 2 * kXBlock;          // 8 KB per layout

struct SmemS {
  uint8_t* raw_mn;   // TMA landing, SWIZZLE_128B_ATOM_32B (consumed MN-major by the first contraction)
  uint8_t* raw_k;    // TMA landing, SWIZZLE_128B (consumed K-major over pixels by the second contraction)
  uint8_t* mn_hi; uint8_t* mn_lo;
  uint8_t* k_hi_x; uint8_t* k_lo_x;
  uint8_t* q_hi; uint8_t* q_lo;   // queries [128 rows][32 e] K-major SW128
  uint64_t* bar_tma; uint64_t* bar_mma;
  uint32_t* tmem_slot;
};
constexpr size_t kSmemSBytes = 1024 + 6 * kXTileS + 2 * 128 * 128 + 64;

__device__ __forceinline__ SmemS carve_s(uint8_t* raw) {
  SmemS s;
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  s.raw_mn = p; p += kXTileS;
  s.raw_k = p; p += kXTileS;
  s.mn_hi = p; p += kXTileS;
  s.mn_lo = p; p += kXTileS;
  s.k_hi_x = p; p += kXTileS;
  s.k_lo_x = p; p += kXTileS;
  s.q_hi = p; p += 128 * 128;
  s.q_lo = p; p += 128 * 128;
  s.bar_tma = reinterpret_cast<uint64_t*>(p); p += 8;
  s.bar_mma = reinterpret_cast<uint64_t*>(p); p += 8;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

__device__ __forceinline__ void issue_x_tma_s(const SmemS& s, const CUtensorMap* map_mn, const CUtensorMap* map_k, int p0,
                                              int row0) {
  mbar_arrive_expect_tx(s.bar_tma, 2 * kXTileS);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    tma_load_2d(s.raw_mn + j * kXBlock, map_mn, p0 + 32 * j, row0, s.bar_tma);
    tma_load_2d(s.raw_k + j * kXBlock, map_k, p0 + 32 * j, row0, s.bar_tma);
  }
}

__device__ __forceinline__ void split_tile(const uint8_t* raw, uint8_t* hi_, uint8_t* lo_, int bytes) {
  const float4* r = reinterpret_cast<const float4*>(raw);
  float4* hi = reinterpret_cast<float4*>(hi_);
  float4* lo = reinterpret_cast<float4*>(lo_);
  for (int idx = threadIdx.x; idx < bytes / 16; idx += kThreads) {
    const float4 v = r[idx];
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    hi[idx] = h;
    lo[idx] = l;
  }
}

__global__ void __launch_bounds__(kThreads) sql_tc_summary_kernel(const __grid_constant__ CUtensorMap map_mn,
                                                                  const __grid_constant__ CUtensorMap map_k,
                                                                  const float* __restrict__ queries, int Q, int n,
                                                                  int tiles_per_chunk,
                                                                  float* __restrict__ partial /*[B][chunks][Q][34]*/) {
  extern __shared__ uint8_t smem_raw[];
  const SmemS s = carve_s(smem_raw);
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int warp = threadIdx.x >> 5;
  constexpr uint32_t kCols = 256;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_mn);
    tma_prefetch_desc(&map_k);
    mbar_init(s.bar_tma, 1);
    mbar_init(s.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(s.tmem_slot, kCols);
    tmem_relinquish();
  }
  const int ntiles = (n + kTileS - 1) / kTileS;
  const int t_begin = chunk * tiles_per_chunk;
  const int t_end = min(ntiles, t_begin + tiles_per_chunk);
  __syncthreads();
  if (threadIdx.x == 0 && t_begin < t_end) issue_x_tma_s(s, &map_mn, &map_k, t_begin * kTileS, b * kE);
  // queries -> [128 rows][32] K-major SW128 hi / lo, rows >= Q zero
  {
    const float* qb = queries + (size_t)b * Q * kE;
    for (int base = threadIdx.x; base < 128 * kE; base += 8 * kThreads) {   // eight loads in flight per trip
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = base + j * kThreads;
        v[j] = (idx >> 5) < Q ? __ldg(qb + idx) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = base + j * kThreads;
        const float hi = tf32_hi(v[j]);
        const uint32_t off = sw128_offset(idx >> 5, idx & 31);
        *reinterpret_cast<float*>(s.q_hi + off) = hi;
        *reinterpret_cast<float*>(s.q_lo + off) = v[j] - hi;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc1 = make_idesc_tf32(128, kTileS, 0, 1);   // A = K (K-major), B = x (MN-major)
  const uint32_t idesc2 = make_idesc_tf32(128, kE, 0, 0);       // A = P (TMEM), B = x (K-major over pixels)
  float m = -INFINITY, l = 0.f, acc[kE];
#pragma unroll
  for (int e = 0; e < kE; ++e) acc[e] = 0.f;
  uint32_t ph_tma = 0, ph_mma = 0;
  for (int t = t_begin; t < t_end; ++t) {
    const int p0 = t * kTileS;
    mbar_wait(s.bar_tma, ph_tma); ph_tma ^= 1;
    split_tile(s.raw_mn, s.mn_hi, s.mn_lo, kXTileS);
    split_tile(s.raw_k, s.k_hi_x, s.k_lo_x, kXTileS);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (t + 1 < t_end) issue_x_tma_s(s, &map_mn, &map_k, p0 + kTileS, b * kE);
      tc_fence_after();
      const uint32_t qh = smem_u32(s.q_hi), ql = smem_u32(s.q_lo), xh = smem_u32(s.mn_hi), xl = smem_u32(s.mn_lo);
      uint32_t a = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t qa = pass == 1 ? ql : qh;
        const uint32_t xb = pass == 2 ? xl : xh;
#pragma unroll
        for (int k = 0; k < kE / 8; ++k) {
          umma_tf32_ss(tmem, make_desc_sw128(qa + k * 32, 16, 1024), make_desc_mn32(xb + k * 1024, kXBlock), idesc1, a);
          a = 1;
        }
      }
      umma_commit(s.bar_mma);
    }
    mbar_wait(s.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    // ---- this thread's query row: tile max, lazy rescale, P = exp(y - m) -> TMEM (hi in place, lo beside it)
    const int valid = min(kTileS, n - p0);   // pixels >= n were zero-filled by TMA: exclude them
    const bool full = valid == kTileS;        // every tile but (possibly) the last: no per-element range checks
    float tmax = -INFINITY;
    for (int c = 0; c < kTileS; c += 16) {
      float v[16];
      tmem_ld16(lane_base + c, v);
      tmem_wait_ld();
      if (full) {
        float t4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t4[j] = fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3]));
        tmax = fmaxf(tmax, fmaxf(fmaxf(t4[0], t4[1]), fmaxf(t4[2], t4[3])));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c + i < valid) tmax = fmaxf(tmax, v[i]);
      }
    }
    float rs = 1.f;
    if (tmax > m + 8.f) {          // lazy: the reference point only moves when exceeded by > 8 (exp args <= 8)
      rs = __expf(m - tmax);       // m = -inf at first: 0
      m = tmax;
    }
    const float nm2 = -m * kLog2eS;   // exp(y - m) = 2^(y log2e - m log2e): one FFMA + one MUFU.EX2 per element
    float ls4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < kTileS; c += 16) {
      float v[16], lo[16];
      tmem_ld16(lane_base + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float pe;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(fmaf(v[i], kLog2eS, nm2)));
        if (!full && c + i >= valid) pe = 0.f;
        ls4[i & 3] += pe;
        const float h = tf32_hi(pe);
        v[i] = h;
        lo[i] = pe - h;
      }
      tmem_st16(lane_base + c, v);
      tmem_st16(lane_base + kTileS + c, lo);
    }
    const float lsum = (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
    tmem_wait_st();
    l = l * rs + lsum;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t xh = smem_u32(s.k_hi_x), xl = smem_u32(s.k_lo_x);
      uint32_t a = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t pa = tmem + (pass == 1 ? kTileS : 0);
        const uint32_t xb = pass == 2 ? xl : xh;
#pragma unroll
        for (int k = 0; k < kTileS / 8; ++k) {
          // B: [32 e rows][32 px] K-major blocks (one per 32 pixels); 8 px = 32 B per k-step
          umma_tf32_ts(tmem + 2 * kTileS, pa + k * 8,
                       make_desc_sw128(xb + (uint32_t)(k >> 2) * kXBlock + (uint32_t)(k & 3) * 32u, 16, 1024), idesc2, a);
          a = 1;
        }
      }
      umma_commit(s.bar_mma);
    }
    mbar_wait(s.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    {
      float v[16];
      tmem_ld16(lane_base + 2 * kTileS, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], rs, v[i]);
      tmem_ld16(lane_base + 2 * kTileS + 16, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[16 + i] = fmaf(acc[16 + i], rs, v[i]);
    }
    tc_fence_before();
    __syncthreads();
  }
  const int q = threadIdx.x;
  if (q < Q) {
    float* out = partial + (((size_t)b * chunks + chunk) * Q + q) * (kE + 2);
    out[0] = m; out[1] = l;
#pragma unroll
    for (int e = 0; e < kE; ++e) out[2 + e] = acc[e];
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kCols);
}

// 2^x as one MUFU.EX2 (callers fold log2(e) and the softmax reference point into x with one FFMA)
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
constexpr float kLog2e = 1.4426950408889634f;

// Per-thread part of sw128_offset(row, lane) for the eight values of row & 7: with a compile-time row the transposed
// store address becomes base + row * 128 + sw[row & 7] (no per-element address arithmetic).
__device__ __forceinline__ void sw128_lane_offsets(int lane, uint32_t (&sw)[8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) sw[r] = ((((uint32_t)lane >> 2) ^ (uint32_t)r) << 4) + (((uint32_t)lane & 3u) << 2);
}

// ================================================================================================
// "Mixed-weight" decomposition (all Q, D <= 128):   logits = Wp (K x) + b = (Wp K) x + b = M x + b,  M [D x 32].
// The depth regression and its backward then contract over E = 32 instead of Q, need no Wp tiles on chip and no
// y -> logits hand-off; the summary path (which does need y) is independent of Wp.  M = Wp K, dWp = dM K^T and
// dK_1 = Wp^T dM are tiny per-sample products left to cuBLAS on the host side (sqlx/sql.py).
//   forward :  pred      = softmax_d(M x + b) . centers                                  sql_tc_pred2_kernel
//   backward:  pass 1    = dM, d_bp, d_centers, d_x (regression path)                    sql_tc_bwd_pred_kernel
//              pass 2    = d_x += summary path, d_K_2                                    sql_tc_bwd_sum_kernel
// ================================================================================================

// M tiles: [DP rows][32 e] K-major SW128, hi / lo
__device__ __forceinline__ void stage_mix(uint8_t* m_hi, uint8_t* m_lo, const float* __restrict__ Mb, int D, int DP) {
  for (int base = threadIdx.x; base < DP * kE; base += 8 * kThreads) {   // eight loads in flight per trip
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + j * kThreads;
      v[j] = (idx < DP * kE && (idx >> 5) < D) ? __ldg(Mb + idx) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + j * kThreads;
      if (idx >= DP * kE) break;
      const float hi = tf32_hi(v[j]);
      const uint32_t off = sw128_offset(idx >> 5, idx & 31);
      *reinterpret_cast<float*>(m_hi + off) = hi;
      if (m_lo) *reinterpret_cast<float*>(m_lo + off) = v[j] - hi;
    }
  }
}

// Z[128 px, DP] = x^T M^T as 3xTF32.  x tiles MN-major (hi / lo), M tiles K-major (hi / lo).
__device__ __forceinline__ void issue_xm(uint32_t xh, uint32_t xl, uint32_t mh, uint32_t ml, uint32_t tm_z, int DP) {
  const uint32_t idesc = make_idesc_tf32(kTile, DP, 1, 0);
  uint32_t acc = 0;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t xa = pass == 1 ? xl : xh;
    const uint32_t mb = pass == 2 ? ml : mh;
#pragma unroll
    for (int k = 0; k < kE / 8; ++k) {
      umma_tf32_ss(tm_z, make_desc_mn32(xa + k * 1024, kXBlock), make_desc_sw128(mb + k * 32, 16, 1024), idesc, acc);
      acc = 1;
    }
  }
}

struct SmemP2 {
  uint8_t *x_raw, *x_hi, *x_lo, *m_hi, *m_lo;
  float *bias, *cen;
  uint64_t *bar_tma, *bar_mma;
  uint32_t* tmem_slot;
};
__host__ __device__ inline size_t smem_p2_bytes(int DP) { return 1024 + 3 * kXTile + 2 * (size_t)DP * 128 + 2 * DP * 4 + 64; }

__global__ void __launch_bounds__(kThreads) sql_tc_pred2_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                const float* __restrict__ Mx /*[B,D,32]*/,
                                                                const float* __restrict__ bp,
                                                                const float* __restrict__ centers, int D, int DP, int n,
                                                                int tiles_per_chunk, uint32_t tmem_cols,
                                                                float* __restrict__ pred) {
  extern __shared__ uint8_t smem_raw[];
  Smem s;   // reuse the field names of the common struct for the helpers (split_x_tile, issue_x_tma, softmax_expect)
  {
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    s.x_raw = p; p += kXTile;
    s.x_hi = p; p += kXTile;
    s.x_lo = p; p += kXTile;
    s.k_hi = p; p += (size_t)DP * 128;    // M hi
    s.k_lo = p; p += (size_t)DP * 128;    // M lo
    s.w_hi = s.w_lo = nullptr;
    s.bias = reinterpret_cast<float*>(p); p += DP * 4;
    s.cen = reinterpret_cast<float*>(p); p += DP * 4;
    s.bar_tma = reinterpret_cast<uint64_t*>(p); p += 8;
    s.bar_mma = reinterpret_cast<uint64_t*>(p); p += 8;
    s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&xmap);
    mbar_init(s.bar_tma, 1);
    mbar_init(s.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(s.tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  __syncthreads();
  if (threadIdx.x == 0 && t_begin < t_end) issue_x_tma(s, &xmap, t_begin * kTile, b * kE);
  stage_mix(s.k_hi, s.k_lo, Mx + (size_t)b * D * kE, D, DP);
  for (int d = threadIdx.x; d < DP; d += kThreads) {
    s.bias[d] = d < D ? __ldg(bp + d) * kLog2eS : -INFINITY;
    s.cen[d] = d < D ? __ldg(centers + (size_t)b * D + d) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  for (int t = t_begin; t < t_end; ++t) {
    const int p0 = t * kTile;
    mbar_wait(s.bar_tma, phase);
    split_x_tile(s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (t + 1 < t_end) issue_x_tma(s, &xmap, p0 + kTile, b * kE);
      tc_fence_after();
      issue_xm(smem_u32(s.x_hi), smem_u32(s.x_lo), smem_u32(s.k_hi), smem_u32(s.k_lo), tmem, DP);
      umma_commit(s.bar_mma);
    }
    mbar_wait(s.bar_mma, phase);
    tc_fence_after();
    const float pr = softmax_expect(s, lane_base, 0, DP);
    const int p = p0 + warp * 32 + lane;
    if (p < n) pred[(size_t)b * n + p] = pr;
    tc_fence_before();
    __syncthreads();
    phase ^= 1;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// backward pass 1 (regression path).  TMEM: [0,DP) logits / dz   [DP,DP+32) d_x tile   [DP+32,DP+80) accumulator
//   accumulator rows d: cols 0..31 = dM[d,e] = sum_p dz[p,d] x[e,p], col 32 = d_bp[d]
//   d_centers[d] = sum_p pi g is accumulated per thread in registers and reduced across the CTA at the end
// ------------------------------------------------------------------------------------------------
constexpr int kAccN = kE + 16;   // x columns + ones column (+ padding to a multiple of 16)
template <int DP>
struct BwdPredSmem {
  static constexpr size_t x_raw = 0, x_hi = kXTile, x_lo = 2 * kXTile, m_hi = 3 * kXTile, m_lo = m_hi + DP * 128,
                          mT = m_lo + DP * 128,                      // [DP/32 atoms][32 e rows][32 d]
                          bx = mT + (DP / 32) * 32 * 128,           // [4 px atoms][48 rows][32 px]: x (TMA) | ones
                          adz = bx + 4 * kAccN * 128,               // [4 px atoms][128 rows][32 px]: dz^T
                          tail = adz + 4 * 128 * 128 + 1024;        // (+1 KB: the final [128][DP+1] reduction scratch)
  static constexpr size_t bytes = 1024 + tail + 2 * DP * 4 + 64;
};

// Software-pipelined tile loop: the d_x / accumulator MMAs of tile t are committed to a second mbarrier and NOT waited
// for; tile t+1's TMA wait, hi/lo split and logits MMA are issued first, and the d_x rows of tile t leave TMEM while that
// logits MMA runs (measured on a B200 against the serial loop: bit-identical results, 8 % faster, r02 candidates.log).
template <int DP>
__global__ void __launch_bounds__(kThreads) sql_tc_bwd_pred_kernel(
    const __grid_constant__ CUtensorMap map_mn, const __grid_constant__ CUtensorMap map_k, const float* __restrict__ Mx,
    const float* __restrict__ bp, const float* __restrict__ centers, const float* __restrict__ g_pred, int D, int n,
    int tiles_per_chunk, float* __restrict__ d_x, float* __restrict__ part_dM /*[cta][D][32]*/,
    float* __restrict__ part_db /*[cta][D]*/, float* __restrict__ part_dc /*[cta][D]*/) {
  using L = BwdPredSmem<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Smem s;
  s.x_raw = base + L::x_raw; s.x_hi = base + L::x_hi; s.x_lo = base + L::x_lo;
  s.k_hi = base + L::m_hi; s.k_lo = base + L::m_lo;
  uint8_t* mT = base + L::mT;
  uint8_t* bx = base + L::bx;
  uint8_t* adz = base + L::adz;
  s.bias = reinterpret_cast<float*>(base + L::tail);
  s.cen = s.bias + DP;
  s.bar_tma = reinterpret_cast<uint64_t*>(s.cen + DP);
  s.bar_mma = s.bar_tma + 1;
  uint64_t* bar_bx = s.bar_tma + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.bar_tma + 3);
  uint64_t* bar_mma2 = s.bar_tma + 4;     // completion of a tile's d_x / accumulator MMAs
  const int b = blockIdx.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 256;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_mn);
    tma_prefetch_desc(&map_k);
    mbar_init(s.bar_tma, 1);
    mbar_init(s.bar_mma, 1);
    mbar_init(bar_bx, 1);
    mbar_init(bar_mma2, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(s.tmem_slot, kCols);
    tmem_relinquish();
  }
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  __syncthreads();
  if (threadIdx.x == 0 && t_begin < t_end) issue_x_tma(s, &map_mn, t_begin * kTile, b * kE);
  const float* Mb = Mx + (size_t)b * D * kE;
  stage_mix(s.k_hi, s.k_lo, Mb, D, DP);
  for (int base = threadIdx.x; base < DP * kE; base += 8 * kThreads) {   // mT[e][d] = M[d][e]; eight loads per trip
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + j * kThreads;
      v[j] = (idx < DP * kE && (idx >> 5) < D) ? __ldg(Mb + idx) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = base + j * kThreads;
      if (idx >= DP * kE) break;
      const int d = idx >> 5, e = idx & 31;
      *reinterpret_cast<float*>(mT + (uint32_t)(d >> 5) * 32u * 128u + sw128_offset(e, d & 31)) = v[j];
    }
  }
  for (int i = threadIdx.x; i < 4 * 16 * 32; i += kThreads) {      // rows 32..47 of every pixel atom: ones | zeros
    const int atom = i / (16 * 32), rem = i - atom * 16 * 32, row = kE + (rem >> 5), col = rem & 31;
    *reinterpret_cast<float*>(bx + atom * kAccN * 128 + sw128_offset(row, col)) = row == kE ? 1.f : 0.f;
  }
  for (int i = threadIdx.x; i < 4 * 128 * 32 / 4; i += kThreads) reinterpret_cast<float4*>(adz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int d = threadIdx.x; d < DP; d += kThreads) {
    s.bias[d] = d < D ? __ldg(bp + d) * kLog2e : -INFINITY;   // logits are handled in base 2: t = z log2e + bias log2e
    s.cen[d] = d < D ? __ldg(centers + (size_t)b * D + d) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t swl[8];
  sw128_lane_offsets(lane, swl);
  constexpr uint32_t tm_z = 0, tm_dx = DP, tm_acc = DP + 32;
  const uint32_t id_acc = make_idesc_tf32(128, kAccN, 0, 0);
  const uint32_t id_dx = make_idesc_tf32(128, 32, 0, 0);
  float dc[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) dc[d] = 0.f;
  uint32_t ph_tma = 0, ph_mma = 0, ph_bx = 0, acc_on = 0;
  uint32_t ph_mma2 = 0;
  bool pending = false;     // the previous tile's d_x rows still sit in TMEM (its MMAs possibly in flight)
  int p_prev = 0;
  float* dxb = d_x + (size_t)b * kE * n;
  auto store_dx = [&](int pp) {
#pragma unroll
    for (int c = 0; c < kE; c += 16) {
      float v[16];
      tmem_ld16(lane_base + tm_dx + c, v);
      tmem_wait_ld();
      if (pp < n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dxb[(size_t)(c + i) * n + pp] = v[i];
      }
    }
  };
  for (int t = t_begin; t < t_end; ++t) {
    const int p0 = t * kTile;
    const int p = p0 + warp * 32 + lane;
    const float gp = p < n ? __ldg(g_pred + (size_t)b * n + p) : 0.f;   // upstream gradient: in flight during the tile
    mbar_wait(s.bar_tma, ph_tma); ph_tma ^= 1;
    split_x_tile(s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      // the previous tile's MMAs read bx and the dz columns of TMEM: they must be done before both are reused
      if (pending) mbar_wait(bar_mma2, ph_mma2);
      if (t + 1 < t_end) issue_x_tma(s, &map_mn, p0 + kTile, b * kE);
      mbar_arrive_expect_tx(bar_bx, kXTile);   // the K-major x rows of the accumulator's B tile (free: previous MMAs waited)
#pragma unroll
      for (int j = 0; j < 4; ++j) tma_load_2d(bx + j * kAccN * 128, &map_k, p0 + 32 * j, b * kE, bar_bx);
      tc_fence_after();
      issue_xm(smem_u32(s.x_hi), smem_u32(s.x_lo), smem_u32(s.k_hi), smem_u32(s.k_lo), tmem + tm_z, DP);
      umma_commit(s.bar_mma);
    }
    if (pending) {     // the previous tile's d_x rows leave TMEM while this tile's logits MMA runs
      mbar_wait(bar_mma2, ph_mma2); ph_mma2 ^= 1;
      tc_fence_after();
      store_dx(p_prev);
      pending = false;       // (the d_x columns are rewritten only after this tile's epilogue and its block barrier)
    }
    mbar_wait(s.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    // pass A: online max / sum / expectation over the base-2 logits t = (z + bias) log2e
    float m = -INFINITY, se = 0.f, sc = 0.f;
    const float4* bias4 = reinterpret_cast<const float4*>(s.bias);
    const float4* cen4 = reinterpret_cast<const float4*>(s.cen);
#pragma unroll
    for (int c = 0; c < DP; c += 16) {
      float v[16];
      tmem_ld16(lane_base + tm_z + c, v);
      tmem_wait_ld();
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bq = bias4[(c >> 2) + j];
        v[4 * j] = fmaf(v[4 * j], kLog2e, bq.x); v[4 * j + 1] = fmaf(v[4 * j + 1], kLog2e, bq.y);
        v[4 * j + 2] = fmaf(v[4 * j + 2], kLog2e, bq.z); v[4 * j + 3] = fmaf(v[4 * j + 3], kLog2e, bq.w);
        cm = fmaxf(cm, fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3])));
      }
      if (cm > m) { const float r = ex2_fast(m - cm); se *= r; sc *= r; m = cm; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 cq = cen4[(c >> 2) + j];
        const float e0 = ex2_fast(v[4 * j] - m), e1 = ex2_fast(v[4 * j + 1] - m);
        const float e2 = ex2_fast(v[4 * j + 2] - m), e3 = ex2_fast(v[4 * j + 3] - m);
        se += (e0 + e1) + (e2 + e3);
        sc += fmaf(e0, cq.x, e1 * cq.y) + fmaf(e2, cq.z, e3 * cq.w);
      }
    }
    const float inv = 1.f / se, pr = sc * inv;
    const float g = gp * inv;   // g / sum folded together
    // pass B: pi g, dz -> registers (d_centers), TMEM (A operand of d_x) and transposed shared memory (A operand of dM)
    uint8_t* adw = adz + warp * 128 * 128;
#pragma unroll
    for (int c = 0; c < DP; c += 16) {
      float v[16];
      tmem_ld16(lane_base + tm_z + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bq = bias4[(c >> 2) + j], cq = cen4[(c >> 2) + j];
        const float bv[4] = {bq.x, bq.y, bq.z, bq.w}, cv[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 4 * j + k, d = c + i;
          const float pg = ex2_fast(fmaf(v[i], kLog2e, bv[k]) - m) * g;
          dc[d] += pg;
          v[i] = pg * (cv[k] - pr);
          *reinterpret_cast<float*>(adw + d * 128 + swl[d & 7]) = v[i];
        }
      }
      tmem_st16(lane_base + tm_z + c, v);
    }
    tmem_wait_st();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      mbar_wait(bar_bx, ph_bx);
      const uint32_t a0 = smem_u32(adz), b0 = smem_u32(bx), mt = smem_u32(mT);
#pragma unroll
      for (int k = 0; k < kTile / 8; ++k) {     // accumulator += dz^T [x | 1]      (K = 128 pixels)
        umma_tf32_ss(tmem + tm_acc, make_desc_sw128(a0 + (k >> 2) * 128 * 128 + (k & 3) * 32, 16, 1024),
                     make_desc_sw128(b0 + (k >> 2) * kAccN * 128 + (k & 3) * 32, 16, 1024), id_acc, acc_on);
        acc_on = 1;
      }
#pragma unroll
      for (int k = 0; k < DP / 8; ++k)          // d_x tile = dz M               (K = DP bins, A from TMEM)
        umma_tf32_ts(tmem + tm_dx, tmem + tm_z + k * 8,
                     make_desc_sw128(mt + (uint32_t)(k >> 2) * 32u * 128u + (uint32_t)(k & 3) * 32u, 16, 1024), id_dx, k > 0);
      umma_commit(bar_mma2);
    }
    ph_bx ^= 1;
    pending = true;
    p_prev = p;
  }
  if (pending) {       // drain: the last tile's d_x rows (the commit also covers every accumulator MMA)
    mbar_wait(bar_mma2, ph_mma2);
    tc_fence_after();
    store_dx(p_prev);
    tc_fence_before();
    __syncthreads();
  }
  // accumulator rows (lane = bin d): dM, d_bp
  const int d_row = threadIdx.x;
  const bool have = t_begin < t_end;
#pragma unroll
  for (int c = 0; c < kAccN; c += 16) {
    float v[16];
    if (have) {
      tmem_ld16(lane_base + tm_acc + c, v);
      tmem_wait_ld();
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    if (d_row < D) {
      if (c < kE) {
#pragma unroll
        for (int i = 0; i < 16; ++i) part_dM[((size_t)cta * D + d_row) * kE + c + i] = v[i];
      } else {
        part_db[(size_t)cta * D + d_row] = v[0];
      }
    }
  }
  // d_centers: reduce the per-thread (per pixel slot) accumulators over the 128 threads through shared memory
  __syncthreads();
  float* red = reinterpret_cast<float*>(adz);   // [128 threads][DP], padded row pitch DP+1
#pragma unroll
  for (int d = 0; d < DP; ++d) red[threadIdx.x * (DP + 1) + d] = dc[d];
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kThreads) {
    float acc = 0.f;
    for (int tt = 0; tt < kThreads; ++tt) acc += red[tt * (DP + 1) + d];
    part_dc[(size_t)cta * D + d] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kCols);
}

// ------------------------------------------------------------------------------------------------
// backward pass 2 (summary path):  a = softmax_pixels(y),  dy = a (t - delta),  t = x^T ds^T
//   d_x (+)= dy K + a ds ,   d_K_2 += dy^T x
// TMEM: [0,QP) y -> dy   [QP,2QP) t -> a   [2QP,2QP+32) d_x tile   [2QP+32,2QP+64) d_K accumulator
// ------------------------------------------------------------------------------------------------
template <int QP>
struct BwdSumSmem {
  static constexpr size_t x_raw = 0, x_hi = kXTile, x_lo = 2 * kXTile, x_k = 3 * kXTile, k_hi = 4 * kXTile,
                          k_lo = k_hi + QP * 128, ds = k_lo + QP * 128, kT = ds + QP * 128,
                          dsT = kT + (QP / 32) * 32 * 128, dyT = dsT + (QP / 32) * 32 * 128, tail = dyT + 4 * 128 * 128;
  static constexpr size_t bytes = 1024 + tail + 3 * QP * 4 + 64;
};

// Same software pipelining as sql_tc_bwd_pred_kernel: tile t's d_x rows leave TMEM while tile t+1's y / t MMAs run.
template <int QP>
__global__ void __launch_bounds__(kThreads) sql_tc_bwd_sum_kernel(
    const __grid_constant__ CUtensorMap map_mn, const __grid_constant__ CUtensorMap map_k,
    const float* __restrict__ queries, const float* __restrict__ summary, const float* __restrict__ row_max,
    const float* __restrict__ row_sum, const float* __restrict__ d_summary, int Q, int n, int tiles_per_chunk,
    int accumulate, float* __restrict__ d_x, float* __restrict__ part_dK /*[cta][Q][32]*/) {
  using L = BwdSumSmem<QP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Smem s;
  s.x_raw = base + L::x_raw; s.x_hi = base + L::x_hi; s.x_lo = base + L::x_lo;
  s.k_hi = base + L::k_hi; s.k_lo = base + L::k_lo;
  uint8_t* x_k = base + L::x_k;
  uint8_t* ds = base + L::ds;
  uint8_t* kT = base + L::kT;
  uint8_t* dsT = base + L::dsT;
  uint8_t* dyT = base + L::dyT;
  float* mq = reinterpret_cast<float*>(base + L::tail);
  float* il = mq + QP;
  float* dl = il + QP;
  s.bar_tma = reinterpret_cast<uint64_t*>(dl + QP);
  s.bar_mma = s.bar_tma + 1;
  uint64_t* bar_xk = s.bar_tma + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.bar_tma + 3);
  uint64_t* bar_mma2 = s.bar_tma + 4;     // completion of a tile's d_x / d_K MMAs
  const int b = blockIdx.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 512;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_mn);
    tma_prefetch_desc(&map_k);
    mbar_init(s.bar_tma, 1);
    mbar_init(s.bar_mma, 1);
    mbar_init(bar_xk, 1);
    mbar_init(bar_mma2, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(s.tmem_slot, kCols);
    tmem_relinquish();
  }
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  __syncthreads();
  if (threadIdx.x == 0 && t_begin < t_end) issue_x_tma(s, &map_mn, t_begin * kTile, b * kE);
  stage_queries(s, queries + (size_t)b * Q * kE, Q, QP);
  {
    const float* qb = queries + (size_t)b * Q * kE;
    const float* dsb = d_summary + (size_t)b * Q * kE;
    for (int base = threadIdx.x; base < QP * kE; base += 4 * kThreads) {   // eight loads in flight per trip
      float kv[4], dv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = base + j * kThreads;
        const bool on = idx < QP * kE && (idx >> 5) < Q;
        kv[j] = on ? __ldg(qb + idx) : 0.f;
        dv[j] = on ? __ldg(dsb + idx) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = base + j * kThreads;
        if (idx >= QP * kE) break;
        const int q = idx >> 5, e = idx & 31;
        *reinterpret_cast<float*>(ds + sw128_offset(q, e)) = dv[j];
        const uint32_t offT = (uint32_t)(q >> 5) * 32u * 128u + sw128_offset(e, q & 31);
        *reinterpret_cast<float*>(kT + offT) = kv[j];
        *reinterpret_cast<float*>(dsT + offT) = dv[j];
      }
    }
    for (int idx = threadIdx.x; idx < 4 * 128 * 32 / 4; idx += kThreads)
      reinterpret_cast<float4*>(dyT)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = threadIdx.x; q < QP; q += kThreads) {
      float m = 0.f, inv = 0.f, delta = 0.f;
      if (q < Q) {
        m = __ldg(row_max + b * Q + q);
        inv = 1.f / __ldg(row_sum + b * Q + q);
        float dsv[kE], smv[kE];
#pragma unroll
        for (int e = 0; e < kE; ++e) { dsv[e] = __ldg(dsb + q * kE + e); smv[e] = __ldg(summary + ((size_t)b * Q + q) * kE + e); }
#pragma unroll
        for (int e = 0; e < kE; ++e) delta = fmaf(dsv[e], smv[e], delta);
      }
      // a[p,q] = exp(y - m) / l = 2^(y log2e + cq),  cq = -m log2e - log2(l)  (-inf for the padded queries: a = 0)
      mq[q] = q < Q ? fmaf(-m, kLog2e, log2f(inv)) : -INFINITY;
      il[q] = delta;
      dl[q] = 0.f;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t swl[8];
  sw128_lane_offsets(lane, swl);
  constexpr uint32_t tm_y = 0, tm_t = QP, tm_dx = 2 * QP, tm_dk = 2 * QP + 32;
  const uint32_t id_t = make_idesc_tf32(128, QP, 1, 0);
  const uint32_t id_32 = make_idesc_tf32(128, 32, 0, 0);
  uint32_t ph_tma = 0, ph_mma = 0, ph_xk = 0, acc_dk = 0;
  uint32_t ph_mma2 = 0;
  bool pending = false;     // the previous tile's d_x rows still sit in TMEM
  int p_prev = 0;
  float prev_old[kE];       // the previous tile's accumulate operands
#pragma unroll
  for (int e = 0; e < kE; ++e) prev_old[e] = 0.f;
  float* dxb = d_x + (size_t)b * kE * n;
  auto store_dx = [&](int pp, const float (&add)[kE]) {
#pragma unroll
    for (int c = 0; c < kE; c += 16) {
      float v[16];
      tmem_ld16(lane_base + tm_dx + c, v);
      tmem_wait_ld();
      if (pp < n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dxb[(size_t)(c + i) * n + pp] = v[i] + add[c + i];
      }
    }
  };
  for (int t = t_begin; t < t_end; ++t) {
    const int p0 = t * kTile;
    const int p = p0 + warp * 32 + lane;
    const bool pin = p < n;
    // accumulate mode: the regression-path d_x written by sql_tc_bwd_pred_kernel is fetched now (32 independent L2
    // loads in flight for the whole tile) and added at the end -- the kernel runs one CTA per SM, registers are free
    float prev[kE];
#pragma unroll
    for (int e = 0; e < kE; ++e) prev[e] = (accumulate && pin) ? __ldcg(dxb + (size_t)e * n + p) : 0.f;
    mbar_wait(s.bar_tma, ph_tma); ph_tma ^= 1;
    split_x_tile(s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      // the previous tile's MMAs read x_k and the dy / a columns of TMEM: done before both are reused
      if (pending) mbar_wait(bar_mma2, ph_mma2);
      if (t + 1 < t_end) issue_x_tma(s, &map_mn, p0 + kTile, b * kE);
      mbar_arrive_expect_tx(bar_xk, kXTile);
#pragma unroll
      for (int j = 0; j < 4; ++j) tma_load_2d(x_k + j * kXBlock, &map_k, p0 + 32 * j, b * kE, bar_xk);
      tc_fence_after();
      issue_y(s, tmem + tm_y, QP);
      const uint32_t xh = smem_u32(s.x_hi), dsa = smem_u32(ds);
#pragma unroll
      for (int k = 0; k < kE / 8; ++k)
        umma_tf32_ss(tmem + tm_t, make_desc_mn32(xh + k * 1024, kXBlock), make_desc_sw128(dsa + k * 32, 16, 1024), id_t,
                     k > 0);
      umma_commit(s.bar_mma);
    }
    if (pending) {     // the previous tile's d_x rows leave TMEM while this tile's y / t MMAs run
      mbar_wait(bar_mma2, ph_mma2); ph_mma2 ^= 1;
      tc_fence_after();
      store_dx(p_prev, prev_old);
      pending = false;
    }
    mbar_wait(s.bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < QP; c += 16) {
      float yv[16], tt[16];
      tmem_ld16(lane_base + tm_y + c, yv);
      tmem_ld16(lane_base + tm_t + c, tt);
      tmem_wait_ld();
      uint8_t* dyw = dyT + warp * 128 * 128;
#pragma unroll
      for (int i4 = 0; i4 < 16; i4 += 4) {
        const float4 cq4 = *reinterpret_cast<const float4*>(mq + c + i4);    // folded exponent offsets
        const float4 dl4 = *reinterpret_cast<const float4*>(il + c + i4);    // delta_q = ds_q . summary_q
        const float cqv[4] = {cq4.x, cq4.y, cq4.z, cq4.w}, dlv[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i4 + j, q = c + i;
          const float a = pin ? ex2_fast(fmaf(yv[i], kLog2e, cqv[j])) : 0.f;
          yv[i] = a * (tt[i] - dlv[j]);
          tt[i] = a;
          *reinterpret_cast<float*>(dyw + q * 128 + swl[q & 7]) = yv[i];
        }
      }
      tmem_st16(lane_base + tm_y + c, yv);
      tmem_st16(lane_base + tm_t + c, tt);
    }
    tmem_wait_st();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      mbar_wait(bar_xk, ph_xk);
      const uint32_t a0 = smem_u32(dyT), xk = smem_u32(x_k), kt = smem_u32(kT), dst = smem_u32(dsT);
#pragma unroll
      for (int k = 0; k < kTile / 8; ++k) {
        umma_tf32_ss(tmem + tm_dk, make_desc_sw128(a0 + (k >> 2) * 128 * 128 + (k & 3) * 32, 16, 1024),
                     make_desc_sw128(xk + (k >> 2) * kXBlock + (k & 3) * 32, 16, 1024), id_32, acc_dk);
        acc_dk = 1;
      }
#pragma unroll
      for (int k = 0; k < QP / 8; ++k)
        umma_tf32_ts(tmem + tm_dx, tmem + tm_y + k * 8,
                     make_desc_sw128(kt + (k >> 2) * 32 * 128 + (k & 3) * 32, 16, 1024), id_32, k > 0);
#pragma unroll
      for (int k = 0; k < QP / 8; ++k)
        umma_tf32_ts(tmem + tm_dx, tmem + tm_t + k * 8,
                     make_desc_sw128(dst + (k >> 2) * 32 * 128 + (k & 3) * 32, 16, 1024), id_32, 1);
      umma_commit(bar_mma2);
    }
    ph_xk ^= 1;
    pending = true;
    p_prev = p;
#pragma unroll
    for (int e = 0; e < kE; ++e) prev_old[e] = prev[e];
  }
  if (pending) {       // drain: the last tile's d_x rows (the commit also covers every d_K MMA)
    mbar_wait(bar_mma2, ph_mma2);
    tc_fence_after();
    store_dx(p_prev, prev_old);
    tc_fence_before();
    __syncthreads();
  }
  {
    const int q = threadIdx.x;
    float* out = part_dK + ((size_t)cta * Q + q) * kE;
    const bool have = t_begin < t_end;
#pragma unroll
    for (int c = 0; c < kE; c += 16) {
      float v[16];
      if (have) {
        tmem_ld16(lane_base + tm_dk + c, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      if (q < Q) {
#pragma unroll
        for (int i = 0; i < 16; ++i) out[c + i] = v[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kCols);
}

// ------------------------------------------------------------------------------------------------
// energy maps  y[b,q,p]   (module-level FullQueryLayer output; also the bring-up kernel of this file)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) sql_tc_energy_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                 const float* __restrict__ queries, int Q, int Qp,
                                                                 int n, int tiles_per_chunk, uint32_t tmem_cols,
                                                                 float* __restrict__ energy) {
  extern __shared__ uint8_t smem_raw[];
  const Smem s = carve(smem_raw, Qp, 0);
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&xmap);
    mbar_init(s.bar_tma, 1);
    mbar_init(s.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(s.tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  __syncthreads();   // barrier init visible before the first TMA
  if (threadIdx.x == 0 && t_begin < t_end) issue_x_tma(s, &xmap, t_begin * kTile, b * kE);
  stage_queries(s, queries + (size_t)b * Q * kE, Q, Qp);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  uint32_t phase = 0;
  for (int t = t_begin; t < t_end; ++t) {
    const int p0 = t * kTile;
    mbar_wait(s.bar_tma, phase);
    split_x_tile(s);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (t + 1 < t_end) issue_x_tma(s, &xmap, p0 + kTile, b * kE);   // x_raw is free again: prefetch
      tc_fence_after();
      issue_y(s, tmem, Qp);
      umma_commit(s.bar_mma);
    }
    mbar_wait(s.bar_mma, phase);
    tc_fence_after();
    const int p = p0 + warp * 32 + lane;
    for (int c = 0; c < Qp; c += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_wait_ld();
      if (p < n) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c + i < Q) energy[((size_t)b * Q + c + i) * n + p] = v[i];
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM and the operand tiles are reused by the next iteration
    phase ^= 1;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace tcsql
}  // namespace sqlx

using namespace sqlx;

namespace {
uint32_t pow2_cols(int c) {
  uint32_t v = 32;
  while ((int)v < c) v <<= 1;
  return v;
}
}  // namespace

// 1 if the tensor-core kernels take this shape (otherwise the caller uses the fp32 kernels)
extern "C" int sqlx_sql_tc_supported(int E, int Q, int D, int n) {
  if (E != tcsql::kE || Q < 1 || Q > 128 || D < 0 || D > 128) return 0;
  if (n % 4 != 0) return 0;   // TMA needs a 16-byte row pitch
  return 1;
}

namespace {
struct TcPlan {
  int Qp, Dp, chunks, tpc;
  uint32_t tmem_cols;
  size_t smem;
};
// ctas_per_sm: how many CTAs of this kernel fit on one SM (shared memory / TMEM), used to size the grid
TcPlan plan_tc(int B, int Q, int D, int n, int tmem_need_cols) {
  TcPlan p;
  p.Qp = (Q + 15) / 16 * 16;
  p.Dp = D > 0 ? (D + 15) / 16 * 16 : 0;
  p.tmem_cols = pow2_cols(tmem_need_cols);
  p.smem = tcsql::smem_bytes(p.Qp, p.Dp);
  int per_sm = (int)((227 * 1024) / (p.smem + 1024));
  const int by_tmem = 512 / (int)p.tmem_cols;
  per_sm = per_sm < by_tmem ? per_sm : by_tmem;
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  const int tiles = ceil_div(n, tcsql::kTile);
  int chunks = (per_sm * kNumSMs) / B;
  chunks = chunks < 1 ? 1 : (chunks > tiles ? tiles : chunks);
  p.tpc = ceil_div(tiles, chunks);
  p.chunks = ceil_div(tiles, p.tpc);
  return p;
}
}  // namespace

namespace sqlx {
// chunk plan of the tensor-core summary kernel (shared with the workspace-size query in sql_fp32.cu)
void tc_summary_plan(int B, int n, int* chunks, int* tiles_per_chunk) {
  const int tiles = ceil_div(n, tcsql::kTileS);
  int c = (2 * kNumSMs) / B;   // 2 CTAs per SM (shared memory 82 KB, 256 TMEM columns each)
  c = c < 1 ? 1 : (c > tiles ? tiles : c);
  *tiles_per_chunk = ceil_div(tiles, c);
  *chunks = ceil_div(tiles, *tiles_per_chunk);
}

// writes per-chunk partial records [B][chunks][Q][34]; the caller runs the split-softmax combine
int tc_summary_partials(const float* x, const float* queries, int B, int Q, int n, float* partial, int* chunks_out,
                        cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  int chunks, tpc;
  tc_summary_plan(B, n, &chunks, &tpc);
  CUtensorMap map_mn, map_k;
  if (int e = make_tensor_map_2d(&map_mn, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = make_tensor_map_2d(&map_k, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 0)) return e;
  if (int e = ensure_dyn_smem(tcsql::sql_tc_summary_kernel, 227 * 1024)) return e;
  {
    ProfScope prof("sql_tc_summary_kernel", st);
    tcsql::sql_tc_summary_kernel<<<dim3(chunks, B), tcsql::kThreads, tcsql::kSmemSBytes, st>>>(map_mn, map_k, queries, Q, n,
                                                                                              tpc, partial);
  }
  *chunks_out = chunks;
  return check_launch("sql_tc_summary_kernel");
}
}  // namespace sqlx

namespace sqlx {
void tc_bwd_plan(int B, int n, int* chunks, int* tiles_per_chunk) {
  const int tiles = ceil_div(n, tcsql::kTile);
  int c = kNumSMs / B;   // one CTA per SM (about 200 KB of shared memory, all 512 TMEM columns)
  c = c < 1 ? 1 : (c > tiles ? tiles : c);
  *tiles_per_chunk = ceil_div(tiles, c);
  *chunks = ceil_div(tiles, *tiles_per_chunk);
}

// ---- mixed-weight decomposition launchers
int tc_pred_mix_fwd(const float* x, const float* Mx, const float* bp, const float* centers, int B, int D, int n,
                    float* pred, cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  const int DP = (D + 15) / 16 * 16;
  const size_t smem = tcsql::smem_p2_bytes(DP);
  const uint32_t cols = pow2_cols(DP);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  per_sm = per_sm > (int)(512 / cols) ? (int)(512 / cols) : per_sm;
  per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);
  const int tiles = ceil_div(n, tcsql::kTile);
  int chunks = (per_sm * kNumSMs) / B;
  chunks = chunks < 1 ? 1 : (chunks > tiles ? tiles : chunks);
  const int tpc = ceil_div(tiles, chunks);
  chunks = ceil_div(tiles, tpc);
  CUtensorMap xmap;
  if (int e = make_tensor_map_2d(&xmap, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = ensure_dyn_smem(tcsql::sql_tc_pred2_kernel, 227 * 1024)) return e;
  ProfScope prof("sql_tc_pred_kernel", st);
  tcsql::sql_tc_pred2_kernel<<<dim3(chunks, B), tcsql::kThreads, smem, st>>>(xmap, Mx, bp, centers, D, DP, n, tpc, cols, pred);
  return check_launch("sql_tc_pred2_kernel");
}

template <int DP>
int launch_bwd_pred(const CUtensorMap& map_mn, const CUtensorMap& map_k, const float* Mx, const float* bp,
                    const float* centers, const float* g_pred, int B, int D, int n, int chunks, int tpc, float* d_x,
                    float* part_dM, float* part_db, float* part_dc, cudaStream_t st) {
  if (int e = ensure_dyn_smem(tcsql::sql_tc_bwd_pred_kernel<DP>, 227 * 1024)) return e;
  ProfScope prof("sql_tc_bwd_pred_kernel", st);
  tcsql::sql_tc_bwd_pred_kernel<DP><<<dim3(chunks, B), tcsql::kThreads, tcsql::BwdPredSmem<DP>::bytes, st>>>(
      map_mn, map_k, Mx, bp, centers, g_pred, D, n, tpc, d_x, part_dM, part_db, part_dc);
  return check_launch("sql_tc_bwd_pred_kernel");
}

int tc_bwd_pred_mix(const float* x, const float* Mx, const float* bp, const float* centers, const float* g_pred, int B,
                    int D, int n, float* d_x, float* part_dM, float* part_db, float* part_dc, int chunks, int tpc,
                    cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  CUtensorMap map_mn, map_k;
  if (int e = make_tensor_map_2d(&map_mn, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = make_tensor_map_2d(&map_k, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 0)) return e;
  if (D <= 64)
    return launch_bwd_pred<64>(map_mn, map_k, Mx, bp, centers, g_pred, B, D, n, chunks, tpc, d_x, part_dM, part_db, part_dc, st);
  return launch_bwd_pred<128>(map_mn, map_k, Mx, bp, centers, g_pred, B, D, n, chunks, tpc, d_x, part_dM, part_db, part_dc, st);
}

template <int QP>
int launch_bwd_sum(const CUtensorMap& map_mn, const CUtensorMap& map_k, const float* queries, const float* summary,
                   const float* row_max, const float* row_sum, const float* d_summary, int B, int Q, int n, int chunks,
                   int tpc, int accumulate, float* d_x, float* part_dK, cudaStream_t st) {
  if (int e = ensure_dyn_smem(tcsql::sql_tc_bwd_sum_kernel<QP>, 227 * 1024)) return e;
  ProfScope prof("sql_tc_bwd_sum_kernel", st);
  tcsql::sql_tc_bwd_sum_kernel<QP><<<dim3(chunks, B), tcsql::kThreads, tcsql::BwdSumSmem<QP>::bytes, st>>>(
      map_mn, map_k, queries, summary, row_max, row_sum, d_summary, Q, n, tpc, accumulate, d_x, part_dK);
  return check_launch("sql_tc_bwd_sum_kernel");
}

int tc_bwd_sum(const float* x, const float* queries, const float* summary, const float* row_max, const float* row_sum,
               const float* d_summary, int B, int Q, int n, int accumulate, float* d_x, float* part_dK, int chunks, int tpc,
               cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  CUtensorMap map_mn, map_k;
  if (int e = make_tensor_map_2d(&map_mn, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = make_tensor_map_2d(&map_k, x, (uint64_t)B * tcsql::kE, (uint64_t)n, 32, 32, 0)) return e;
  if (Q <= 64)
    return launch_bwd_sum<64>(map_mn, map_k, queries, summary, row_max, row_sum, d_summary, B, Q, n, chunks, tpc, accumulate,
                              d_x, part_dK, st);
  return launch_bwd_sum<128>(map_mn, map_k, queries, summary, row_max, row_sum, d_summary, B, Q, n, chunks, tpc, accumulate,
                             d_x, part_dK, st);
}
}  // namespace sqlx

extern "C" int sqlx_sql_energy_tc(const float* x, const float* queries, int B, int E, int Q, int n, float* energy,
                                  void* stream) {
  SQLX_REQUIRE(x && queries && energy, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape B=%d n=%d", B, n);
  SQLX_REQUIRE(sqlx_sql_tc_supported(E, Q, 0, n), "shape E=%d Q=%d n=%d is not supported by the tensor-core path", E, Q, n);
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const TcPlan p = plan_tc(B, Q, 0, n, (Q + 15) / 16 * 16);
  CUtensorMap xmap;
  if (int e = make_tensor_map_2d(&xmap, x, (uint64_t)B * E, (uint64_t)n, 32, 32, /*atom32=*/1)) return e;
  if (int e = ensure_dyn_smem(tcsql::sql_tc_energy_kernel, 227 * 1024)) return e;
  ProfScope prof("sql_tc_energy_kernel", st);
  tcsql::sql_tc_energy_kernel<<<dim3(p.chunks, B), tcsql::kThreads, p.smem, st>>>(xmap, queries, Q, p.Qp, n, p.tpc,
                                                                                 p.tmem_cols, energy);
  return check_launch("sql_tc_energy_kernel");
}
