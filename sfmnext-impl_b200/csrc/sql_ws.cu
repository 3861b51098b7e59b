// SQL decoder tail on tcgen05 / TMEM / TMA, warp-specialised generation (round 2).
//
// Round 1's kernels ran every phase of a tile in series on ONE warpgroup: TMA wait -> hi/lo split -> tcgen05.mma -> wait ->
// per-thread softmax epilogue out of TMEM -> tcgen05.mma -> wait.  With one warp per scheduler the epilogue paid the full
// latency of every instruction and the tensor pipe idled under it (ncu, round 1: 6 % warps active, 11-25 % tensor pipe).
// Here a CTA is 17 warps with fixed roles:
//
//   warp 16 (one elected lane)   issues every TMA load and every tcgen05.mma, commits them to mbarriers; never touches data
//   warps 0..15                  TWO GROUPS of eight epilogue warps that ping-pong: group g owns the steps s = g (mod 2) and
//                                the TMEM accumulator Z[g], so one group's TMEM loads / exponentials / stores overlap the
//                                other group's barrier waits and the tensor pipe works on both.  tcgen05.ld / st reach the 32
//                                TMEM lanes 32 * (warp % 4): warps q and q + 4 of a group own the same 32 pixels and split
//                                the COLUMNS (bins / queries) in two
//
// i.e. the logits accumulator is DOUBLE-BUFFERED in TMEM and every wait of one step is covered by work of the other.  A step is (128-pixel tile, 64-column half): with the per-pixel softmax statistics saved
// by the forward kernel (max and 1 / sum of the base-2 logits, 8 bytes per pixel) the backward needs no max / sum pass and
// the halves of a 128-bin row are independent, so one kernel body serves D, Q <= 64 (one half) and <= 128 (two halves).
//
// Precision is unchanged from round 1: forward contractions 3xTF32 (hi/lo split of both operands), gradient contractions
// single-pass TF32 with fp32 accumulation.
//
// Reference lines: networks/layers.py:17-20, networks/depth_decoder_QTR.py:61,70 and their autograd.
#include "common.cuh"
#include "tc_common.cuh"

#include <math.h>

#include <atomic>

namespace sqlx {
namespace wsql {

using namespace tc;

constexpr int kE = 32;                 // embedding channels
constexpr int kTile = 128;             // pixels per tile = UMMA M
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGrpWarps = 8;            // two ping-pong groups of epilogue warps
constexpr int kGrpThreads = kGrpWarps * 32;
constexpr int kThreads = kEpiThreads + 32;      // + the control warp.  17 warps: one scheduler hosts five, its 16 K registers
                                                // cap every thread at 96 registers (what __launch_bounds__ enforces)
constexpr int kXBlock = 32 * kE * 4;   // one [32 e][32 px] block: 4 KB
constexpr int kXTile = 4 * kXBlock;    // 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kAccN = kE + 16;         // [x | 1 | zero padding] columns of the pixel-contraction accumulator

__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// named barrier over `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// all epilogue threads: [rows][32] fp32 global matrix -> K-major SW128 hi (and lo) tiles with `rows_pad` rows
__device__ __forceinline__ void stage_kmajor(uint8_t* hi, uint8_t* lo, const float* __restrict__ src, int rows, int rows_pad,
                                             int tid, int nthreads) {
  for (int base = tid; base < rows_pad * kE; base += 4 * nthreads) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + j * nthreads;
      v[j] = (idx < rows_pad * kE && (idx >> 5) < rows) ? __ldg(src + idx) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + j * nthreads;
      if (idx >= rows_pad * kE) break;
      const uint32_t off = sw128_offset(idx >> 5, idx & 31);
      const float h = lo ? tf32_hi(v[j]) : v[j];
      *reinterpret_cast<float*>(hi + off) = h;
      if (lo) *reinterpret_cast<float*>(lo + off) = v[j] - h;
    }
  }
}

// all epilogue threads: transposed [32 e rows][rows_pad cols] in 32-column SW128 atoms (single precision value)
__device__ __forceinline__ void stage_transposed(uint8_t* dst, const float* __restrict__ src, int rows, int rows_pad, int tid,
                                                 int nthreads) {
  for (int base = tid; base < rows_pad * kE; base += 4 * nthreads) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + j * nthreads;
      v[j] = (idx < rows_pad * kE && (idx >> 5) < rows) ? __ldg(src + idx) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = base + j * nthreads;
      if (idx >= rows_pad * kE) break;
      const int r = idx >> 5, e = idx & 31;
      *reinterpret_cast<float*>(dst + (uint32_t)(r >> 5) * 32u * 128u + sw128_offset(e, r & 31)) = v[j];
    }
  }
}

// One group of epilogue warps (NT threads): x_lo = x - tf32(x) of a freshly landed tile (same swizzled layout: elementwise).
// The landing buffer itself is the "hi" operand: the tensor core reads only the tf32 part of a 32-bit operand (the 13 low
// mantissa bits are ignored -- measured: forward depth identical to the explicitly truncated operand to 2e-6,
// tests/test_sql_tc_gpu.py::test_ws_kernels_match_v1), so it needs no pass of its own.  SQLX_WS_TRUNC=1 truncates it in place.
#ifndef SQLX_WS_TRUNC
#define SQLX_WS_TRUNC 0
#endif
template <int NT>
__device__ __forceinline__ void split_x(uint8_t* x_hi, uint8_t* x_lo, int tid) {
  float4* hi = reinterpret_cast<float4*>(x_hi);
  float4* lo = reinterpret_cast<float4*>(x_lo);
  constexpr int PER = kXTile / 16 / NT;
  float4 v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) v[i] = hi[tid + i * NT];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    float4 h, l;
    h.x = tf32_hi(v[i].x); h.y = tf32_hi(v[i].y); h.z = tf32_hi(v[i].z); h.w = tf32_hi(v[i].w);
    l.x = v[i].x - h.x; l.y = v[i].y - h.y; l.z = v[i].z - h.z; l.w = v[i].w - h.w;
    if (SQLX_WS_TRUNC) hi[tid + i * NT] = h;
    lo[tid + i * NT] = l;
  }
}

// control lane: one 128-pixel x tile = four [32 e][32 px] boxes (pixels >= n are zero-filled by TMA)
__device__ __forceinline__ void tma_x_tile(uint8_t* dst, uint32_t atom_stride, const CUtensorMap* map, int p0, int row0,
                                           uint64_t* bar) {
  mbar_arrive_expect_tx(bar, kXTile);
#pragma unroll
  for (int j = 0; j < 4; ++j) tma_load_2d(dst + j * atom_stride, map, p0 + 32 * j, row0, bar);
}

// control lane: D[128 px, N] = x^T B^T as 3xTF32; x tiles MN-major (hi / lo), B tiles K-major rows (hi / lo).
// The ONE issuing thread is on the critical path of every step, so it works from descriptors built once before the tile
// loop: a UMMA shared-memory descriptor holds (address >> 4) in its low 14 bits, and every operand of a step sits at
// base + a compile-time byte offset inside the CTA's < 256 KB window, so "descriptor + (offset >> 4)" is one 64-bit add
// (ncu, first warp-specialised cut: 39 % of the epilogue warps' samples sat in mbarrier waits behind a control lane that
// rebuilt 72 descriptors -- shifts, masks, ors -- per step).
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

__device__ __forceinline__ void mma_x_b3(uint64_t dxh, uint64_t dxl, uint64_t dbh, uint64_t dbl, uint32_t tm_d, uint32_t idesc) {
  uint32_t acc = 0;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t xa = pass == 1 ? dxl : dxh;
    const uint64_t bb = pass == 2 ? dbl : dbh;
#pragma unroll
    for (int k = 0; k < kE / 8; ++k) {
      // A: MN-major, 8 e-rows = 1024 B per k-step; B: K-major, 8 e = 32 B per k-step inside the 128-B row
      umma_tf32_ss(tm_d, desc_add(xa, k * 1024), desc_add(bb, k * 32), idesc, acc);
      acc = 1;
    }
  }
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles: afterwards lanes 2i and 2i+1 hold the total of v[i]
// in v[0].
__device__ __forceinline__ void warp_reduce16(float (&v)[16]) {
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

// ================================================================================================================
// forward: pred[b,p] = sum_d softmax_d(M x + b)[d] c_d, and the per-pixel softmax statistics for the backward
//   TMEM: Z[0] = [0, DP), Z[1] = [DP, 2 DP)
// ================================================================================================================
constexpr int kPredStages = 6;     // TMA ring: a slot is refilled when its tile's MMA completes, 4 tile times before it is needed

template <int DP>
struct PredSmem {
  static constexpr size_t x_ring = 0, x_lo = kPredStages * kXTile, m_hi = x_lo + 2 * kXTile, m_lo = m_hi + DP * 128,
                          part = m_lo + DP * 128,                   // [4 column groups][128 px] float4 (m, se, sc, -)
                          tail = part + 4 * 128 * 16;
  static constexpr size_t bytes = 1024 + tail + 2 * DP * 4 + 256;
};

template <int DP>
__global__ void __launch_bounds__(kThreads, 1) sql_ws_pred_kernel(const __grid_constant__ CUtensorMap map_mn,
                                                                  const float* __restrict__ Mx, const float* __restrict__ bp,
                                                                  const float* __restrict__ centers, int D, int n,
                                                                  int tiles_per_chunk, float* __restrict__ pred,
                                                                  float* __restrict__ stat_m, float* __restrict__ stat_inv) {
  using L = PredSmem<DP>;
  constexpr int NST = kPredStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* x_ring = base + L::x_ring;     // [NST] landing buffers = "hi" operands after the in-place truncation
  uint8_t* x_lo = base + L::x_lo;         // [2]
  uint8_t* m_hi = base + L::m_hi;
  uint8_t* m_lo = base + L::m_lo;
  float4* part = reinterpret_cast<float4*>(base + L::part);
  float* bias2 = reinterpret_cast<float*>(base + L::tail);
  float* cen = bias2 + DP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + DP);
  uint64_t* bar_full = bars;                 // [NST] TMA: tile landed in ring slot
  uint64_t* bar_split = bars + NST;          // [2]   16 warps: slot truncated in place, x_lo[i] written
  uint64_t* bar_z = bars + NST + 2;          // [2]   MMA commit: Z[i] ready (ring slot and x_lo[i] free)
  uint64_t* bar_zfree = bars + NST + 4;      // [2]   16 warps: Z[i] read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NST + 6);
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 2 * DP;
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  const int nsteps = max(t_end - t_begin, 0);
  if (threadIdx.x == kEpiThreads) {
    tma_prefetch_desc(&map_mn);
    for (int i = 0; i < NST; ++i) mbar_init(bar_full + i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_split + i, kGrpWarps); mbar_init(bar_z + i, 1); mbar_init(bar_zfree + i, kGrpWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {
    tmem_alloc(tmem_slot, kCols);
    tmem_relinquish();
  }
  __syncthreads();
  if (threadIdx.x == kEpiThreads)
    for (int i = 0; i < NST && i < nsteps; ++i)
      tma_x_tile(x_ring + i * kXTile, kXBlock, &map_mn, (t_begin + i) * kTile, b * kE, bar_full + i);
  if (warp < kEpiWarps) {
    stage_kmajor(m_hi, m_lo, Mx + (size_t)b * D * kE, D, DP, threadIdx.x, kEpiThreads);
    for (int d = threadIdx.x; d < DP; d += kEpiThreads) {
      bias2[d] = d < D ? __ldg(bp + d) * kLog2e : -INFINITY;   // base-2 logits; padded bins never win the softmax
      cen[d] = d < D ? __ldg(centers + (size_t)b * D + d) : 0.f;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kEpiWarps) {
    // ---------------------------------------------------------------- control lane
    if (elect_one()) {
      const uint64_t dx_ring = make_desc_mn32(smem_u32(x_ring), kXBlock), dx_lo = make_desc_mn32(smem_u32(x_lo), kXBlock);
      const uint64_t dm_hi = make_desc_sw128(smem_u32(m_hi), 16, 1024), dm_lo = make_desc_sw128(smem_u32(m_lo), 16, 1024);
      const uint32_t idesc = make_idesc_tf32(kTile, DP, 1, 0);
      int slot = 0;
      for (int s = 0; s < nsteps; ++s) {
        mbar_wait(bar_split + (s & 1), (s >> 1) & 1);                     // operands of tile s are ready
        if (s >= 2) mbar_wait(bar_zfree + (s & 1), ((s >> 1) - 1) & 1);   // the epilogue of step s-2 has read Z[s&1]
        tc_fence_after();
        mma_x_b3(desc_add(dx_ring, slot * kXTile), desc_add(dx_lo, (s & 1) * kXTile), dm_hi, dm_lo, tmem + (s & 1) * DP, idesc);
        umma_commit(bar_z + (s & 1));
        slot = slot + 1 == NST ? 0 : slot + 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps: group `grp` owns tiles s = grp (mod 2)
    const int grp = warp >> 3, wl = warp & 7, q = wl & 3, cg = wl >> 2, gtid = threadIdx.x & (kGrpThreads - 1);
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    constexpr int CW = DP / 2;                     // columns per thread and tile
    float4* gpart = part + grp * 2 * 128;          // [2 column halves][128 px] of this group
    auto split_tile = [&](int i) {                 // tile i (same parity as the group): lo -> x_lo[i & 1]
      mbar_wait(bar_full + i % NST, (i / NST) & 1);
      split_x<kGrpThreads>(x_ring + (i % NST) * kXTile, x_lo + (i & 1) * kXTile, gtid);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split + (i & 1));
    };
    if (grp < nsteps) split_tile(grp);
    for (int s = grp; s < nsteps; s += 2) {
      const int p = (t_begin + s) * kTile + q * 32 + lane;
      mbar_wait(bar_z + grp, (s >> 1) & 1);              // logits of tile s are in Z[grp]; its ring slot and x_lo[grp] are free
      tc_fence_after();
      if (gtid == 0 && s + NST < nsteps)                 // refill the slot: NST - 2 tile times before that tile is split
        tma_x_tile(x_ring + (s % NST) * kXTile, kXBlock, &map_mn, (t_begin + s + NST) * kTile, b * kE, bar_full + s % NST);
      if (s + 2 < nsteps) split_tile(s + 2);             // operands of this group's next tile
      // this thread's CW bins of its pixel: online max / sum / expectation in base 2
      float m = -INFINITY, se = 0.f, sc = 0.f;
#pragma unroll
      for (int c = 0; c < CW; c += 32) {
        float v[32];
        tmem_ld16(lane_base + grp * DP + cg * CW + c, *reinterpret_cast<float(*)[16]>(v));
        tmem_ld16(lane_base + grp * DP + cg * CW + c + 16, *reinterpret_cast<float(*)[16]>(v + 16));
        tmem_wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(bias2 + cg * CW + c);
        const float4* c4 = reinterpret_cast<const float4*>(cen + cg * CW + c);
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bq = b4[j];
          v[4 * j] = fmaf(v[4 * j], kLog2e, bq.x); v[4 * j + 1] = fmaf(v[4 * j + 1], kLog2e, bq.y);
          v[4 * j + 2] = fmaf(v[4 * j + 2], kLog2e, bq.z); v[4 * j + 3] = fmaf(v[4 * j + 3], kLog2e, bq.w);
          cm = fmaxf(cm, fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3])));
        }
        if (cm > m) { const float r = ex2f(m - cm); se *= r; sc *= r; m = cm; }
        const float ms = fmaxf(m, -1e30f);     // every bin so far padding (-inf): keep the exponents -inf, not NaN
        float se2 = 0.f, sc2 = 0.f;            // two accumulation chains
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 cq = c4[j];
          const float e0 = ex2f(v[4 * j] - ms), e1 = ex2f(v[4 * j + 1] - ms);
          const float e2 = ex2f(v[4 * j + 2] - ms), e3 = ex2f(v[4 * j + 3] - ms);
          if (j & 1) { se2 += (e0 + e1) + (e2 + e3); sc2 += fmaf(e0, cq.x, e1 * cq.y) + fmaf(e2, cq.z, e3 * cq.w); }
          else { se += (e0 + e1) + (e2 + e3); sc += fmaf(e0, cq.x, e1 * cq.y) + fmaf(e2, cq.z, e3 * cq.w); }
        }
        se += se2; sc += sc2;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_zfree + grp);       // Z[grp] may be overwritten by the MMA of tile s+2
      // combine the two column halves of the pixel (warps q and q + 4 of the group) through shared memory
      gpart[cg * 128 + q * 32 + lane] = make_float4(m, se, sc, 0.f);
      named_sync(1 + grp * 4 + q, 64);
      const float4 pa = gpart[q * 32 + lane], pb = gpart[128 + q * 32 + lane];
      const float M = fmaxf(pa.x, pb.x);
      const float ra = ex2f(pa.x - M), rb = ex2f(pb.x - M);   // (a half whose bins are all padding has m = -inf, se = 0)
      const float SE = fmaf(pa.y, ra, pb.y * rb), SC = fmaf(pa.z, ra, pb.z * rb);
      named_sync(1 + grp * 4 + q, 64);                   // `gpart` is rewritten by the group's next tile
      if (p < n) {
        const float inv = 1.f / SE;
        if (cg == 0) { pred[(size_t)b * n + p] = SC * inv; stat_m[(size_t)b * n + p] = M; }
        else stat_inv[(size_t)b * n + p] = inv;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, kCols);
}

// ================================================================================================================
// backward pass 1 (regression path): d_x = dz M,  dM = dz^T x,  d_bp = sum_p dz,  d_centers = sum_p pi g
//   dz[p,d] = pi[p,d] g[p] (c_d - pred[p]),  pi = 2^(t - m2) * inv  from the saved statistics (no max / sum pass)
//   step = (tile, 64-bin half).  TMEM: Z[0] = [0,64)  Z[1] = [64,128)  dx[0..2] = [128 + 32 i, +32)
//                                      acc[h] = [224 + 48 h, +48)   rows = bins of half h, cols = [e | ones]
// ================================================================================================================
template <int DP>
struct BwdPredSmem {
  static constexpr int NH = DP / 64;
  static constexpr int NST = DP > 64 ? 2 : 3;      // TMA ring of x tiles (landing buffer = "hi" operand, truncated in place)
  static constexpr size_t x_ring = 0, x_lo = NST * kXTile, m_hi = x_lo + kXTile, m_lo = m_hi + DP * 128,
                          mT = m_lo + DP * 128,                     // [DP/32 atoms][32 e rows][32 d]
                          bx = mT + (DP / 32) * 32 * 128,           // [2 buffers][4 px atoms][48 rows][32 px]: x (TMA) | ones
                          adz = bx + 2 * 4 * kAccN * 128,           // [4 px atoms][128 rows][32 px]: dz^T, rows >= 64 zero
                          tail = adz + 4 * 128 * 128;
  static constexpr size_t bytes = 1024 + tail + 2 * DP * 4 + 128 + kEpiWarps * 32 * 4;
  static_assert(bytes <= 227 * 1024, "shared-memory budget");
};

template <int DP>
__global__ void __launch_bounds__(kThreads, 1) sql_ws_bwd_pred_kernel(
    const __grid_constant__ CUtensorMap map_mn, const __grid_constant__ CUtensorMap map_k, const float* __restrict__ Mx,
    const float* __restrict__ bp, const float* __restrict__ centers, const float* __restrict__ g_pred,
    const float* __restrict__ pred, const float* __restrict__ stat_m, const float* __restrict__ stat_inv, int D, int n,
    int tiles_per_chunk, float* __restrict__ d_x, float* __restrict__ part_dM /*[cta][D][32]*/,
    float* __restrict__ part_db /*[cta][D]*/, float* __restrict__ part_dc /*[cta][D]*/) {
  using L = BwdPredSmem<DP>;
  constexpr int NH = L::NH, NST = L::NST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* x_ring = base + L::x_ring;
  uint8_t* x_lo = base + L::x_lo;
  uint8_t* m_hi = base + L::m_hi;
  uint8_t* m_lo = base + L::m_lo;
  uint8_t* mT = base + L::mT;
  uint8_t* bx = base + L::bx;
  uint8_t* adz = base + L::adz;
  float* bias2 = reinterpret_cast<float*>(base + L::tail);
  float* cen = bias2 + DP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + DP);
  uint64_t* bar_full = bars;            // [NST <= 3] TMA: x tile landed in its ring slot
  uint64_t* bar_split = bars + 3;       // 16 warps: slot truncated in place, x_lo written
  uint64_t* bar_bx = bars + 4;          // [2] TMA: K-major x rows of bx[i] landed
  uint64_t* bar_z = bars + 6;           // [2] MMA commit: logits of a step in Z[i]
  uint64_t* bar_epi = bars + 8;         // [2] 16 warps: dz of a step in Z[i] (TMEM) and adz (shared)
  uint64_t* bar_m2 = bars + 10;         // [2] MMA commit: d_x / accumulator MMAs of a step done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* dc_part = reinterpret_cast<float*>(bars + 14);   // [16 warps][32]
  const int b = blockIdx.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 512;
  constexpr uint32_t tm_z = 0, tm_dx = 128, tm_acc = 224;
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  const int ntiles = max(t_end - t_begin, 0);
  const int nsteps = ntiles * NH;
  if (threadIdx.x == kEpiThreads) {
    tma_prefetch_desc(&map_mn);
    tma_prefetch_desc(&map_k);
    for (int i = 0; i < NST; ++i) mbar_init(bar_full + i, 1);
    mbar_init(bar_split, kGrpWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_bx + i, 1); mbar_init(bar_z + i, 1); mbar_init(bar_epi + i, kGrpWarps); mbar_init(bar_m2 + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {
    tmem_alloc(tmem_slot, kCols);
    tmem_relinquish();
  }
  __syncthreads();
  if (threadIdx.x == kEpiThreads && ntiles > 0) {
    for (int i = 0; i < NST && i < ntiles; ++i)
      tma_x_tile(x_ring + i * kXTile, kXBlock, &map_mn, (t_begin + i) * kTile, b * kE, bar_full + i);
    tma_x_tile(bx, kAccN * 128, &map_k, t_begin * kTile, b * kE, bar_bx);
  }
  if (warp < kEpiWarps) {
    const float* Mb = Mx + (size_t)b * D * kE;
    stage_kmajor(m_hi, m_lo, Mb, D, DP, threadIdx.x, kEpiThreads);
    stage_transposed(mT, Mb, D, DP, threadIdx.x, kEpiThreads);
    for (int i = threadIdx.x; i < 2 * 4 * 16 * 32; i += kEpiThreads) {     // rows 32..47 of every pixel atom: ones | zeros
      const int atom = i / (16 * 32), rem = i - atom * 16 * 32, row = kE + (rem >> 5), col = rem & 31;
      *reinterpret_cast<float*>(bx + atom * kAccN * 128 + sw128_offset(row, col)) = row == kE ? 1.f : 0.f;
    }
    for (int i = threadIdx.x; i < 4 * 128 * 32 / 4; i += kEpiThreads)
      reinterpret_cast<float4*>(adz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = threadIdx.x; d < DP; d += kEpiThreads) {
      bias2[d] = d < D ? __ldg(bp + d) * kLog2e : -INFINITY;
      cen[d] = d < D ? __ldg(centers + (size_t)b * D + d) : 0.f;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  float* dxb = d_x + (size_t)b * kE * n;

  if (warp == kEpiWarps) {
    // ---------------------------------------------------------------- control lane
    if (elect_one()) {
      const uint32_t id_acc = make_idesc_tf32(64, kAccN, 0, 0);      // M = 64: only the 64 bin rows of the half are read (the
                                                                      // A read is what an SS instruction costs: 44 -> 28 cycles)
      const uint32_t id_dx = make_idesc_tf32(128, 32, 0, 0);
      const uint32_t id_z = make_idesc_tf32(kTile, 64, 1, 0);
      const uint64_t dx_ring = make_desc_mn32(smem_u32(x_ring), kXBlock), dx_lo = make_desc_mn32(smem_u32(x_lo), kXBlock);
      const uint64_t dm_hi = make_desc_sw128(smem_u32(m_hi), 16, 1024), dm_lo = make_desc_sw128(smem_u32(m_lo), 16, 1024);
      const uint64_t d_mT = make_desc_sw128(smem_u32(mT), 16, 1024), d_adz = make_desc_sw128(smem_u32(adz), 16, 1024);
      const uint64_t d_bx = make_desc_sw128(smem_u32(bx), 16, 1024);
      // d_x and accumulator MMAs of step sp = (tile ti, half h): its dz is in Z[sp&1] (TMEM) and adz (shared)
      auto issue_mma2 = [&](int sp, int ti, int h) {
        mbar_wait(bar_epi + (sp & 1), (sp >> 1) & 1);
        if (h == 0) mbar_wait(bar_bx + (ti & 1), (ti >> 1) & 1);
        tc_fence_after();
        const uint64_t mt = desc_add(d_mT, h * 2 * 32 * 128), b0 = desc_add(d_bx, (ti & 1) * 4 * kAccN * 128);
        const uint32_t t_dx = tmem + tm_dx + (ti % 3) * 32, t_z = tmem + tm_z + (sp & 1) * 64, t_acc = tmem + tm_acc + h * kAccN;
#pragma unroll
        for (int k = 0; k < 64 / 8; ++k)          // d_x tile (+)= dz_h M_h        (K = 64 bins, A from TMEM)
          umma_tf32_ts(t_dx, t_z + k * 8, desc_add(mt, (k >> 2) * 32 * 128 + (k & 3) * 32), id_dx, (h > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k)       // acc_h += dz_h^T [x | 1]       (K = 128 pixels)
          umma_tf32_ss(t_acc, desc_add(d_adz, (k >> 2) * 128 * 128 + (k & 3) * 32),
                       desc_add(b0, (k >> 2) * kAccN * 128 + (k & 3) * 32), id_acc, (ti > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_m2 + (sp & 1));
      };
      int ti = 0, h = 0, slot = 0, pti = 0, ph = 0;     // (tile, half, ring slot) of step s; (tile, half) of step s-1
      for (int s = 0; s < nsteps; ++s) {
        if (h == 0) mbar_wait(bar_split, ti & 1);   // ring slot ti % NST and x_lo hold the operands of tile ti
        tc_fence_after();
        // logits of (tile, half) -> Z[s&1].  Its previous contents (dz of step s-2) were last read by the d_x MMA of step
        // s-2, issued before this one: tcgen05.mma of one thread execute in issue order.
        mma_x_b3(desc_add(dx_ring, slot * kXTile), dx_lo, desc_add(dm_hi, h * 64 * 128), desc_add(dm_lo, h * 64 * 128),
                 tmem + tm_z + (s & 1) * 64, id_z);
        umma_commit(bar_z + (s & 1));
        if (s > 0) issue_mma2(s - 1, pti, ph);
        pti = ti; ph = h;
        if (++h == NH) { h = 0; ++ti; slot = slot + 1 == NST ? 0 : slot + 1; }
      }
      if (nsteps > 0) issue_mma2(nsteps - 1, pti, ph);
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps: group `grp` owns steps s = grp (mod 2)
    // (two halves per tile: group h owns half h of every tile; one half: the groups alternate tiles)
    const int grp = warp >> 3, wl = warp & 7, q = wl & 3, cg = wl >> 2, gtid = threadIdx.x & (kGrpThreads - 1);
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t lane_hi = (uint32_t)lane >> 2, lane_lo4 = ((uint32_t)lane & 3u) << 2;   // swizzled column of this lane
    float dc[32];                          // d_centers partials of this thread's 32 bins (the same bins at every step)
#pragma unroll
    for (int i = 0; i < 32; ++i) dc[i] = 0.f;
    auto store_dx = [&](int ti) {          // this thread's 16 channels of its pixel of tile ti: TMEM -> global
      const int pp = (t_begin + ti) * kTile + q * 32 + lane;
      float v[16];
      tmem_ld16(lane_base + tm_dx + (ti % 3) * 32 + cg * 16, v);
      tmem_wait_ld();
      if (pp < n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dxb[(size_t)(cg * 16 + i) * n + pp] = v[i];
      }
    };
    auto split_tile = [&](int i) {         // tile i: lo -> x_lo (the ring slot itself is the hi operand)
      mbar_wait(bar_full + i % NST, (i / NST) & 1);
      split_x<kGrpThreads>(x_ring + (i % NST) * kXTile, x_lo, gtid);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split);
    };
    if (grp == 0 && ntiles > 0) split_tile(0);
    for (int s = grp; s < nsteps; s += 2) {
      const int ti = s / NH, h = s - ti * NH;
      const int p = (t_begin + ti) * kTile + q * 32 + lane;
      // per-pixel terms: in flight while waiting for the logits
      const bool pin = p < n;
      const size_t o = (size_t)b * n + p;
      const float g = pin ? __ldg(g_pred + o) : 0.f;
      const float inv = pin ? __ldg(stat_inv + o) : 0.f;
      const float m2 = pin ? __ldg(stat_m + o) : 0.f;
      const float pr = pin ? __ldg(pred + o) : 0.f;
      const float gi = g * inv;
      mbar_wait(bar_z + grp, (s >> 1) & 1);                // logits of step s in Z[grp]
      tc_fence_after();
      if (h == NH - 1) {                                   // the tile's last logits MMA is done: its ring slot and x_lo are free
        if (gtid == 0 && ti + NST < ntiles)
          tma_x_tile(x_ring + (ti % NST) * kXTile, kXBlock, &map_mn, (t_begin + ti + NST) * kTile, b * kE, bar_full + ti % NST);
        if (ti + 1 < ntiles) split_tile(ti + 1);           // next tile's operands first: its MMA overlaps this epilogue
      }
      // ---- dz of this thread's 32 bins of its pixel: TMEM -> registers -> TMEM (the A operand of the d_x MMA)
      float v[32];
      {
        const int c0 = h * 64 + cg * 32;
        tmem_ld16(lane_base + tm_z + grp * 64 + cg * 32, *reinterpret_cast<float(*)[16]>(v));
        tmem_ld16(lane_base + tm_z + grp * 64 + cg * 32 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
        tmem_wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(bias2 + c0);
        const float4* c4 = reinterpret_cast<const float4*>(cen + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bq = b4[j], cq = c4[j];
          const float bv[4] = {bq.x, bq.y, bq.z, bq.w}, cv[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * j + k;
            const float pg = ex2f(fmaf(v[i], kLog2e, bv[k]) - m2) * gi;
            dc[i] += pg;
            v[i] = pg * (cv[k] - pr);
          }
        }
        tmem_st16(lane_base + tm_z + grp * 64 + cg * 32, *reinterpret_cast<float(*)[16]>(v));
        tmem_st16(lane_base + tm_z + grp * 64 + cg * 32 + 16, *reinterpret_cast<float(*)[16]>(v + 16));
      }
      // ---- the transposed copy (A operand of the accumulator MMA) goes to the single adz tile: wait until the MMAs of
      // the previous step -- the other group's -- have read it.  Everything above overlapped them.
      if (s > 0) {
        mbar_wait(bar_m2 + ((s - 1) & 1), ((s - 1) >> 1) & 1);
        tc_fence_after();
        if (h == 0 && gtid == 0 && ti + 1 < ntiles)        // all MMAs of tile ti-1 are done: bx[(ti+1)&1] is free, refill it
          tma_x_tile(bx + (size_t)((ti + 1) & 1) * 4 * kAccN * 128, kAccN * 128, &map_k, (t_begin + ti + 1) * kTile, b * kE,
                     bar_bx + ((ti + 1) & 1));
      } else if (gtid == 0 && ntiles > 1) {
        tma_x_tile(bx + (size_t)4 * kAccN * 128, kAccN * 128, &map_k, (t_begin + 1) * kTile, b * kE, bar_bx + 1);
      }
      {
        uint8_t* adw = adz + q * 128 * 128;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int row = cg * 32 + i;
          *reinterpret_cast<float*>(adw + row * 128 + ((lane_hi ^ (uint32_t)(row & 7)) << 4) + lane_lo4) = v[i];
        }
      }
      tmem_wait_st();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_epi + grp);
      // the previous tile's d_x rows (complete: bar_m2 of its last step was waited for above) leave TMEM now
      if (s > 0 && h == 0) store_dx(ti - 1);
    }
    if (nsteps > 0 && grp == ((nsteps - 1) & 1)) {         // the group of the last step drains the last tile
      mbar_wait(bar_m2 + ((nsteps - 1) & 1), ((nsteps - 1) >> 1) & 1);
      tc_fence_after();
      store_dx(ntiles - 1);
    }
    // every MMA has completed before the accumulators are read: the last commit covers all earlier ones
    named_sync(1, kEpiThreads);
    tc_fence_after();
    // ---- accumulators: rows = bins of half h, cols [0,32) dM, col 32 d_bp.  An M = 64 accumulator keeps row m in TMEM lane
    // 32 (m / 16) + m % 16 (tools/tc_probe3.cu): lanes 0..15 of every lane quarter.  Warps (q, cg) of group 0 read half 0,
    // of group 1 half 1 (one half: group 0 only); three 16-column chunks per warp pair
    {
      const int h = grp;
      if (h < NH) {
        const int d = lane < 16 ? h * 64 + q * 16 + lane : D;
        for (int ch = cg; ch < 3; ch += 2) {
          float a[16];
          if (ntiles > 0) {
            tmem_ld16(lane_base + tm_acc + h * kAccN + ch * 16, a);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = 0.f;
          }
          if (d < D) {
            if (ch < 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) part_dM[((size_t)cta * D + d) * kE + ch * 16 + i] = a[i];
            } else {
              part_db[(size_t)cta * D + d] = a[0];
            }
          }
        }
      }
    }
    // ---- d_centers: warp totals of this thread's 32 bins, then the pixel quarters (and, with one half, the two groups)
    {
      float lo16[16], hi16[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { lo16[i] = dc[i]; hi16[i] = dc[16 + i]; }
      warp_reduce16(lo16);
      warp_reduce16(hi16);
      if (!(lane & 1)) {
        dc_part[warp * 32 + (lane >> 1)] = lo16[0];
        dc_part[warp * 32 + 16 + (lane >> 1)] = hi16[0];
      }
    }
    named_sync(1, kEpiThreads);
    for (int i = threadIdx.x; i < NH * 64; i += kEpiThreads) {
      const int h = i >> 6, c = i & 63, cgc = c >> 5, k = c & 31;
      float acc = 0.f;
      if (NH == 2) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) acc += dc_part[(h * 8 + cgc * 4 + qq) * 32 + k];
      } else {
#pragma unroll
        for (int w8 = 0; w8 < 2; ++w8)
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) acc += dc_part[(w8 * 8 + cgc * 4 + qq) * 32 + k];
      }
      if (i < D) part_dc[(size_t)cta * D + i] = acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, kCols);
}


// ================================================================================================================
// backward pass 2 (summary path):  a = softmax_pixels(y),  dy = a (t - delta),  t = x^T ds^T
//   d_x (+)= dy K + a ds ,   d_K_2 += dy^T x
//   step = (tile, 64-query half).  TMEM per step buffer i: y -> dy = [128 i, +64)   t -> a = [128 i + 64, +64)
//                                  dx[0..2] = [256 + 32 j, +32)    dK[h] = [352 + 32 h, +32)  rows = queries of half h
// ================================================================================================================
template <int QP>
struct BwdSumSmem {
  static constexpr int NH = QP / 64;
  static constexpr int NST = QP > 64 ? 2 : 3;
  static constexpr size_t x_ring = 0, x_lo = NST * kXTile, x_k = x_lo + kXTile /* [2] */, k_hi = x_k + 2 * kXTile,
                          k_lo = k_hi + QP * 128,
                          ds = k_lo + QP * 128,                    // [QP rows][32 e]                 d_summary (K-major)
                          kT = ds + QP * 128,                      // [QP/32 atoms][32 e rows][32 q]  queries transposed
                          dsT = kT + (QP / 32) * 32 * 128,         // same shape                      d_summary transposed
                          dyT = dsT + (QP / 32) * 32 * 128,        // [4 px atoms][128 rows][32 px]   dy^T, rows >= 64 zero
                          tail = dyT + 4 * 128 * 128;
  static constexpr size_t bytes = 1024 + tail + 2 * QP * 4 + 128;
  static_assert(bytes <= 227 * 1024, "shared-memory budget");
};

template <int QP>
__global__ void __launch_bounds__(kThreads, 1) sql_ws_bwd_sum_kernel(
    const __grid_constant__ CUtensorMap map_mn, const __grid_constant__ CUtensorMap map_k,
    const float* __restrict__ queries, const float* __restrict__ summary, const float* __restrict__ row_max,
    const float* __restrict__ row_sum, const float* __restrict__ d_summary, int Q, int n, int tiles_per_chunk,
    int accumulate, float* __restrict__ d_x, float* __restrict__ part_dK /*[cta][Q][32]*/) {
  using L = BwdSumSmem<QP>;
  constexpr int NH = L::NH, NST = L::NST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* x_ring = base + L::x_ring;
  uint8_t* x_lo = base + L::x_lo;
  uint8_t* x_k = base + L::x_k;
  uint8_t* k_hi = base + L::k_hi;
  uint8_t* k_lo = base + L::k_lo;
  uint8_t* dsm = base + L::ds;
  uint8_t* kT = base + L::kT;
  uint8_t* dsT = base + L::dsT;
  uint8_t* dyT = base + L::dyT;
  float* cqv = reinterpret_cast<float*>(base + L::tail);     // a[p,q] = 2^(y log2e + cq),  cq = -m log2e - log2(l)
  float* dlv = cqv + QP;                                      // delta_q = ds_q . summary_q
  uint64_t* bars = reinterpret_cast<uint64_t*>(dlv + QP);
  uint64_t* bar_full = bars;            // [NST <= 3] TMA: x tile landed in its ring slot
  uint64_t* bar_split = bars + 3;       // 8 warps: x_lo written
  uint64_t* bar_xk = bars + 4;          // [2] TMA: K-major x tile landed in x_k[i]
  uint64_t* bar_z = bars + 6;           // [2] MMA commit: y, t of a step in buffer i
  uint64_t* bar_epi = bars + 8;         // [2] 8 warps: dy, a of a step in buffer i (TMEM) and dyT (shared)
  uint64_t* bar_m2 = bars + 10;         // [2] MMA commit: d_x / d_K MMAs of a step done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int b = blockIdx.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 512;
  constexpr uint32_t tm_dx = 256, tm_dk = 352;
  const int t_begin = blockIdx.x * tiles_per_chunk;
  const int t_end = min((n + kTile - 1) / kTile, t_begin + tiles_per_chunk);
  const int ntiles = max(t_end - t_begin, 0);
  const int nsteps = ntiles * NH;
  if (threadIdx.x == kEpiThreads) {
    tma_prefetch_desc(&map_mn);
    tma_prefetch_desc(&map_k);
    for (int i = 0; i < NST; ++i) mbar_init(bar_full + i, 1);
    mbar_init(bar_split, kGrpWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_xk + i, 1); mbar_init(bar_z + i, 1); mbar_init(bar_epi + i, kGrpWarps); mbar_init(bar_m2 + i, 1); }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {
    tmem_alloc(tmem_slot, kCols);
    tmem_relinquish();
  }
  __syncthreads();
  if (threadIdx.x == kEpiThreads && ntiles > 0) {
    for (int i = 0; i < NST && i < ntiles; ++i)
      tma_x_tile(x_ring + i * kXTile, kXBlock, &map_mn, (t_begin + i) * kTile, b * kE, bar_full + i);
    tma_x_tile(x_k, kXBlock, &map_k, t_begin * kTile, b * kE, bar_xk);
  }
  if (warp < kEpiWarps) {
    const float* qb = queries + (size_t)b * Q * kE;
    const float* dsb = d_summary + (size_t)b * Q * kE;
    stage_kmajor(k_hi, k_lo, qb, Q, QP, threadIdx.x, kEpiThreads);
    stage_kmajor(dsm, nullptr, dsb, Q, QP, threadIdx.x, kEpiThreads);
    stage_transposed(kT, qb, Q, QP, threadIdx.x, kEpiThreads);
    stage_transposed(dsT, dsb, Q, QP, threadIdx.x, kEpiThreads);
    for (int i = threadIdx.x; i < 4 * 128 * 32 / 4; i += kEpiThreads)
      reinterpret_cast<float4*>(dyT)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = threadIdx.x; q < QP; q += kEpiThreads) {
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
      cqv[q] = q < Q ? fmaf(-m, kLog2e, log2f(inv)) : -INFINITY;    // (-inf for the padded queries: a = 0)
      dlv[q] = delta;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  float* dxb = d_x + (size_t)b * kE * n;

  if (warp == kEpiWarps) {
    // ---------------------------------------------------------------- control lane
    if (elect_one()) {
      const uint32_t id_y = make_idesc_tf32(kTile, 64, 1, 0);       // y, t: A = x (MN-major), B = K / ds rows (K-major)
      const uint32_t id_32 = make_idesc_tf32(128, 32, 0, 0);        // d_x
      const uint32_t id_dk = make_idesc_tf32(64, 32, 0, 0);         // d_K: M = 64, the 64 query rows of the half
      const uint64_t dx_ring = make_desc_mn32(smem_u32(x_ring), kXBlock), dx_lo = make_desc_mn32(smem_u32(x_lo), kXBlock);
      const uint64_t dk_hi = make_desc_sw128(smem_u32(k_hi), 16, 1024), dk_lo = make_desc_sw128(smem_u32(k_lo), 16, 1024);
      const uint64_t d_ds = make_desc_sw128(smem_u32(dsm), 16, 1024), d_kT = make_desc_sw128(smem_u32(kT), 16, 1024);
      const uint64_t d_dsT = make_desc_sw128(smem_u32(dsT), 16, 1024), d_dyT = make_desc_sw128(smem_u32(dyT), 16, 1024);
      const uint64_t d_xk = make_desc_sw128(smem_u32(x_k), 16, 1024);
      auto issue_mma2 = [&](int sp, int ti, int h) {
        mbar_wait(bar_epi + (sp & 1), (sp >> 1) & 1);
        if (h == 0) mbar_wait(bar_xk + (ti & 1), (ti >> 1) & 1);
        tc_fence_after();
        const uint32_t t_dx = tmem + tm_dx + (ti % 3) * 32, t_dy = tmem + (sp & 1) * 128, t_a = t_dy + 64;
        const uint32_t t_dk = tmem + tm_dk + h * 32;
        const uint64_t kt = desc_add(d_kT, h * 2 * 32 * 128), dst = desc_add(d_dsT, h * 2 * 32 * 128);
#pragma unroll
        for (int k = 0; k < kTile / 8; ++k)       // d_K_h += dy_h^T x            (K = 128 pixels)
          umma_tf32_ss(t_dk, desc_add(d_dyT, (k >> 2) * 128 * 128 + (k & 3) * 32),
                       desc_add(d_xk, (ti & 1) * kXTile + (k >> 2) * kXBlock + (k & 3) * 32), id_dk, (ti > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 64 / 8; ++k)          // d_x tile (+)= dy_h K_h        (K = 64 queries, A from TMEM)
          umma_tf32_ts(t_dx, t_dy + k * 8, desc_add(kt, (k >> 2) * 32 * 128 + (k & 3) * 32), id_32, (h > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 64 / 8; ++k)          //              + a_h ds_h
          umma_tf32_ts(t_dx, t_a + k * 8, desc_add(dst, (k >> 2) * 32 * 128 + (k & 3) * 32), id_32, 1u);
        umma_commit(bar_m2 + (sp & 1));
      };
      int ti = 0, h = 0, slot = 0, pti = 0, ph = 0;
      for (int s = 0; s < nsteps; ++s) {
        if (h == 0) mbar_wait(bar_split, ti & 1);
        tc_fence_after();
        const uint64_t xa = desc_add(dx_ring, slot * kXTile);
        mma_x_b3(xa, dx_lo, desc_add(dk_hi, h * 64 * 128), desc_add(dk_lo, h * 64 * 128), tmem + (s & 1) * 128, id_y);
#pragma unroll
        for (int k = 0; k < kE / 8; ++k)          // t_h = x^T ds_h^T (single pass: it only enters the gradient)
          umma_tf32_ss(tmem + (s & 1) * 128 + 64, desc_add(xa, k * 1024), desc_add(d_ds, h * 64 * 128 + k * 32), id_y, k > 0);
        umma_commit(bar_z + (s & 1));
        if (s > 0) issue_mma2(s - 1, pti, ph);
        pti = ti; ph = h;
        if (++h == NH) { h = 0; ++ti; slot = slot + 1 == NST ? 0 : slot + 1; }
      }
      if (nsteps > 0) issue_mma2(nsteps - 1, pti, ph);
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps: group `grp` owns steps s = grp (mod 2)
    const int grp = warp >> 3, wl = warp & 7, q = wl & 3, cg = wl >> 2, gtid = threadIdx.x & (kGrpThreads - 1);
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t lane_hi = (uint32_t)lane >> 2, lane_lo4 = ((uint32_t)lane & 3u) << 2;
    auto store_dx = [&](int ti) {          // this thread's 16 channels of its pixel of tile ti: TMEM (+ old value) -> global
      const int pp = (t_begin + ti) * kTile + q * 32 + lane;
      float old[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) old[i] = (accumulate && pp < n) ? __ldcg(dxb + (size_t)(cg * 16 + i) * n + pp) : 0.f;
      float v[16];
      tmem_ld16(lane_base + tm_dx + (ti % 3) * 32 + cg * 16, v);
      tmem_wait_ld();
      if (pp < n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dxb[(size_t)(cg * 16 + i) * n + pp] = v[i] + old[i];
      }
    };
    auto split_tile = [&](int i) {
      mbar_wait(bar_full + i % NST, (i / NST) & 1);
      split_x<kGrpThreads>(x_ring + (i % NST) * kXTile, x_lo, gtid);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split);
    };
    if (grp == 0 && ntiles > 0) split_tile(0);
    for (int s = grp; s < nsteps; s += 2) {
      const int ti = s / NH, h = s - ti * NH;
      const int p = (t_begin + ti) * kTile + q * 32 + lane;
      const bool pin = p < n;
      mbar_wait(bar_z + grp, (s >> 1) & 1);                // y, t of step s in buffer grp
      tc_fence_after();
      if (h == NH - 1) {                                   // the tile's last y / t MMAs are done: its ring slot and x_lo are free
        if (gtid == 0 && ti + NST < ntiles)
          tma_x_tile(x_ring + (ti % NST) * kXTile, kXBlock, &map_mn, (t_begin + ti + NST) * kTile, b * kE, bar_full + ti % NST);
        if (ti + 1 < ntiles) split_tile(ti + 1);
      }
      // ---- this thread's 32 queries of its pixel: a = softmax_pixels(y), dy = a (t - delta); both back to TMEM
      float dy[32];
      {
        const int c0 = h * 64 + cg * 32;
        const uint32_t ty = lane_base + grp * 128 + cg * 32, tt = ty + 64;
#pragma unroll
        for (int c = 0; c < 32; c += 16) {
          float yv[16], tv[16];
          tmem_ld16(ty + c, yv);
          tmem_ld16(tt + c, tv);
          tmem_wait_ld();
#pragma unroll
          for (int i4 = 0; i4 < 16; i4 += 4) {
            const float4 cq4 = *reinterpret_cast<const float4*>(cqv + c0 + c + i4);
            const float4 dl4 = *reinterpret_cast<const float4*>(dlv + c0 + c + i4);
            const float cq[4] = {cq4.x, cq4.y, cq4.z, cq4.w}, dl[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = i4 + j;
              const float a = pin ? ex2f(fmaf(yv[i], kLog2e, cq[j])) : 0.f;
              dy[c + i] = a * (tv[i] - dl[j]);
              tv[i] = a;
              yv[i] = dy[c + i];
            }
          }
          tmem_st16(ty + c, yv);
          tmem_st16(tt + c, tv);
        }
      }
      // ---- dy^T (A operand of the d_K MMA) goes to the single dyT tile once the previous step's MMAs have read it
      if (s > 0) {
        mbar_wait(bar_m2 + ((s - 1) & 1), ((s - 1) >> 1) & 1);
        tc_fence_after();
        if (h == 0 && gtid == 0 && ti + 1 < ntiles)        // all MMAs of tile ti-1 are done: x_k[(ti+1)&1] is free, refill it
          tma_x_tile(x_k + (size_t)((ti + 1) & 1) * kXTile, kXBlock, &map_k, (t_begin + ti + 1) * kTile, b * kE,
                     bar_xk + ((ti + 1) & 1));
      } else if (gtid == 0 && ntiles > 1) {
        tma_x_tile(x_k + kXTile, kXBlock, &map_k, (t_begin + 1) * kTile, b * kE, bar_xk + 1);
      }
      {
        uint8_t* dyw = dyT + q * 128 * 128;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int row = cg * 32 + i;
          *reinterpret_cast<float*>(dyw + row * 128 + ((lane_hi ^ (uint32_t)(row & 7)) << 4) + lane_lo4) = dy[i];
        }
      }
      tmem_wait_st();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_epi + grp);
      if (s > 0 && h == 0) store_dx(ti - 1);
    }
    if (nsteps > 0 && grp == ((nsteps - 1) & 1)) {
      mbar_wait(bar_m2 + ((nsteps - 1) & 1), ((nsteps - 1) >> 1) & 1);
      tc_fence_after();
      store_dx(ntiles - 1);
    }
    named_sync(1, kEpiThreads);
    tc_fence_after();
    // ---- d_K accumulators: rows = queries of half h; an M = 64 accumulator keeps row m in TMEM lane 32 (m / 16) + m % 16
    // (lanes 0..15 of every lane quarter); group h reads half h
    {
      const int h = grp;
      if (h < NH) {
        const int qq = lane < 16 ? h * 64 + q * 16 + lane : Q;
        float a[16];
        if (ntiles > 0) {
          tmem_ld16(lane_base + tm_dk + h * 32 + cg * 16, a);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = 0.f;
        }
        if (qq < Q) {
#pragma unroll
          for (int i = 0; i < 16; ++i) part_dK[((size_t)cta * Q + qq) * kE + cg * 16 + i] = a[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, kCols);
}

// ================================================================================================================
// forward summaries (FullQueryLayer, networks/layers.py:17-20):  y = K x,  a = softmax over PIXELS,  summary = a x^T
//
// Here lane = query row and columns = pixels: the softmax runs ALONG a thread's own columns, so the running maximum and
// denominator live in registers and need no cross-lane traffic; P = exp(y - m) goes back to TMEM as the A operand of the
// second contraction, and the queries themselves (hi / lo, static for the CTA) sit in TMEM as the A operand of the first.
// Both contractions are therefore A-in-TMEM instructions whose only shared-memory traffic is the 1 KB x operand: measured
// (tools/mma_chain_probe.cu) 16 cycles per M = 128, N = 32, K = 8 instruction against 40 with A in shared memory (the
// 4 KB A read at 128 B / clk is what an SS-mode instruction of small N costs).
//
//   step        one 32-pixel tile.  FOUR groups of four epilogue warps take the steps round-robin; the control lane runs
//               three steps ahead:  ... MMA1(s) | MMA2(s-3) | MMA1(s+1) | MMA2(s-2) ...  (tcgen05.mma executes in issue
//               order, so MMA1(s+4) may overwrite the y / P buffer of group s % 4 right behind MMA2(s))
//   online      every (group, lane) keeps its own running (m, l) and its own summary accumulator row in TMEM: the second
//   softmax     contraction ACCUMULATES into it across the steps of the group, and the row is rescaled in place only when
//               the maximum moves by more than 8 (rare after the first tiles); the partial states of a query are merged
//               through shared memory at the end and leave as one record per (chunk, query)
//   Q <= 64     the 128 TMEM lanes hold the 64 queries TWICE: lanes 0..63 own pixels 0..15 of the tile, lanes 64..127
//               pixels 16..31.  Each lane half stores P_hi in its own 16 columns and P_lo in the 16 columns the other half
//               owns (what y left there is not needed), and the second contraction runs once per half into separate
//               accumulators whose other 64 rows are ignored
//   TMEM        group g at columns 96 g:  [0, 32) y -> P_hi,  [32, 64) P_lo,  [64, 96) S                    (Q > 64)
//                                         [0, 32) y -> P_hi | P_lo crossed,  [32, 64) S_A,  [64, 96) S_B    (Q <= 64)
//               [384, 416) K_hi,  [416, 448) K_lo
// ================================================================================================================
constexpr int kSumSlot = 4 * kXBlock;      // ring slot of a step: x MN-major hi | lo | x K-major hi | lo   (4 KB each)
constexpr uint32_t kSumGrpCols = 96, kSumKCol = 384;
constexpr int kSumSplitWarps = 3;          // warps 17..19: each writes the lo images of every third landed step (keeps the
                                           // split and its fence.proxy.async off the epilogue groups' step cycle; one warp per
                                           // step so that three steps' fences overlap).  20 warps: five per scheduler, 96
                                           // registers per thread
constexpr int kSumThreads = kThreads + 32 * kSumSplitWarps;

template <int QP>
struct SumSmem {
#ifndef SQLX_SUM_NSLOT
#define SQLX_SUM_NSLOT 12
#endif
  static constexpr int NSLOT = SQLX_SUM_NSLOT;
  static constexpr int NPART = QP == 64 ? 8 : 4;       // partial softmax states per query at the end
  static constexpr int REC = kE + 3;                   // m, l, 32 sums (odd stride: conflict-free)
  static constexpr size_t ring = 0, tail = (size_t)NSLOT * kSumSlot;
  static constexpr size_t bytes = 1024 + tail + 512;
  static_assert(bytes <= 227 * 1024, "shared-memory budget");
  static_assert((size_t)NPART * QP * REC * 4 <= (size_t)NSLOT * kSumSlot, "merge buffer reuses the ring");
};

template <int QP>
__global__ void __launch_bounds__(kSumThreads, 1) sql_ws_summary_kernel(const __grid_constant__ CUtensorMap map_mn,
                                                                     const __grid_constant__ CUtensorMap map_k,
                                                                     const float* __restrict__ queries, int Q, int n,
                                                                     int steps_per_chunk,
                                                                     float* __restrict__ partial /*[B][chunks][Q][34]*/) {
  using L = SumSmem<QP>;
  constexpr bool kPair = QP == 64;
  constexpr int CW = kPair ? 16 : 32;                  // pixels per thread and step
  constexpr int NSLOT = L::NSLOT;
  constexpr uint32_t kSOff = kPair ? 32 : 64;          // accumulator columns inside the group's block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = base + L::ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + L::tail);
  uint64_t* bar_full = bars;                    // [NSLOT] TMA: both images of a step landed
  uint64_t* bar_split = bars + NSLOT;           // [NSLOT] split warps: lo images of the step in the slot written
  uint64_t* bar_z = bars + 2 * NSLOT;           // [4] MMA commit: y of the group's step (and every earlier MMA) done
  uint64_t* bar_epi = bars + 2 * NSLOT + 4;     // [4] 4 warps: P_hi / P_lo of the group's step in TMEM
  uint64_t* bar_done = bars + 2 * NSLOT + 8;    // MMA commit: everything done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSLOT + 9);
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kCols = 512;
  const int s_begin = chunk * steps_per_chunk;
  const int nsteps = max(min((n + 31) / 32, s_begin + steps_per_chunk) - s_begin, 0);
  auto tma_step = [&](int i) {                  // both images of step i into its ring slot
    uint8_t* slot = ring + (size_t)(i % NSLOT) * kSumSlot;
    mbar_arrive_expect_tx(bar_full + i % NSLOT, 2 * kXBlock);
    tma_load_2d(slot, &map_mn, (s_begin + i) * 32, b * kE, bar_full + i % NSLOT);
    tma_load_2d(slot + 2 * kXBlock, &map_k, (s_begin + i) * 32, b * kE, bar_full + i % NSLOT);
  };
  if (threadIdx.x == kEpiThreads) {
    tma_prefetch_desc(&map_mn);
    tma_prefetch_desc(&map_k);
    for (int i = 0; i < NSLOT; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_split + i, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_z + i, 1); mbar_init(bar_epi + i, 4); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {
    tmem_alloc(tmem_slot, kCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == kEpiThreads)
    for (int i = 0; i < NSLOT && i < nsteps; ++i) tma_step(i);
  if (warp < 4) {
    // queries -> TMEM rows (lane = query; Q <= 64: lanes 64.. repeat lanes 0..; rows past Q are zero), hi / lo
    const int r = warp * 32 + lane, qi = kPair ? (r & 63) : r;
    const float4* qrow = reinterpret_cast<const float4*>(queries + ((size_t)b * Q + (qi < Q ? qi : 0)) * kE);
    float hi[kE], lo[kE];
#pragma unroll
    for (int j = 0; j < kE / 4; ++j) {
      const float4 v = qi < Q ? __ldg(qrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      hi[4 * j] = tf32_hi(v.x); hi[4 * j + 1] = tf32_hi(v.y); hi[4 * j + 2] = tf32_hi(v.z); hi[4 * j + 3] = tf32_hi(v.w);
      lo[4 * j] = v.x - hi[4 * j]; lo[4 * j + 1] = v.y - hi[4 * j + 1]; lo[4 * j + 2] = v.z - hi[4 * j + 2];
      lo[4 * j + 3] = v.w - hi[4 * j + 3];
    }
    const uint32_t tk = tmem + ((uint32_t)(warp * 32) << 16) + kSumKCol;
#pragma unroll
    for (int c = 0; c < kE; c += 16) {
      tmem_st16(tk + c, *reinterpret_cast<float(*)[16]>(hi + c));
      tmem_st16(tk + kE + c, *reinterpret_cast<float(*)[16]>(lo + c));
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == kEpiWarps) {
    // ---------------------------------------------------------------- control lane
    if (elect_one()) {
      const uint32_t id_y = make_idesc_tf32(128, 32, 0, 1);      // y: A = K rows (TMEM), B = x (MN-major, one 32-px atom)
      const uint32_t id_s = make_idesc_tf32(128, 32, 0, 0);      // S: A = P (TMEM),  B = x rows = channels, K = pixels
      const uint32_t tk_hi = tmem + kSumKCol, tk_lo = tk_hi + kE;
      const uint64_t d_mn = make_desc_mn32(smem_u32(ring), kXBlock);
      const uint64_t d_xk = make_desc_sw128(smem_u32(ring) + 2 * kXBlock, 16, 1024);
      // The issuing thread must never let the tensor pipe drain: the tcgen05.mma queue is only a few instructions deep
      // (issue time = execution time, tools/mma_chain_probe.cu) and even a SATISFIED mbarrier wait between two batches
      // costs ~70 idle cycles (tools/mma_mix_probe.cu: 16 -> 22 cycles per instruction with one wait per 12).  So the
      // barrier of the NEXT batch is probed (mbarrier.test_wait, non-blocking) BEFORE the current batch is issued, while
      // the queue is still full, and the blocking wait runs only when that probe failed.
      auto issue_mma1 = [&](int i, uint32_t so) {
        const uint32_t ty = tmem + (i & 3) * kSumGrpCols;
        const uint64_t xh = desc_add(d_mn, so), xl = desc_add(d_mn, so + kXBlock);
#pragma unroll
        for (int k = 0; k < kE / 8; ++k) umma_tf32_ts(ty, tk_hi + k * 8, desc_add(xh, k * 1024), id_y, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kE / 8; ++k) umma_tf32_ts(ty, tk_lo + k * 8, desc_add(xh, k * 1024), id_y, 1u);
#pragma unroll
        for (int k = 0; k < kE / 8; ++k) umma_tf32_ts(ty, tk_hi + k * 8, desc_add(xl, k * 1024), id_y, 1u);
        umma_commit(bar_z + (i & 3));
      };
      auto issue_mma2 = [&](int j, uint32_t so, uint32_t first) {   // first = 0: the group's first step initialises its accumulator
        const uint32_t tb = tmem + (j & 3) * kSumGrpCols;
        const uint64_t xh = desc_add(d_xk, so), xl = desc_add(d_xk, so + kXBlock);
        if (!kPair) {
          umma_tf32_ts(tb + kSOff, tb, xh, id_s, first);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_tf32_ts(tb + kSOff, tb + k * 8, desc_add(xh, k * 32), id_s, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ts(tb + kSOff, tb + 32 + k * 8, desc_add(xh, k * 32), id_s, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32_ts(tb + kSOff, tb + k * 8, desc_add(xl, k * 32), id_s, 1u);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {                          // lane half h: pixels [16 h, 16 h + 16)
            const uint32_t own = tb + 16 * h, oth = tb + 16 * (1 - h), td = tb + kSOff + 32 * h;
            umma_tf32_ts(td, own, desc_add(xh, (2 * h) * 32), id_s, first);
            umma_tf32_ts(td, own + 8, desc_add(xh, (2 * h + 1) * 32), id_s, 1u);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              umma_tf32_ts(td, oth + k * 8, desc_add(xh, (2 * h + k) * 32), id_s, 1u);
              umma_tf32_ts(td, own + k * 8, desc_add(xl, (2 * h + k) * 32), id_s, 1u);
            }
          }
        }
      };
      // The loop is unrolled by the ring depth (a multiple of the four groups): slot offsets, TMEM columns and barrier
      // addresses of every step of a round are compile-time constants, so between two tcgen05.mma there is nothing but
      // one uniform 64-bit add per descriptor (the rolled loop spent ~380 of its 765 cycles per step in R2UR moves and
      // slot arithmetic, with the tensor pipe idle: the queue drains while the issuing thread computes).  Measured
      // 46 -> 40 us at config 2.  (Compile-time TMEM / shared-memory bases were tried on top: SLOWER, 47 us -- the inline
      // PTX operands are "r" registers, so every constant is materialised in a vector register and moved with R2UR.)
      static_assert(NSLOT % 4 == 0, "ring depth must be a multiple of the number of epilogue groups");
      if (nsteps > 0) mbar_wait(bar_split, 0);
      for (int base_i = 0; base_i < nsteps + 3; base_i += NSLOT) {
        const uint32_t ph = (uint32_t)(base_i / NSLOT) & 1u;          // phase parity of the ring barriers in this round
#pragma unroll
        for (int u = 0; u < NSLOT; ++u) {
          const int i = base_i + u, j = i - 3;
          if (i >= nsteps + 3) break;
          const bool has1 = i < nsteps, has2 = j >= 0, hasn = i + 1 < nsteps;
          const int uj = (u + NSLOT - 3) % NSLOT, un = (u + 1) % NSLOT;      // slots of step j and step i + 1
          const uint32_t ph_n = un == 0 ? ph ^ 1u : ph;
          // probe what the next two batches need while the queue still holds the previous batch
          const bool r_epi = has2 && mbar_test(bar_epi + (uj & 3), (uint32_t)(j >> 2) & 1u);
          if (has1) {
            tc_fence_after();
            issue_mma1(u, (uint32_t)u * kSumSlot);
          }
          const bool r_split = hasn && mbar_test(bar_split + un, ph_n);
          if (has2) {
            if (!r_epi) mbar_wait(bar_epi + (uj & 3), (uint32_t)(j >> 2) & 1u);
            tc_fence_after();
            if (j < 4) issue_mma2(uj, (uint32_t)uj * kSumSlot, 0u);
            else issue_mma2(uj, (uint32_t)uj * kSumSlot, 1u);
          }
          if (hasn && !r_split) mbar_wait(bar_split + un, ph_n);
        }
      }
      umma_commit(bar_done);
    }
  } else if (warp > kEpiWarps) {
    // ---------------------------------------------------------------- split warps: lo images of the steps as they land
    // (the landing buffers themselves are the hi operands)
    constexpr int PER = 2 * kXBlock / 16 / 32;                               // float4 per thread and step
    for (int i = warp - (kEpiWarps + 1); i < nsteps; i += kSumSplitWarps) {
      uint8_t* slot = ring + (size_t)(i % NSLOT) * kSumSlot;
      mbar_wait(bar_full + i % NSLOT, (i / NSLOT) & 1);
#pragma unroll
      for (int j0 = 0; j0 < PER; j0 += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = lane + (j0 + j) * 32;                              // [0, 256): MN-major image, [256, 512): K-major
          v[j] = *reinterpret_cast<const float4*>(slot + (idx >> 8) * 2 * kXBlock + (idx & 255) * 16);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = lane + (j0 + j) * 32;
          v[j].x -= tf32_hi(v[j].x); v[j].y -= tf32_hi(v[j].y); v[j].z -= tf32_hi(v[j].z); v[j].w -= tf32_hi(v[j].w);
          *reinterpret_cast<float4*>(slot + kXBlock + (idx >> 8) * 2 * kXBlock + (idx & 255) * 16) = v[j];
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_split + i % NSLOT);
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps: group `grp` owns steps s = grp (mod 4)
    const int grp = warp >> 2, q = warp & 3, gtid = threadIdx.x & 127;
    const int half = kPair ? (q >> 1) : 0;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + grp * kSumGrpCols;
    const uint32_t t_own = lane_base + (kPair ? 16 * half : 0), t_lo = lane_base + (kPair ? 16 * (1 - half) : 32);
    const uint32_t t_S = lane_base + kSOff + 32 * half;
    float m = -INFINITY, l = 0.f;
    for (int s = grp; s < nsteps; s += 4) {
      mbar_wait(bar_z + grp, (s >> 2) & 1);        // y of step s is in the group's buffer; every MMA of steps <= s-4 is done
      tc_fence_after();
      if (gtid == 0 && s >= 4 && s - 4 + NSLOT < nsteps) tma_step(s - 4 + NSLOT);     // the slot of step s-4 is free
      float v[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 16) tmem_ld16(t_own + c, *reinterpret_cast<float(*)[16]>(v + c));
      tmem_wait_ld();
      // pixels >= n were zero-filled by TMA: exclude them
      const int valid = min(32, n - (s_begin + s) * 32) - (kPair ? 16 * half : 0);
      const bool full = valid >= CW;
      float tmax = -INFINITY;
      if (full) {
        float t4[CW / 4];
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) t4[j] = fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3]));
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) tmax = fmaxf(tmax, t4[j]);
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if (i < valid) tmax = fmaxf(tmax, v[i]);
      }
      // lazy reference point: it moves only when exceeded by more than 8 (exp arguments stay <= 8)
      const bool need = tmax > m + 8.f;
      if (__any_sync(0xffffffffu, need)) {
        const float rs = need ? __expf(m - tmax) : 1.f;      // m = -inf at first: 0
        if (need) m = tmax;
        l *= rs;
        if (s >= 4) {                                        // rescale this thread's accumulator row in place
#pragma unroll
          for (int c = 0; c < kE; c += 16) {
            float a[16];
            tmem_ld16(t_S + c, a);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] *= rs;
            tmem_st16(t_S + c, a);
          }
        }
      }
      const float nm2 = -m * kLog2e;                         // exp(y - m) = 2^(y log2e - m log2e)
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < CW; c += 16) {
        float lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pe = ex2f(fmaf(v[c + i], kLog2e, nm2));
          if (!full && c + i >= valid) pe = 0.f;
          ls[i & 3] += pe;
          const float h = tf32_hi(pe);
          v[c + i] = h;
          lo[i] = pe - h;
        }
        tmem_st16(t_own + c, *reinterpret_cast<float(*)[16]>(v + c));
        tmem_st16(t_lo + c, lo);
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_epi + grp);
    }
    // ---- merge the partial states of every query (4 groups x lane halves) and write one record per (chunk, query)
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* mb = reinterpret_cast<float*>(ring);
    {
      const int r = q * 32 + lane, qi = kPair ? (r & 63) : r, part = kPair ? grp * 2 + half : grp;
      float* rec = mb + ((size_t)part * QP + qi) * L::REC;
      const bool live = grp < nsteps;
      rec[0] = m; rec[1] = l;
#pragma unroll
      for (int c = 0; c < kE; c += 16) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = 0.f;
        if (live) {
          tmem_ld16(t_S + c, a);
          tmem_wait_ld();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) rec[2 + c + i] = a[i];
      }
    }
    named_sync(1, kEpiThreads);
    {
      constexpr int NCH = kEpiThreads / QP, CPT = kE / NCH;      // channel blocks per query, channels per thread
      const int qi = threadIdx.x % QP, cb = threadIdx.x / QP;
      if (qi < Q) {
        float M = -INFINITY;
#pragma unroll
        for (int pt = 0; pt < L::NPART; ++pt) M = fmaxf(M, mb[((size_t)pt * QP + qi) * L::REC]);
        float Ls = 0.f, acc[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) acc[i] = 0.f;
#pragma unroll
        for (int pt = 0; pt < L::NPART; ++pt) {
          const float* rec = mb + ((size_t)pt * QP + qi) * L::REC;
          const float w = rec[0] == -INFINITY ? 0.f : __expf(rec[0] - M);
          Ls = fmaf(rec[1], w, Ls);
#pragma unroll
          for (int i = 0; i < CPT; ++i) acc[i] = fmaf(rec[2 + cb * CPT + i], w, acc[i]);
        }
        float* out = partial + (((size_t)b * chunks + chunk) * Q + qi) * (kE + 2);
        if (cb == 0) { out[0] = M; out[1] = Ls; }
#pragma unroll
        for (int i = 0; i < CPT; ++i) out[2 + cb * CPT + i] = acc[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, kCols);
}

}  // namespace wsql

// ------------------------------------------------------------------------------------------------
// launchers (called from the C ABI in sql_fp32.cu)
// ------------------------------------------------------------------------------------------------
// SMs the one-CTA-per-SM kernels may occupy (sqlx_sql_set_sm_budget): a caller that runs a communication kernel beside the
// summary-path backward (the in-step gradient all-reduce) leaves it room instead of making it queue for SMs
std::atomic<int> g_sm_budget{kNumSMs};

void ws_plan(int B, int n, int* chunks, int* tiles_per_chunk) {
  const int tiles = ceil_div(n, wsql::kTile);
  int c = g_sm_budget.load() / B;   // one CTA per SM
  c = c < 1 ? 1 : (c > tiles ? tiles : c);
  *tiles_per_chunk = ceil_div(tiles, c);
  *chunks = ceil_div(tiles, *tiles_per_chunk);
}

template <int DP>
static int launch_ws_pred(const CUtensorMap& map, const float* Mx, const float* bp, const float* centers, int B, int D, int n,
                          float* pred, float* stat_m, float* stat_inv, cudaStream_t st) {
  int chunks, tpc;
  ws_plan(B, n, &chunks, &tpc);
  if (int e = ensure_dyn_smem(wsql::sql_ws_pred_kernel<DP>, wsql::PredSmem<DP>::bytes)) return e;
  ProfScope prof("sql_tc_pred_kernel", st);
  wsql::sql_ws_pred_kernel<DP><<<dim3(chunks, B), wsql::kThreads, wsql::PredSmem<DP>::bytes, st>>>(
      map, Mx, bp, centers, D, n, tpc, pred, stat_m, stat_inv);
  return check_launch("sql_ws_pred_kernel");
}

int ws_pred_fwd(const float* x, const float* Mx, const float* bp, const float* centers, int B, int D, int n, float* pred,
                float* stat_m, float* stat_inv, cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  CUtensorMap map;
  if (int e = make_tensor_map_2d(&map, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (D <= 64) return launch_ws_pred<64>(map, Mx, bp, centers, B, D, n, pred, stat_m, stat_inv, st);
  return launch_ws_pred<128>(map, Mx, bp, centers, B, D, n, pred, stat_m, stat_inv, st);
}

template <int DP>
static int launch_ws_bwd_pred(const CUtensorMap& map_mn, const CUtensorMap& map_k, const float* Mx, const float* bp,
                              const float* centers, const float* g_pred, const float* pred, const float* stat_m,
                              const float* stat_inv, int B, int D, int n, int chunks, int tpc, float* d_x, float* part_dM,
                              float* part_db, float* part_dc, cudaStream_t st) {
  if (int e = ensure_dyn_smem(wsql::sql_ws_bwd_pred_kernel<DP>, wsql::BwdPredSmem<DP>::bytes)) return e;
  ProfScope prof("sql_tc_bwd_pred_kernel", st);
  wsql::sql_ws_bwd_pred_kernel<DP><<<dim3(chunks, B), wsql::kThreads, wsql::BwdPredSmem<DP>::bytes, st>>>(
      map_mn, map_k, Mx, bp, centers, g_pred, pred, stat_m, stat_inv, D, n, tpc, d_x, part_dM, part_db, part_dc);
  return check_launch("sql_ws_bwd_pred_kernel");
}

int ws_bwd_pred(const float* x, const float* Mx, const float* bp, const float* centers, const float* g_pred,
                const float* pred, const float* stat_m, const float* stat_inv, int B, int D, int n, float* d_x,
                float* part_dM, float* part_db, float* part_dc, int chunks, int tpc, cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  CUtensorMap map_mn, map_k;
  if (int e = make_tensor_map_2d(&map_mn, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = make_tensor_map_2d(&map_k, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 0)) return e;
  if (D <= 64)
    return launch_ws_bwd_pred<64>(map_mn, map_k, Mx, bp, centers, g_pred, pred, stat_m, stat_inv, B, D, n, chunks, tpc, d_x,
                                  part_dM, part_db, part_dc, st);
  return launch_ws_bwd_pred<128>(map_mn, map_k, Mx, bp, centers, g_pred, pred, stat_m, stat_inv, B, D, n, chunks, tpc, d_x,
                                 part_dM, part_db, part_dc, st);
}

template <int QP>
static int launch_ws_bwd_sum(const CUtensorMap& map_mn, const CUtensorMap& map_k, const float* queries, const float* summary,
                             const float* row_max, const float* row_sum, const float* d_summary, int B, int Q, int n,
                             int chunks, int tpc, int accumulate, float* d_x, float* part_dK, cudaStream_t st) {
  if (int e = ensure_dyn_smem(wsql::sql_ws_bwd_sum_kernel<QP>, wsql::BwdSumSmem<QP>::bytes)) return e;
  ProfScope prof("sql_tc_bwd_sum_kernel", st);
  wsql::sql_ws_bwd_sum_kernel<QP><<<dim3(chunks, B), wsql::kThreads, wsql::BwdSumSmem<QP>::bytes, st>>>(
      map_mn, map_k, queries, summary, row_max, row_sum, d_summary, Q, n, tpc, accumulate, d_x, part_dK);
  return check_launch("sql_ws_bwd_sum_kernel");
}

int ws_bwd_sum(const float* x, const float* queries, const float* summary, const float* row_max, const float* row_sum,
               const float* d_summary, int B, int Q, int n, int accumulate, float* d_x, float* part_dK, int chunks, int tpc,
               cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  CUtensorMap map_mn, map_k;
  if (int e = make_tensor_map_2d(&map_mn, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = make_tensor_map_2d(&map_k, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 0)) return e;
  if (Q <= 64)
    return launch_ws_bwd_sum<64>(map_mn, map_k, queries, summary, row_max, row_sum, d_summary, B, Q, n, chunks, tpc, accumulate,
                                 d_x, part_dK, st);
  return launch_ws_bwd_sum<128>(map_mn, map_k, queries, summary, row_max, row_sum, d_summary, B, Q, n, chunks, tpc, accumulate,
                                d_x, part_dK, st);
}

void ws_summary_plan(int B, int n, int* chunks, int* steps_per_chunk) {
  const int steps = ceil_div(n, 32);
  int c = kNumSMs / B;   // one CTA per SM
  c = c < 1 ? 1 : (c > steps ? steps : c);
  *steps_per_chunk = ceil_div(steps, c);
  *chunks = ceil_div(steps, *steps_per_chunk);
}

template <int QP>
static int launch_ws_summary(const CUtensorMap& map_mn, const CUtensorMap& map_k, const float* queries, int B, int Q, int n,
                             int chunks, int spc, float* partial, cudaStream_t st) {
  if (int e = ensure_dyn_smem(wsql::sql_ws_summary_kernel<QP>, wsql::SumSmem<QP>::bytes)) return e;
  ProfScope prof("sql_tc_summary_kernel", st);
  wsql::sql_ws_summary_kernel<QP><<<dim3(chunks, B), wsql::kSumThreads, wsql::SumSmem<QP>::bytes, st>>>(map_mn, map_k, queries, Q,
                                                                                                  n, spc, partial);
  return check_launch("sql_ws_summary_kernel");
}

// writes per-chunk partial records [B][chunks][Q][34]; the caller runs the split-softmax combine
int ws_summary_partials(const float* x, const float* queries, int B, int Q, int n, float* partial, int* chunks_out,
                        cudaStream_t st) {
  SQLX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  int chunks, spc;
  ws_summary_plan(B, n, &chunks, &spc);
  CUtensorMap map_mn, map_k;
  if (int e = make_tensor_map_2d(&map_mn, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 1)) return e;
  if (int e = make_tensor_map_2d(&map_k, x, (uint64_t)B * wsql::kE, (uint64_t)n, 32, 32, 0)) return e;
  *chunks_out = chunks;
  if (Q <= 64) return launch_ws_summary<64>(map_mn, map_k, queries, B, Q, n, chunks, spc, partial, st);
  return launch_ws_summary<128>(map_mn, map_k, queries, B, Q, n, chunks, spc, partial, st);
}

}  // namespace sqlx

/* Limit the grid of the one-CTA-per-SM SQL kernels launched from now on to `sms` SMs (clamped to [8, 148]); returns the
 * previous budget.  Workspaces sized with the full budget stay sufficient. */
extern "C" int sqlx_sql_set_sm_budget(int sms) {
  sms = sms < 8 ? 8 : (sms > sqlx::kNumSMs ? sqlx::kNumSMs : sms);
  return sqlx::g_sm_budget.exchange(sms);
}
