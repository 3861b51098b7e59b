// The three tiny per-sample products of the mixed-weight decomposition (sql_tc.cu):
//
//   forward   M[b,d,e]   = sum_q Wp[d,q] K[b,q,e]                  logits = Wp (K x) + b = (Wp K) x + b = M x + b
//   backward  dWp[d,q]   = sum_b sum_e dM[b,d,e] K[b,q,e]          (1x1 conv weight gradient, depth_decoder_QTR.py:28,61)
//             dK[b,q,e] += sum_d Wp[d,q] dM[b,d,e]                 (regression-path part of the query gradient)
//
// They replace three cuBLAS SGEMM launches + an einsum's two layout copies + an add inside the step:
// one launch forward, one backward, fixed summation order (deterministic), no workspace.  Sizes are tiny
// (B*D*E <= 16*128*64 outputs, Q <= 128 terms): one warp per output row keeps every load coalesced.
#include "common.cuh"

namespace sqlx {

// grid: B*D warps; lane = e (loops for E > 32).  Wp row broadcast, K rows coalesced.
__global__ void sql_mix_weights_kernel(const float* __restrict__ Wp, const float* __restrict__ K, int B, int Q, int D,
                                       int E, float* __restrict__ M) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * D) return;
  const int b = warp / D, d = warp - b * D;
  const float* wrow = Wp + (size_t)d * Q;
  const float* kb = K + (size_t)b * Q * E;
  for (int e = lane; e < E; e += 32) {
    float acc = 0.f;
    for (int q0 = 0; q0 < Q; q0 += 16) {         // 32 independent loads in flight per trip (the loop is latency-bound)
      float wv[16], kv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool on = q0 + j < Q;
        wv[j] = on ? __ldg(wrow + q0 + j) : 0.f;
        kv[j] = on ? __ldg(kb + (size_t)(q0 + j) * E + e) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(wv[j], kv[j], acc);
    }
    M[((size_t)b * D + d) * E + e] = acc;
  }
}

// One launch, two roles.  Warps [0, nW): dWp[d,q] = sum_b sum_e dM[b,d,e] K[b,q,e] (lane = e, then a warp sum), nW = D*Q
// when d_Wp is wanted (else 0); warps [nW, nW + B*Q): dK[b,q,:] (+)= sum_d Wp[d,q] dM[b,d,:] (lane = e).
__global__ void sql_mix_weights_bwd_kernel(const float* __restrict__ dM, const float* __restrict__ K,
                                           const float* __restrict__ Wp, int B, int Q, int D, int E, int accumulate_dK,
                                           float* __restrict__ dWp, float* __restrict__ dK) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nW = dWp ? D * Q : 0;
  if (warp < nW) {
    const int d = warp / Q, q = warp - d * Q;
    float acc = 0.f;
    for (int e = lane; e < E; e += 32) {
      for (int b0 = 0; b0 < B; b0 += 16) {       // all samples' loads in flight together
        float dv[16], kv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool on = b0 + j < B;
          dv[j] = on ? __ldg(dM + ((size_t)(b0 + j) * D + d) * E + e) : 0.f;
          kv[j] = on ? __ldg(K + ((size_t)(b0 + j) * Q + q) * E + e) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fmaf(dv[j], kv[j], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dWp[(size_t)d * Q + q] = acc;
    return;
  }
  const int w2 = warp - nW;
  if (!dK || w2 >= B * Q) return;
  const int b = w2 / Q, q = w2 - b * Q;
  const float* dmb = dM + (size_t)b * D * E;
  for (int e = lane; e < E; e += 32) {
    float acc = 0.f;
    for (int d0 = 0; d0 < D; d0 += 16) {
      float wv[16], mv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool on = d0 + j < D;
        wv[j] = on ? __ldg(Wp + (size_t)(d0 + j) * Q + q) : 0.f;
        mv[j] = on ? __ldg(dmb + (size_t)(d0 + j) * E + e) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(wv[j], mv[j], acc);
    }
    float* o = dK + ((size_t)b * Q + q) * E + e;
    *o = accumulate_dK ? *o + acc : acc;
  }
}

}  // namespace sqlx

using namespace sqlx;

extern "C" int sqlx_sql_mix_weights(const float* Wp, const float* queries, int B, int Q, int D, int E, float* Mx,
                                    void* stream) {
  SQLX_REQUIRE(Wp && queries && Mx, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && Q > 0 && D > 0 && E > 0, "non-positive shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long threads = (long long)B * D * 32;
  ProfScope prof("sql_mix_weights_kernel", st);
  sql_mix_weights_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(Wp, queries, B, Q, D, E, Mx);
  return check_launch("sql_mix_weights_kernel");
}

extern "C" int sqlx_sql_mix_weights_bwd(const float* d_Mx, const float* queries, const float* Wp, int B, int Q, int D,
                                        int E, int accumulate_d_queries, float* d_Wp, float* d_queries, void* stream) {
  SQLX_REQUIRE(d_Mx && queries && Wp && (d_Wp || d_queries), "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && Q > 0 && D > 0 && E > 0, "non-positive shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long threads = ((d_Wp ? (long long)D * Q : 0) + (d_queries ? (long long)B * Q : 0)) * 32;
  ProfScope prof("sql_mix_weights_bwd_kernel", st);
  sql_mix_weights_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_Mx, queries, Wp, B, Q, D, E,
                                                                                 accumulate_d_queries, d_Wp, d_queries);
  return check_launch("sql_mix_weights_bwd_kernel");
}
