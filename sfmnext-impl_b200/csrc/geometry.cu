// Module-level geometry drop-ins: BackprojectDepth (layers.py:186-215) and Project3D (layers.py:236-258),
// forward and backward.  The fused training path (photo_fwd.cu / photo_bwd.cu) never materialises the
// [B,4,N] point cloud or the [B,H,W,2] sampling grid; these kernels exist so the reference's own
// nn.Module call sites (trainer.py:423-425) keep working one-for-one.
#include "common.cuh"

namespace sqlx {

__global__ void backproject_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ invK, int H, int W,
                                       float* __restrict__ points) {
  const int b = blockIdx.y, N = H * W;
  const float* iK = invK + b * 16;
  const float k00 = iK[0], k01 = iK[1], k02 = iK[2], k10 = iK[4], k11 = iK[5], k12 = iK[6], k20 = iK[8], k21 = iK[9],
              k22 = iK[10];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N; idx += gridDim.x * blockDim.x) {
    const int v = idx / W, u = idx - v * W;
    const float d = __ldg(depth + (size_t)b * N + idx);
    float* out = points + (size_t)b * 4 * N + idx;
    out[0] = d * (k00 * u + k01 * v + k02);
    out[N] = d * (k10 * u + k11 * v + k12);
    out[2 * (size_t)N] = d * (k20 * u + k21 * v + k22);
    out[3 * (size_t)N] = 1.f;
  }
}

__global__ void backproject_bwd_kernel(const float* __restrict__ g_points, const float* __restrict__ invK, int H, int W,
                                       float* __restrict__ d_depth) {
  const int b = blockIdx.y, N = H * W;
  const float* iK = invK + b * 16;
  const float k00 = iK[0], k01 = iK[1], k02 = iK[2], k10 = iK[4], k11 = iK[5], k12 = iK[6], k20 = iK[8], k21 = iK[9],
              k22 = iK[10];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N; idx += gridDim.x * blockDim.x) {
    const int v = idx / W, u = idx - v * W;
    const float* g = g_points + (size_t)b * 4 * N + idx;
    d_depth[(size_t)b * N + idx] = g[0] * (k00 * u + k01 * v + k02) + g[N] * (k10 * u + k11 * v + k12) +
                                   g[2 * (size_t)N] * (k20 * u + k21 * v + k22);
  }
}

__device__ __forceinline__ void make_P(const float* __restrict__ K, const float* __restrict__ T, float* P) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(K[i * 4 + k], T[k * 4 + j], acc);
      P[i * 4 + j] = acc;
    }
}

__global__ void project_fwd_kernel(const float* __restrict__ points, const float* __restrict__ K,
                                   const float* __restrict__ T, int H, int W, float eps, float* __restrict__ grid) {
  __shared__ float P[12];
  const int b = blockIdx.y, N = H * W;
  if (threadIdx.x == 0) make_P(K + b * 16, T + b * 16, P);
  __syncthreads();
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N; idx += gridDim.x * blockDim.x) {
    const float* pt = points + (size_t)b * 4 * N + idx;
    const float X = pt[0], Y = pt[N], Z = pt[2 * (size_t)N], Wh = pt[3 * (size_t)N];
    const float c0 = P[0] * X + P[1] * Y + P[2] * Z + P[3] * Wh;
    const float c1 = P[4] * X + P[5] * Y + P[6] * Z + P[7] * Wh;
    const float c2 = P[8] * X + P[9] * Y + P[10] * Z + P[11] * Wh;
    const float z = c2 + eps;
    float2 o;
    o.x = ((c0 / z) / (float)(W - 1) - 0.5f) * 2.f;
    o.y = ((c1 / z) / (float)(H - 1) - 0.5f) * 2.f;
    reinterpret_cast<float2*>(grid)[(size_t)b * N + idx] = o;
  }
}

__global__ void project_bwd_kernel(const float* __restrict__ points, const float* __restrict__ K,
                                   const float* __restrict__ T, const float* __restrict__ g_grid, int H, int W,
                                   float eps, float* __restrict__ d_points, float* __restrict__ dP /*[B,12] zeroed*/) {
  __shared__ float P[12];
  __shared__ float red[32];
  const int b = blockIdx.y, N = H * W;
  if (threadIdx.x == 0) make_P(K + b * 16, T + b * 16, P);
  __syncthreads();
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  const float sx = 2.f / (float)(W - 1), sy = 2.f / (float)(H - 1);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N; idx += gridDim.x * blockDim.x) {
    const float* pt = points + (size_t)b * 4 * N + idx;
    const float pv[4] = {pt[0], pt[N], pt[2 * (size_t)N], pt[3 * (size_t)N]};
    const float c0 = P[0] * pv[0] + P[1] * pv[1] + P[2] * pv[2] + P[3] * pv[3];
    const float c1 = P[4] * pv[0] + P[5] * pv[1] + P[6] * pv[2] + P[7] * pv[3];
    const float c2 = P[8] * pv[0] + P[9] * pv[1] + P[10] * pv[2] + P[11] * pv[3];
    const float iz = 1.f / (c2 + eps);
    const float2 g = reinterpret_cast<const float2*>(g_grid)[(size_t)b * N + idx];
    const float g0 = g.x * sx * iz, g1 = g.y * sy * iz;
    const float g2 = -(g0 * c0 + g1 * c1) * iz;
    if (d_points) {
      float* dp = d_points + (size_t)b * 4 * N + idx;
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[(size_t)j * N] = g0 * P[j] + g1 * P[4 + j] + g2 * P[8 + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] = fmaf(g0, pv[j], acc[j]);
      acc[4 + j] = fmaf(g1, pv[j], acc[4 + j]);
      acc[8 + j] = fmaf(g2, pv[j], acc[8 + j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float t = block_sum(acc[i], red);
    if (threadIdx.x == 0) atomicAdd(dP + b * 12 + i, t);
  }
}

// dT[b] = K[b][:3,:]^T dP[b]
__global__ void dT_from_dP12_kernel(const float* __restrict__ K, const float* __restrict__ dP, int B,
                                    float* __restrict__ dT) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 16) return;
  const int e = idx & 15, b = idx >> 4, i = e >> 2, j = e & 3;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) acc += K[b * 16 + k * 4 + i] * dP[b * 12 + k * 4 + j];
  dT[idx] = acc;
}

}  // namespace sqlx

using namespace sqlx;

static int check_geo(int B, int H, int W) {
  SQLX_REQUIRE(B > 0 && B <= 65535 && H > 1 && W > 1, "bad shape B=%d H=%d W=%d", B, H, W);
  return SQLX_OK;
}

extern "C" int sqlx_backproject_fwd(const float* depth, const float* inv_K, int B, int H, int W, float* points,
                                    void* stream) {
  if (int e = check_geo(B, H, W)) return e;
  SQLX_REQUIRE(depth && inv_K && points, "NULL pointer argument");
  dim3 grid(min(ceil_div(H * W, 256), 2 * kNumSMs), B);
  backproject_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(depth, inv_K, H, W, points);
  return check_launch("backproject_fwd_kernel");
}

extern "C" int sqlx_backproject_bwd(const float* g_points, const float* inv_K, int B, int H, int W, float* d_depth,
                                    void* stream) {
  if (int e = check_geo(B, H, W)) return e;
  SQLX_REQUIRE(g_points && inv_K && d_depth, "NULL pointer argument");
  dim3 grid(min(ceil_div(H * W, 256), 2 * kNumSMs), B);
  backproject_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g_points, inv_K, H, W, d_depth);
  return check_launch("backproject_bwd_kernel");
}

extern "C" int sqlx_project_fwd(const float* points, const float* K, const float* T, int B, int H, int W, float eps,
                                float* grid_out, void* stream) {
  if (int e = check_geo(B, H, W)) return e;
  SQLX_REQUIRE(points && K && T && grid_out, "NULL pointer argument");
  dim3 grid(min(ceil_div(H * W, 256), 2 * kNumSMs), B);
  project_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(points, K, T, H, W, eps, grid_out);
  return check_launch("project_fwd_kernel");
}

extern "C" int sqlx_project_bwd(const float* points, const float* K, const float* T, const float* g_grid, int B, int H,
                                int W, float eps, float* d_points, float* d_T, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (int e = check_geo(B, H, W)) return e;
  SQLX_REQUIRE(points && K && T && g_grid && d_T, "NULL pointer argument");
  SQLX_REQUIRE(workspace && workspace_bytes >= sizeof(float) * 12 * (size_t)B, "workspace too small (need 48*B bytes)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* dP = reinterpret_cast<float*>(workspace);
  if (cudaMemsetAsync(dP, 0, sizeof(float) * 12 * (size_t)B, st) != cudaSuccess) return check_launch("cudaMemsetAsync");
  dim3 grid(min(ceil_div(H * W, 256), 64), B);
  project_bwd_kernel<<<grid, 256, 0, st>>>(points, K, T, g_grid, H, W, eps, d_points, dP);
  if (int e = check_launch("project_bwd_kernel")) return e;
  dT_from_dP12_kernel<<<ceil_div(B * 16, 128), 128, 0, st>>>(K, dP, B, d_T);
  return check_launch("dT_from_dP12_kernel");
}
