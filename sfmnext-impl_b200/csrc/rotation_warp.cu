// inverse_rotation_warp of the indoor trainer's rectification step: layers.py:460-479.
//   R = euler2mat(rot);  P = K R;  w(u,v) = depth_to_3d(ones, K)(u,v) = ((u - cx)/fx, (v - cy)/fy, 1)   (kornia
//   geometry.depth.depth_to_3d / unproject_points: an un-pinned dependency absent from the reference tree; its published
//   pin-hole rule is restated here and in the test-side restatement);  c = P w;  pix = c.xy / (c.z + 1e-7);
//   out = F.grid_sample(img, normalised(pix), padding_mode="zeros", align_corners=True)   -- i.e. a bilinear gather at the
//   un-normalised position pix itself, taps outside the frame contributing zero.
// The kernel takes the 3x3 matrix P (the caller builds K . euler2mat(rot) with three tiny differentiable torch ops) and
// returns, in the backward, dL/dP = sum_pixels dL/dc w^T (fixed-order two-stage reduction); autograd carries it to `rot`.
#include "common.cuh"

namespace sqlx {

struct Tap2 {
  int x0, y0;
  float fx, fy;
};
__device__ __forceinline__ Tap2 taps_of(float px, float py) {
  Tap2 t;
  const float fx0 = floorf(px), fy0 = floorf(py);
  t.x0 = (int)fx0; t.y0 = (int)fy0;
  t.fx = px - fx0; t.fy = py - fy0;
  return t;
}
__device__ __forceinline__ float at0(const float* __restrict__ pl, int H, int W, int y, int x) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(pl + (size_t)y * W + x) : 0.f;     // padding_mode="zeros"
}

__global__ void rotation_warp_fwd_kernel(const float* __restrict__ img, const float* __restrict__ P,
                                         const float* __restrict__ K, int B, int H, int W, float* __restrict__ out) {
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / plane), r = (int)(i - (size_t)b * plane);
    const int v = r / W, u = r - v * W;
    const float* Kb = K + b * 9;
    const float* Pb = P + b * 9;
    const float w0 = ((float)u - Kb[2]) / Kb[0], w1 = ((float)v - Kb[5]) / Kb[4];
    const float c0 = Pb[0] * w0 + Pb[1] * w1 + Pb[2], c1 = Pb[3] * w0 + Pb[4] * w1 + Pb[5], c2 = Pb[6] * w0 + Pb[7] * w1 + Pb[8];
    const float z = c2 + 1e-7f;
    // the reference normalises by (W-1, H-1) and grid_sample(align_corners=True) un-normalises again
    const float px = ((((c0 / z) / (float)(W - 1) - 0.5f) * 2.f + 1.f) * 0.5f) * (float)(W - 1);
    const float py = ((((c1 / z) / (float)(H - 1) - 0.5f) * 2.f + 1.f) * 0.5f) * (float)(H - 1);
    const Tap2 t = taps_of(px, py);
    const float w00 = (1.f - t.fx) * (1.f - t.fy), w01 = t.fx * (1.f - t.fy), w10 = (1.f - t.fx) * t.fy, w11 = t.fx * t.fy;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = img + ((size_t)b * 3 + c) * plane;
      out[((size_t)b * 3 + c) * plane + r] = at0(pl, H, W, t.y0, t.x0) * w00 + at0(pl, H, W, t.y0, t.x0 + 1) * w01 +
                                             at0(pl, H, W, t.y0 + 1, t.x0) * w10 + at0(pl, H, W, t.y0 + 1, t.x0 + 1) * w11;
    }
  }
}

constexpr int kRwBlocks = 64;   // blocks per sample of the backward reduction

// partial [B][kRwBlocks][9]; dP [B][9] written by the last block of each sample (fixed-order sums)
__global__ void rotation_warp_bwd_kernel(const float* __restrict__ img, const float* __restrict__ P,
                                         const float* __restrict__ K, const float* __restrict__ g_out, int H, int W,
                                         float* __restrict__ partial, unsigned int* __restrict__ counter,
                                         float* __restrict__ dP) {
  __shared__ float red[32];
  __shared__ int is_last;
  const int b = blockIdx.y;
  const size_t plane = (size_t)H * W;
  const float* Kb = K + b * 9;
  const float* Pb = P + b * 9;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < (int)plane; r += gridDim.x * blockDim.x) {
    const int v = r / W, u = r - v * W;
    const float w0 = ((float)u - Kb[2]) / Kb[0], w1 = ((float)v - Kb[5]) / Kb[4];
    const float c0 = Pb[0] * w0 + Pb[1] * w1 + Pb[2], c1 = Pb[3] * w0 + Pb[4] * w1 + Pb[5], c2 = Pb[6] * w0 + Pb[7] * w1 + Pb[8];
    const float z = c2 + 1e-7f, rz = 1.f / z;
    const float qx = c0 * rz, qy = c1 * rz;
    const float px = (((qx / (float)(W - 1) - 0.5f) * 2.f + 1.f) * 0.5f) * (float)(W - 1);
    const float py = (((qy / (float)(H - 1) - 0.5f) * 2.f + 1.f) * 0.5f) * (float)(H - 1);
    const Tap2 t = taps_of(px, py);
    float gx = 0.f, gy = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = img + ((size_t)b * 3 + c) * plane;
      const float a = at0(pl, H, W, t.y0, t.x0), bq = at0(pl, H, W, t.y0, t.x0 + 1);
      const float cq = at0(pl, H, W, t.y0 + 1, t.x0), dq = at0(pl, H, W, t.y0 + 1, t.x0 + 1);
      const float g = __ldg(g_out + ((size_t)b * 3 + c) * plane + r);
      gx += g * ((bq - a) * (1.f - t.fy) + (dq - cq) * t.fy);
      gy += g * ((cq - a) * (1.f - t.fx) + (dq - bq) * t.fx);
    }
    // pix = c.xy / z  (the normalise / un-normalise round trip has unit derivative)
    const float g0 = gx * rz, g1 = gy * rz, g2 = -(gx * qx + gy * qy) * rz;
    acc[0] += g0 * w0; acc[1] += g0 * w1; acc[2] += g0;
    acc[3] += g1 * w0; acc[4] += g1 * w1; acc[5] += g1;
    acc[6] += g2 * w0; acc[7] += g2 * w1; acc[8] += g2;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float tsum = block_sum(acc[k], red);
    if (threadIdx.x == 0) partial[((size_t)b * gridDim.x + blockIdx.x) * 9 + k] = tsum;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(&counter[b], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 9) {
    float s = 0.f;
    for (int j = 0; j < (int)gridDim.x; ++j) s += *(volatile float*)&partial[((size_t)b * gridDim.x + j) * 9 + threadIdx.x];
    dP[b * 9 + threadIdx.x] = s;
  }
  if (threadIdx.x == 0) counter[b] = 0u;
}

}  // namespace sqlx

using namespace sqlx;

extern "C" size_t sqlx_rotation_warp_workspace_bytes(int B) {
  return B > 0 ? 256 + sizeof(unsigned int) * (size_t)B + 16 + sizeof(float) * (size_t)B * kRwBlocks * 9 : 0;
}

/* out[b,c,v,u] = bilinear(img[b,c], pix(u,v)), zeros outside, pix = (P_b w).xy / ((P_b w).z + 1e-7),
 * w = ((u - K[0][2]) / K[0][0], (v - K[1][2]) / K[1][1], 1): layers.py:460-479 with P = K . euler2mat(rot). */
extern "C" int sqlx_rotation_warp_fwd(const float* img, const float* P, const float* K3, int B, int H, int W, float* out,
                                      void* stream) {
  SQLX_REQUIRE(img && P && K3 && out, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 1 && W > 1 && (long long)H * W < (1ll << 30), "bad shape B=%d H=%d W=%d", B, H, W);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * H * W;
  const int blocks = (int)((total + 255) / 256 < 8 * kNumSMs ? (total + 255) / 256 : 8 * kNumSMs);
  ProfScope prof("rotation_warp_fwd_kernel", st);
  rotation_warp_fwd_kernel<<<blocks, 256, 0, st>>>(img, P, K3, B, H, W, out);
  return check_launch("rotation_warp_fwd_kernel");
}

/* d_P [B,3,3] = dL/dP given g_out [B,3,H,W]; workspace zero-initialised once by the caller (left zero) */
extern "C" int sqlx_rotation_warp_bwd(const float* img, const float* P, const float* K3, const float* g_out, int B, int H,
                                      int W, float* d_P, void* workspace, size_t workspace_bytes, void* stream) {
  SQLX_REQUIRE(img && P && K3 && g_out && d_P && workspace, "NULL pointer argument");
  SQLX_REQUIRE(B > 0 && H > 1 && W > 1 && (long long)H * W < (1ll << 30), "bad shape B=%d H=%d W=%d", B, H, W);
  SQLX_REQUIRE(workspace_bytes >= sqlx_rotation_warp_workspace_bytes(B), "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 256 + sizeof(unsigned int) * (size_t)B);
  partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(partial) + 15) & ~(uintptr_t)15);
  ProfScope prof("rotation_warp_bwd_kernel", st);
  rotation_warp_bwd_kernel<<<dim3(kRwBlocks, B), 256, 0, st>>>(img, P, K3, g_out, H, W, partial, counter, d_P);
  return check_launch("rotation_warp_bwd_kernel");
}
