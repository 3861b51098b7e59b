// Bring-up probe (not part of the product): can ONE shared-memory image of an x tile, written by TMA with
// SWIZZLE_128B_ATOM_32B, serve both as the MN-major tf32 operand (known good) and as a K-major operand whose descriptor
// names layout type 1 (SWIZZLE_128B_BASE32B)?  B logical [32 n][32 k] stored as rows = n, 32 k contiguous, 32-B chunks
// XOR (row & 3); A [128 m][32 k] in the ordinary K-major SWIZZLE_128B layout.  D = A B^T over K = 32 (four k-steps).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sfmnext-impl_b200/csrc/tc_common.cuh"
using namespace sqlx::tc;

struct Cfg { uint32_t lbo, sbo, layout; int kstep_bytes; };

__global__ void probe(Cfg c, const float* A, const float* B, float* D) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = base;            // 16 KB
  uint8_t* sb = base + 16384;    // 4 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) *(float*)(sa + sw128_offset(i / 32, i % 32)) = A[i];
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    *(float*)(sb + (c.layout == 1 ? b32_offset(n, k) : sw128_offset(n, k))) = B[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, 32, 0, 0);
    for (int k = 0; k < 4; ++k)
      umma_tf32_ss(tmem, make_desc_sw128(smem_u32(sa) + k * 32, 16, 1024),
                   make_desc(smem_u32(sb) + k * c.kstep_bytes, c.lbo, c.sbo, c.layout), idesc, k > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int cc = 0; cc < 32; cc += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cc, v);
    tmem_wait_ld();
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * 32 + cc + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> A(128 * 32), B(32 * 32), D(128 * 32);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : B) v = (float)(rand() % 5 - 2);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  Cfg cfgs[] = {
    {16, 1024, 2, 32},     // reference: ordinary SWIZZLE_128B K-major
    {16, 1024, 1, 32},     // BASE32B image, K-major descriptor, 8-row groups 1 KB apart
    {16, 512, 1, 32},
    {1024, 16, 1, 32},
    {512, 1024, 1, 32},
  };
  for (auto& c : cfgs) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, 40000>>>(c, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
      float exp = 0.f;
      for (int k = 0; k < 32; ++k) exp += A[m * 32 + k] * B[n * 32 + k];
      if (D[m * 32 + n] != exp) ++bad;
    }
    printf("layout=%u lbo=%u sbo=%u kstep=%d : %s, mismatches %d/4096\n", c.layout, c.lbo, c.sbo, c.kstep_bytes,
           cudaGetErrorString(e), bad);
  }
  return 0;
}
