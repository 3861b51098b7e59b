// Bring-up probe for tcgen05 operand layouts (not part of the product): one UMMA M=128,N=32,K=8 kind::tf32,
// A and B written to shared memory by threads under a chosen layout, D dumped to the host.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sfmnext-impl_b200/csrc/tc_common.cuh"
using namespace sqlx::tc;

struct Cfg { int a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int use_mask; };
__device__ __forceinline__ uint32_t b32_offset(int krow, int mn) {  // SWIZZLE_128B_BASE32B: 32-B chunks ^ (row & 3)
  return (uint32_t)krow * 128u + ((((uint32_t)mn >> 3) ^ ((uint32_t)krow & 3u)) << 5) + (((uint32_t)mn & 7u) << 2);
}
__device__ __forceinline__ uint64_t desc_b32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)1 << 61; return d; }


// A logical [128 m][8 k], B logical [32 n][8 k]
__global__ void probe(Cfg c, const float* A, const float* B, float* D) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = base;            // 16 KB region
  uint8_t* sb = base + 16384;    // 16 KB region
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) ((float*)base)[i] = 0.f;
  __syncthreads();
  // A
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int m = i / 8, k = i % 8;
    uint32_t off;
    if (c.a_mn) off = (m / 32) * 4096 + b32_offset(k, m % 32);   // blocks of 32 m; rows = k; 32 m contiguous
    else off = sw128_offset(m, k);                                   // rows = m (128 B pitch), k contiguous
    *(float*)(sa + off) = A[i];
  }
  for (int i = threadIdx.x; i < 32 * 8; i += blockDim.x) {
    const int n = i / 8, k = i % 8;
    uint32_t off;
    if (c.b_mn) off = b32_offset(k, n);
    else off = sw128_offset(n, k);
    *(float*)(sb + off) = B[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 32); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, 32, c.a_mn, c.b_mn);
    const uint64_t da = c.a_mn ? desc_b32(smem_u32(sa), c.a_lbo, c.a_sbo) : make_desc_sw128(smem_u32(sa), c.a_lbo, c.a_sbo);
    const uint64_t db = c.b_mn ? desc_b32(smem_u32(sb), c.b_lbo, c.b_sbo) : make_desc_sw128(smem_u32(sb), c.b_lbo, c.b_sbo);
    if (c.use_mask) {
      uint32_t z = 0, acc = 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
    } else {
      umma_tf32_ss(tmem, da, db, idesc, 0);
    }
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
  std::vector<float> A(128 * 8), B(32 * 8), D(128 * 32);
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 8; ++k) A[m * 8 + k] = m * 10 + k;
  for (int n = 0; n < 32; ++n) for (int k = 0; k < 8; ++k) B[n * 8 + k] = (n == k) ? 1.f : (n == 8 + k ? 2.f : 0.f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  Cfg cfgs[] = {
    {0, 0, 16, 1024, 16, 1024, 0},      // both K-major (reference configuration)
    {0, 0, 16, 1024, 16, 1024, 1},      // same with explicit lane mask operand
    {1, 0, 4096, 512, 16, 1024, 0},    // A MN-major BASE32B, LBO = 32-pixel block stride, SBO = 4-row k group
    {1, 0, 512, 4096, 16, 1024, 0},    // swapped
    {0, 1, 16, 1024, 4096, 512, 0},    // B MN-major BASE32B
    {0, 1, 16, 1024, 512, 4096, 0},
    {1, 1, 4096, 512, 4096, 512, 0},
  };
  for (auto& c : cfgs) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, 40000>>>(c, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
      float exp = n < 8 ? A[m * 8 + n] : (n < 16 ? 2.f * A[m * 8 + n - 8] : 0.f);
      if (D[m * 32 + n] != exp) ++bad;
    }
    printf("cfg a_mn=%d b_mn=%d a(lbo=%u,sbo=%u) b(lbo=%u,sbo=%u) mask=%d : %s, mismatches %d/4096\n", c.a_mn, c.b_mn,
           c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, c.use_mask, cudaGetErrorString(e), bad);
    for (int m : {0, 1, 9, 33, 100}) {
      printf("   m=%3d:", m);
      for (int n = 0; n < 18; ++n) printf(" %6.0f", D[m * 32 + n]);
      printf("\n");
    }
  }
  return 0;
}
