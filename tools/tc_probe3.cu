// Bring-up probe (not part of the product): where do the 64 rows of an M = 64, cta_group::1 tcgen05.mma accumulator live in
// TMEM?  A [64 m][32 k] K-major SWIZZLE_128B with A[m][k] = (k == 0) ? m + 1 : 0, B [32 n][32 k] with B[n][k] = (k == 0),
// so D[m][n] = m + 1 for every n: each TMEM lane that holds row m reads back m + 1; untouched lanes keep the 0xff fill.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../sfmnext-impl_b200/csrc/tc_common.cuh"
using namespace sqlx::tc;

__global__ void probe(float* D, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = base;            // 8 KB: 64 rows
  uint8_t* sb = base + 16384;    // 4 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) *(float*)(sa + sw128_offset(i / 32, i % 32)) = (i % 32 == 0) ? (float)(i / 32 + 1) : 0.f;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) *(float*)(sb + sw128_offset(i / 32, i % 32)) = (i % 32 == 0) ? 1.f : 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 64); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  {   // fill the accumulator region with a marker
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = -7.f;
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 16, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_tf32(64, 32, 0, 0);
    const long long t0 = clock64();
    for (int rep = 0; rep < 64; ++rep)
      for (int k = 0; k < 4; ++k)
        umma_tf32_ss(tmem, make_desc_sw128(smem_u32(sa) + k * 32, 16, 1024), make_desc_sw128(smem_u32(sb) + k * 32, 16, 1024), idesc,
                     (k > 0 || rep > 0) ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
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
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<float> D(128 * 32);
  float* dD; long long* dc; long long hc;
  cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dc, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  probe<<<1, 128, 40000>>>(dD, dc);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
  printf("M=64 N=32 SS: %s; %.1f cycles per instruction (256 instructions)\n", cudaGetErrorString(e), (double)hc / 256.0);
  printf("lane -> value / 64 of column 0 (row + 1; -7 = untouched), and whether all 32 columns agree\n");
  for (int l = 0; l < 128; ++l) {
    bool same = true;
    for (int c = 1; c < 32; ++c) same &= D[l * 32 + c] == D[l * 32];
    printf("%3d:%6.1f%s%s", l, D[l * 32] / 64.f, same ? " " : "*", (l % 8 == 7) ? "\n" : "  ");
  }
  return 0;
}
