// Microbenchmark (not part of the product): issue cost of the exact tcgen05.mma mix of the summary kernel's control lane:
//   form 1: A in TMEM, B MN-major tf32 (SWIZZLE_128B_BASE32B), N = 32      (y = K x)
//   form 2: A in TMEM, B K-major (SWIZZLE_128B), N = 32                    (S += P x^T)
// in batches of 12 with / without a tcgen05.commit after every batch.
#include <cstdio>
#include <cstdlib>
#include "../sfmnext-impl_b200/csrc/tc_common.cuh"
using namespace sqlx::tc;

template <int MODE>   // 0: form 2 only   1: form 1 only   2: alternate batches   +4: commit after every batch of 12
__global__ void probe(int NI, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sb = base;
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_barrier_init(); mbar_arrive(&bar3); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t id1 = make_idesc_tf32(128, 32, 0, 1), id2 = make_idesc_tf32(128, 32, 0, 0);
    const uint64_t d1 = make_desc_mn32(smem_u32(sb), 4096), d2 = make_desc_sw128(smem_u32(sb) + 8192, 16, 1024);
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      for (int it = 0; it < NI; it += 24) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const bool f1 = (MODE & 3) == 1 || ((MODE & 3) == 2 && half == 0);
          const uint32_t d = tmem + (uint32_t)((it / 24) & 3) * 96 + (f1 ? 0 : 64);
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            if (f1) umma_tf32_ts(d, tmem + 384 + (k & 3) * 8 + (k >= 4 && k < 8 ? 32 : 0), d1 + (uint64_t)((k & 3) * 64), id1, 1u);
            else umma_tf32_ts(d, tmem + ((it / 24) & 3) * 96 + (k & 3) * 8, d2 + (uint64_t)((k & 3) * 2), id2, 1u);
          }
          if ((MODE & 4) && half == 0) umma_commit(&bar2);
          if (MODE & 8) { mbar_wait(&bar3, 0); if (MODE & 16) tc_fence_after(); }
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  const int NI = 960;
  auto run = [&](auto kern, const char* name) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    kern<<<1, 128, 40000>>>(NI, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s issue %6.1f  total %6.1f cycles / mma\n", name, (double)h[0] / NI, (double)h[1] / NI);
  };
  run(probe<0>, "form 2 (B K-major)");
  run(probe<1>, "form 1 (B MN-major)");
  run(probe<2>, "alternating batches of 12");
  run(probe<4>, "form 2, commit every 24");
  run(probe<5>, "form 1, commit every 24");
  run(probe<6>, "alternating, commit every 24");
  run(probe<2 + 8>, "alternating, satisfied mbar_wait every 12");
  run(probe<6 + 8>, "alternating, commit/24 + wait/12");
  run(probe<6 + 8 + 16>, "alternating, commit/24 + wait/12 + fence");
  return 0;
}
