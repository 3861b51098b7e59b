// Microbenchmark (not part of the product): cost of a stream of tcgen05.mma kind::tf32 M=128, K=8 instructions issued by one
// thread, as a function of N, of how many INDEPENDENT accumulators the stream alternates between, and of where A lives
// (shared memory / TMEM).  Answers: is a chain of small accumulating MMAs latency-bound, and does interleaving help?
#include <cstdio>
#include <cstdlib>
#include "../sfmnext-impl_b200/csrc/tc_common.cuh"
using namespace sqlx::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
               : "+r"(laneid), "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}

template <int CH, int TS>
__global__ void probe(int N, int NI, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = base;             // 16 KB: A [128][32] K-major
  uint8_t* sb = base + 16384;     // 32 KB: B [256][32] K-major
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (slot != 0) __trap();        // the CTA owns all 512 columns: base 0, known at compile time (keeps every
  const uint32_t tmem = 0;        // UTCHMMA operand in uniform registers: no ELECT / BRA.U.ANY waterfall per instruction)
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t da = make_desc_sw128(smem_u32(sa), 16, 1024), db = make_desc_sw128(smem_u32(sb), 16, 1024);
    const uint32_t a_col = 480;   // A operand columns in TS mode (8 columns per instruction)
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      for (int it = 0; it < NI; it += 4 * CH) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const uint32_t d = tmem + (uint32_t)c * N;
            if (TS) umma_tf32_ts(d, tmem + a_col + kk * 8, db + 2 * kk, idesc, 1u);
            else umma_tf32_ss(d, da + 2 * kk, db + 2 * kk, idesc, 1u);
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
  printf("%-4s %-6s %-6s %10s %10s\n", "mode", "N", "chains", "issue cyc", "cyc/mma");
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 32, 64, 128, 256})
      for (int chains : {1, 2, 3, 4, 6, 8}) {
        if (chains * N > 448) continue;
        auto run = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000); kern<<<1, 128, 60000>>>(N, NI, d); };
        if (ts) { if (chains == 1) run(probe<1, 1>); else if (chains == 2) run(probe<2, 1>); else if (chains == 3) run(probe<3, 1>); else if (chains == 4) run(probe<4, 1>); else if (chains == 6) run(probe<6, 1>); else run(probe<8, 1>); }
        else { if (chains == 1) run(probe<1, 0>); else if (chains == 2) run(probe<2, 0>); else if (chains == 3) run(probe<3, 0>); else if (chains == 4) run(probe<4, 0>); else if (chains == 6) run(probe<6, 0>); else run(probe<8, 0>); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-4s %-6d %-6d %10.1f %10.1f\n", ts ? "TS" : "SS", N, chains, (double)h[0] / NI, (double)h[1] / NI);
      }
  return 0;
}
