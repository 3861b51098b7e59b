// Microbenchmark (B200 bring-up): MUFU.EX2 and MUFU.RCP issue rate per SM, alone and mixed with FFMA -- the softmax
// epilogues of the SQL kernels execute one ex2 per (pixel, bin).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
__device__ __forceinline__ float ex2f(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcpf(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (i + 1) * 1e-3f + threadIdx.x * 1e-7f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.f;
      else if (MODE == 1) a[i] = rcpf(a[i] + 1.5f);
      else if (MODE == 2) { a[i] = ex2f(a[i]) - 1.f; a[i] = fmaf(a[i], 0.5f, 0.1f); a[i] = fmaf(a[i], 0.5f, 0.1f); a[i] = fmaf(a[i], 0.5f, 0.1f); }
      else a[i] = fmaf(a[i], 0.999f, 0.001f);
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, float* out, int threads) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = 148 * 2;
  k<MODE><<<blocks, threads>>>(out, 1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<blocks, threads>>>(out, 1.0001f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads * ITERS * 8;
  // per SM per clock at 1.965 GHz (nominal; the probe prints the raw rate too)
  printf("%-34s %3d thr/CTA %8.3f ms  %8.2f Gop/s  = %.2f ops/clk/SM @1.965GHz\n", name, threads, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 512 * sizeof(float));
  run<0>("ex2 (+1 FADD)", out, 512); run<0>("ex2 (+1 FADD)", out, 128);
  run<1>("rcp (+1 FADD)", out, 512);
  run<2>("ex2 + FADD + 3 FFMA", out, 512);
  run<3>("FFMA only", out, 512);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
