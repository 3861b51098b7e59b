// Microbenchmark (B200 bring-up): issue cost of packed fp32x2 arithmetic (FFMA2 / FADD2, sm_100) against scalar FFMA /
// FADD, alone and interleaved with shared-memory loads -- decides whether packing the box-filter / SSIM arithmetic of the
// photometric kernels pays.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/f32x2_probe.cu -o tools/f32x2_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  __shared__ float sh[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = seed * i;
  __syncthreads();
  float a[8], b = seed + threadIdx.x * 1e-6f, c = 0.999f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (i + 1);
  float2 p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
  float ls = 0.f;
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {          // 8 independent scalar FFMA chains
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, b);
    } else if (MODE == 1) {   // the same 8 FMAs as 4 FFMA2
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = __ffma2_rn(p[i], c2, b2);
    } else if (MODE == 2) {   // 8 scalar FFMA + 4 LDS
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) ls += sh[(threadIdx.x + it * 4 + i * 37) & 2047];
    } else if (MODE == 3) {   // 4 FFMA2 + 4 LDS
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = __ffma2_rn(p[i], c2, b2);
#pragma unroll
      for (int i = 0; i < 4; ++i) ls += sh[(threadIdx.x + it * 4 + i * 37) & 2047];
    } else if (MODE == 4) {   // 8 scalar FADD
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i] + b;
    } else if (MODE == 5) {   // 4 FADD2
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = __fadd2_rn(p[i], b2);
    }
  }
  float r = ls;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) r += p[i].x + p[i].y;
  out[blockIdx.x * 256 + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(out, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<blocks, 256>>>(out, 1.0001f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double fma = (double)blocks * 256 * ITERS * 8;
  printf("%-28s %8.3f ms   %7.2f T fp-op-pairs/s (8 per thread-iteration)\n", name, ms, fma / ms / 1e9);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0>("8 x FFMA", out);
  run<1>("4 x FFMA2", out);
  run<2>("8 x FFMA + 4 LDS", out);
  run<3>("4 x FFMA2 + 4 LDS", out);
  run<4>("8 x FADD", out);
  run<5>("4 x FADD2", out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
