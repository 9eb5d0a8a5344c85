// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/ffma2_bench scripts/ffma2_bench.cu && scripts/_bin/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<unsigned long long&>(d))
               : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float s) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 a = make_float2(s, s * 0.999f), b = make_float2(0.001f, 0.002f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
        else { float2 t = acc[i]; acc[i] = b; ffma2(acc[i], t, a); }
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
  float* out;
  const int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&out, blocks * 256 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) bench<0><<<blocks, 256>>>(out, iters, 0.9999f); else bench<1><<<blocks, 256>>>(out, iters, 0.9999f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fmas = (double)blocks * 256 * iters * 8 * 8 * 2;
      if (rep == 2) printf("%s: %.3f ms, %.1f T-FMA/s (%.1f TFLOP/s)\n", mode == 0 ? "FFMA  (scalar)" : "FFMA2 (f32x2) ", ms, fmas / ms * 1e-9, 2 * fmas / ms * 1e-9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
