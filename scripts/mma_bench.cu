// Micro-benchmark: issue rate of legacy mma.sync bf16 shapes on sm_100a (m16n8k8 vs m16n8k16), 8 warps / CTA, 4 independent accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/mma_bench scripts/mma_bench.cu && scripts/_bin/mma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int K16, int NACC>
__global__ void __launch_bounds__(256) bench(float* out, int iters, unsigned seed) {
  float d[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (K16)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        else
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                       : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int K16, int NACC>
static void run(const char* name, float* out, int ctas_per_sm) {
  const int blocks = 148 * ctas_per_sm, iters = 2048;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    bench<K16, NACC><<<blocks, 256>>>(out, iters, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double mmas = (double)blocks * 8 * iters * 4 * NACC;          // warp-level MMA instructions
  const double per_sm_per_clk = mmas / 148 / (ms * 1e-3 * 1.965e9);
  printf("%-28s %d CTAs/SM x 8 warps, %d accumulators: %.3f ms, %.3f MMA / clk / SM (at 1965 MHz), %.1f dense TFLOP/s\n", name, ctas_per_sm, NACC, ms,
         per_sm_per_clk, mmas * (K16 ? 4096.0 : 2048.0) / ms * 1e-9);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  run<0, 4>("mma.m16n8k8  bf16", out, 2);
  run<1, 4>("mma.m16n8k16 bf16", out, 2);
  run<0, 8>("mma.m16n8k8  bf16", out, 2);
  run<1, 8>("mma.m16n8k16 bf16", out, 2);
  run<0, 8>("mma.m16n8k8  bf16", out, 4);
  run<1, 8>("mma.m16n8k16 bf16", out, 4);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
