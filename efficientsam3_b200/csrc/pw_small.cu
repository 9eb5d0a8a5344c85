// Narrow pointwise convolutions (K, N <= 64 channels) on the CUDA cores:  out[m][n] = sum_k a[m][k] w[n][k] (+ residual[m][n]).
//
// The stage-0 / stage-1 1x1 convs of the students (16 / 32 / 64 channels at 512^2 .. 256^2 pixels; efficientvit/nn/ops.py:273-367) and
// their input gradients in the training step are 8.4 M-row, 16..64-column matrices: 0.3-0.5 GB of traffic and almost no math.  On the
// 128 x BN tcgen05 tiles of gemm_tc they ran at 0.4-1.5 TB/s (profiles/r2h_train_table.md: `gemm_tc[K=16,N=16]` 1.28 ms per launch for
// 537 MB): a 128-byte-swizzled TMA box of 64 channels is three quarters zero-fill at K = 16 and a tile is 4 KB of output.  Here a
// thread owns one pixel row: it reads its K channels with 16-byte loads (a warp covers 32 consecutive rows = one contiguous block),
// keeps them in registers as fp32, multiplies by the weight matrix held in shared memory as fp32 (broadcast float4 reads) and
// writes N bf16 channels with 16-byte stores.  fp32 accumulation, same rounding points as the tensor-core path (bf16 operands,
// fp32 sum, one rounding to bf16 at the end).
#include "common.cuh"

namespace es3 {
namespace {

template <int K, int N>
__global__ void __launch_bounds__(256) pw_small_kernel(const bf16* __restrict__ a, long long lda, const bf16* __restrict__ w, long long ldw,
                                                       bf16* __restrict__ out, long long ldo, const bf16* __restrict__ res, long long ldr,
                                                       long long M) {
  __shared__ __align__(16) float s_w[N * K];
  for (int i = threadIdx.x; i < N * K; i += 256) s_w[i] = __bfloat162float(w[(long long)(i / K) * ldw + (i % K)]);
  __syncthreads();
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  float x[K];
  const uint4* ap = reinterpret_cast<const uint4*>(a + m * lda);
#pragma unroll
  for (int j = 0; j < K / 8; ++j) unpack8(__ldg(ap + j), x + 8 * j);
  uint4* op = reinterpret_cast<uint4*>(out + m * ldo);
  const uint4* rp = res ? reinterpret_cast<const uint4*>(res + m * ldr) : nullptr;
#pragma unroll
  for (int n0 = 0; n0 < N; n0 += 8) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < K; k += 4) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 wv = *reinterpret_cast<const float4*>(s_w + (n0 + e) * K + k);     // warp-uniform address: broadcast
        acc[e] = fmaf(x[k], wv.x, acc[e]);
        acc[e] = fmaf(x[k + 1], wv.y, acc[e]);
        acc[e] = fmaf(x[k + 2], wv.z, acc[e]);
        acc[e] = fmaf(x[k + 3], wv.w, acc[e]);
      }
    }
    if (rp) {
      float r[8];
      unpack8(__ldg(rp + n0 / 8), r);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += r[e];
    }
    op[n0 / 8] = pack8(acc);
  }
}

template <int K, int N>
int launch_pw_small(const void* a, long long lda, const void* w, long long ldw, void* out, long long ldo, const void* res, long long ldr,
                    long long M, cudaStream_t st) {
  pw_small_kernel<K, N><<<(unsigned)((M + 255) / 256), 256, 0, st>>>((const bf16*)a, lda, (const bf16*)w, ldw, (bf16*)out, ldo,
                                                                     (const bf16*)res, ldr, M);
  ES3_LAUNCH_CHECK("pw_small_kernel");
  return 0;
}

}  // namespace
}  // namespace es3

using namespace es3;

/* out[m * ldo + n] = sum_k a[m * lda + k] * w[n * ldw + k] (+ residual[m * ldr + n]); bf16 in / out, fp32 accumulation; K, N in
 * {16, 32, 64} and not both 64.  Returns -1 (no error set) for any other shape or for unaligned operands: the caller uses
 * es3_gemm_bf16. */
extern "C" int es3_pw_small_bf16(const void* a, long long lda, const void* w, long long ldw, void* out, long long ldo, const void* residual,
                                 long long ldr, long long M, int N, int K, void* stream) {
  const bool okdim = (K == 16 || K == 32 || K == 64) && (N == 16 || N == 32 || N == 64) && !(K == 64 && N == 64);
  if (!okdim || M <= 0 || lda % 8 || ldo % 8 || (residual && ldr % 8) ||
      (((uintptr_t)a | (uintptr_t)out | (uintptr_t)residual) & 15) != 0)
    return -1;
  cudaStream_t st = (cudaStream_t)stream;
#define ES3_PW(KK, NN) if (K == KK && N == NN) return launch_pw_small<KK, NN>(a, lda, w, ldw, out, ldo, residual, ldr, M, st)
  ES3_PW(16, 16); ES3_PW(16, 32); ES3_PW(16, 64); ES3_PW(32, 16); ES3_PW(32, 32); ES3_PW(32, 64); ES3_PW(64, 16); ES3_PW(64, 32);
#undef ES3_PW
  return -1;
}
