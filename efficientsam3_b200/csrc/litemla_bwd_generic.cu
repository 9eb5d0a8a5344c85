// Backward of the ReLU linear attention for head dim 16 | 32 (efficientvit_b2: LiteMLA dim 32, ops.py:592-621) -- the
// generic-dim counterpart of es3_litemla_attn_bwd (train_bwd.cu, dim 16 only), paired with es3_litemla_attn_generic
// (litemla.cu): it consumes that kernel's partial KV sums [B][heads2][ceil(HW/128)][DIM+1][DIM].
//   q' = relu(q), k' = relu(k), vpad = [v, 1];  KV[j][i] = sum_n vpad[n][j] k'[n][i];  o[n][j] = sum_i KV[j][i] q'[n][i];
//   y[n][j] = o[n][j] / (o[n][DIM] + eps).   With r = 1 / (o[DIM] + eps):  do[j] = dy[j] r,  do[DIM] = -r sum_j dy[j] y[j];
//   dKV[j][i] = sum_n do[n][j] q'[n][i];  dq'[i] = sum_j KV[j][i] do[j];  dv[j] = sum_i dKV[j][i] k'[i];  dk'[i] = sum_j vpad[j] dKV[j][i].
// Thread per token, fp32, two-stage deterministic dKV reduction.  GPU parity: tests/test_zz_train_gpu.py
// (test_litemla_attn_bwd_generic, the efficientvit_b2 training step), green on a B200 since round 2.
#include "common.cuh"

namespace es3 {
namespace {
constexpr int LG_PX = 128;

template <int DIM>
__device__ __forceinline__ void load_row(const bf16* p, float* f) {
#pragma unroll
  for (int i = 0; i < DIM / 8; ++i) unpack8(__ldg(reinterpret_cast<const uint4*>(p) + i), f + 8 * i);
}

template <int DIM>
__device__ __forceinline__ void sum_parts(const float* __restrict__ src, int nchunk, float* dst, int tid, int nthr) {
  constexpr int E = (DIM + 1) * DIM;
  for (int e = tid; e < E; e += nthr) {
    float a = 0.f;
    for (int c = 0; c < nchunk; ++c) a += src[(long long)c * E + e];
    dst[e] = a;
  }
}

// do[0..DIM] of one token from q' (already relu'd), dy and KV (shared memory)
template <int DIM>
__device__ __forceinline__ void token_do(const float* skv, const float* q, const float* dy, float eps, float* dof) {
  float den = 0.f;
#pragma unroll
  for (int i = 0; i < DIM; ++i) den = fmaf(skv[DIM * DIM + i], q[i], den);
  const float r = 1.f / (den + eps);
  float dot = 0.f;
#pragma unroll 4
  for (int j = 0; j < DIM; ++j) {
    float o = 0.f;
#pragma unroll
    for (int i = 0; i < DIM; ++i) o = fmaf(skv[j * DIM + i], q[i], o);
    dof[j] = dy[j] * r;
    dot = fmaf(dy[j], o * r, dot);
  }
  dof[DIM] = -r * dot;
}

// grid (ceil(HW / 128), heads2, B), block 256.  dkv_part [B][heads2][nchunk_b][DIM+1][DIM].
template <int DIM>
__global__ void __launch_bounds__(256) litemla_dkv_g_kernel(const bf16* __restrict__ ms, long long ld, const bf16* __restrict__ dy,
                                                            long long lddy, const float* __restrict__ kv_part, int nchunk_f,
                                                            float* __restrict__ dkv_part, int HW, float eps) {
  constexpr int E = (DIM + 1) * DIM;
  __shared__ float skv[E];
  __shared__ float s_q[LG_PX][DIM + 1];
  __shared__ float s_do[LG_PX][DIM + 2];
  const int tid = threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z, heads2 = gridDim.y;
  sum_parts<DIM>(kv_part + ((long long)b * heads2 + h) * nchunk_f * E, nchunk_f, skv, tid, 256);
  __syncthreads();
  if (tid < LG_PX) {
    const int n = blockIdx.x * LG_PX + tid;
    float q[DIM], dyv[DIM], dof[DIM + 1];
    if (n < HW) {
      load_row<DIM>(ms + ((long long)b * HW + n) * ld + h * 3 * DIM, q);
      load_row<DIM>(dy + ((long long)b * HW + n) * lddy + h * DIM, dyv);
#pragma unroll
      for (int i = 0; i < DIM; ++i) q[i] = fmaxf(q[i], 0.f);
      token_do<DIM>(skv, q, dyv, eps, dof);
    } else {
#pragma unroll
      for (int i = 0; i < DIM; ++i) q[i] = 0.f;
#pragma unroll
      for (int j = 0; j <= DIM; ++j) dof[j] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < DIM; ++i) s_q[tid][i] = q[i];
#pragma unroll
    for (int j = 0; j <= DIM; ++j) s_do[tid][j] = dof[j];
  }
  __syncthreads();
  float* dst = dkv_part + (((long long)b * heads2 + h) * gridDim.x + blockIdx.x) * E;
  for (int e = tid; e < E; e += 256) {
    const int j = e / DIM, i = e % DIM;
    float a = 0.f;
#pragma unroll 8
    for (int n = 0; n < LG_PX; ++n) a = fmaf(s_do[n][j], s_q[n][i], a);
    dst[e] = a;
  }
}

// grid (ceil(HW / 128), heads2, B), block 128: thread = token.  dms [B][HW][lddms] in the q|k|v layout of ms.
template <int DIM>
__global__ void __launch_bounds__(128) litemla_dqkv_g_kernel(const bf16* __restrict__ ms, long long ld, const bf16* __restrict__ dy,
                                                             long long lddy, const float* __restrict__ kv_part, int nchunk_f,
                                                             const float* __restrict__ dkv_part, int nchunk_b, bf16* __restrict__ dms,
                                                             long long lddms, int HW, float eps) {
  constexpr int E = (DIM + 1) * DIM;
  __shared__ float skv[E];
  __shared__ float sdkv[E];
  const int tid = threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z, heads2 = gridDim.y;
  sum_parts<DIM>(kv_part + ((long long)b * heads2 + h) * nchunk_f * E, nchunk_f, skv, tid, 128);
  sum_parts<DIM>(dkv_part + ((long long)b * heads2 + h) * nchunk_b * E, nchunk_b, sdkv, tid, 128);
  __syncthreads();
  const int n = blockIdx.x * LG_PX + tid;
  if (n >= HW) return;
  const bf16* row = ms + ((long long)b * HW + n) * ld + h * 3 * DIM;
  bf16* orow = dms + ((long long)b * HW + n) * lddms + h * 3 * DIM;
  float q[DIM], dof[DIM + 1];
  {
    float dyv[DIM], qr[DIM];
    load_row<DIM>(row, q);
    load_row<DIM>(dy + ((long long)b * HW + n) * lddy + h * DIM, dyv);
#pragma unroll
    for (int i = 0; i < DIM; ++i) qr[i] = fmaxf(q[i], 0.f);
    token_do<DIM>(skv, qr, dyv, eps, dof);
  }
  // dq, 8 channels at a time
#pragma unroll
  for (int c8 = 0; c8 < DIM / 8; ++c8) {
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = c8 * 8 + u;
      float a = 0.f;
#pragma unroll
      for (int j = 0; j <= DIM; ++j) a = fmaf(skv[j * DIM + i], dof[j], a);
      o[u] = q[i] > 0.f ? a : 0.f;
    }
    reinterpret_cast<uint4*>(orow)[c8] = pack8(o);
  }
  float k[DIM], v[DIM];
  load_row<DIM>(row + DIM, k);
  load_row<DIM>(row + 2 * DIM, v);
  // dk
#pragma unroll
  for (int c8 = 0; c8 < DIM / 8; ++c8) {
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = c8 * 8 + u;
      float a = sdkv[DIM * DIM + i];                 // vpad[DIM] = 1
#pragma unroll
      for (int j = 0; j < DIM; ++j) a = fmaf(v[j], sdkv[j * DIM + i], a);
      o[u] = k[i] > 0.f ? a : 0.f;
    }
    reinterpret_cast<uint4*>(orow + DIM)[c8] = pack8(o);
  }
  // dv
#pragma unroll
  for (int c8 = 0; c8 < DIM / 8; ++c8) {
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = c8 * 8 + u;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < DIM; ++i) a = fmaf(sdkv[j * DIM + i], fmaxf(k[i], 0.f), a);
      o[u] = a;
    }
    reinterpret_cast<uint4*>(orow + 2 * DIM)[c8] = pack8(o);
  }
}

template <int DIM>
int launch(const bf16* ms, long long ld, const bf16* dy, long long lddy, const float* kv_part, int nchunk_f, float* dkv_ws, bf16* dms,
           long long lddms, int B, int HW, int heads2, float eps, cudaStream_t st) {
  const int nchunk_b = ceil_div(HW, LG_PX);
  dim3 grid(nchunk_b, heads2, B);
  litemla_dkv_g_kernel<DIM><<<grid, 256, 0, st>>>(ms, ld, dy, lddy, kv_part, nchunk_f, dkv_ws, HW, eps);
  ES3_LAUNCH_CHECK("litemla_dkv_g_kernel");
  litemla_dqkv_g_kernel<DIM><<<grid, 128, 0, st>>>(ms, ld, dy, lddy, kv_part, nchunk_f, dkv_ws, nchunk_b, dms, lddms, HW, eps);
  ES3_LAUNCH_CHECK("litemla_dqkv_g_kernel");
  return 0;
}
}  // namespace
}  // namespace es3

using namespace es3;

extern "C" long long es3_litemla_bwd_generic_ws_floats(int B, int HW, int heads2, int dim) {
  return (long long)B * heads2 * ceil_div(HW, LG_PX) * (dim + 1) * dim;
}

/* Backward of es3_litemla_attn_generic; kv_part = the workspace that call filled (nchunk_f = ceil(HW / 128)). */
extern "C" int es3_litemla_attn_bwd_generic(const void* ms, long long ld, const void* dy, long long lddy, const float* kv_part,
                                            int nchunk_f, float* dkv_ws, void* dms, long long lddms, int B, int HW, int heads2, int dim,
                                            float eps, void* stream) {
  ES3_REQUIRE(dim == 16 || dim == 32, "es3_litemla_attn_bwd_generic: dim=%d not instantiated (16, 32)", dim);
  ES3_REQUIRE(ld >= 3 * dim * heads2 && ld % 8 == 0 && lddy % 8 == 0 && lddms % 8 == 0 && lddy >= dim * heads2 && lddms >= 3 * dim * heads2,
              "es3_litemla_attn_bwd_generic: bad strides ld=%lld lddy=%lld lddms=%lld", ld, lddy, lddms);
  ES3_REQUIRE(nchunk_f == ceil_div(HW, 128), "es3_litemla_attn_bwd_generic: kv_part must come from es3_litemla_attn_generic (nchunk %d != %d)",
              nchunk_f, ceil_div(HW, 128));
  cudaStream_t st = (cudaStream_t)stream;
  if (dim == 16)
    return launch<16>((const bf16*)ms, ld, (const bf16*)dy, lddy, kv_part, nchunk_f, dkv_ws, (bf16*)dms, lddms, B, HW, heads2, eps, st);
  return launch<32>((const bf16*)ms, ld, (const bf16*)dy, lddy, kv_part, nchunk_f, dkv_ws, (bf16*)dms, lddms, B, HW, heads2, eps, st);
}
