// Backward kernels of the TinyViT student (sam3/sam3/backbones/tiny_vit.py) that the shared training kernels of train_bwd.cu do
// not cover:
//   es3_layernorm_bwd        nn.LayerNorm backward over bf16 rows (Attention.norm :262, Mlp.norm :205): dx (+ an optional residual
//                            gradient), dgamma += sum dy xhat, dbeta += sum dy
//   es3_win_attn_bias_bwd    backward of es3_win_attn_bias_bf16 (Attention.forward :264-293 on window partitions) on a token map
//                            whose extent is a multiple of the window: dqkv in the forward's [q|k|v]-per-head layout and the
//                            per-window score gradients dS (fp32); the bias gradient is the column sum of dS over the windows
//                            (es3_colsum_f32, fixed order)
// Correctness-first CUDA-core formulations (thread per row / per key, fixed-order reductions, no atomics).  GPU parity:
// tests/test_zz_train_gpu.py (test_layernorm_bwd, test_win_attn_bias_bwd, the TinyViT training step), green on a B200 since round 2.
#include "common.cuh"

namespace es3 {
namespace {

// ------------------------------------------------------------------------------------------ LayerNorm backward
// y = (x - mu) rstd gamma + beta.  With g = dy gamma, xh = (x - mu) rstd:  dx = rstd (g - mean(g) - xh mean(g xh)).
// grid = nblk, block 256 = 8 warps; a warp owns a row at a time (rows blockIdx.x * rows_per_block + warp, + 8, ...); lane l owns the
// 8-channel vectors l, l + 32, ... (C <= 1024: at most 4 per lane) and keeps its dgamma / dbeta partial sums in registers over all
// the rows of the block; the 8 warps are then summed through shared memory.  part [nblk][2][C].
constexpr int LNB_MAXV = 4;
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            const float* __restrict__ gamma, const bf16* __restrict__ dres, float eps,
                                                            bf16* __restrict__ dx, long long M, int C, long long rows_per_block,
                                                            float* __restrict__ part) {
  __shared__ float red[8][1024];       // cross-warp sums, dgamma then dbeta (32 KB)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CV = C >> 3;
  float dg[LNB_MAXV][8], db[LNB_MAXV][8];
#pragma unroll
  for (int v = 0; v < LNB_MAXV; ++v)
#pragma unroll
    for (int k = 0; k < 8; ++k) dg[v][k] = db[v][k] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  const float inv_c = 1.f / (float)C;
  for (long long r = r0 + warp; r < r1; r += 8) {
    float xv[LNB_MAXV][8], gv[LNB_MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < LNB_MAXV; ++v) {
      const int cv = lane + 32 * v;
      if (cv < CV) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * C + cv * 8)), xv[v]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += xv[v][k];
      }
    }
    const float mu = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < LNB_MAXV; ++v) {
      const int cv = lane + 32 * v;
      if (cv < CV) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = xv[v][k] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int v = 0; v < LNB_MAXV; ++v) {
      const int cv = lane + 32 * v;
      if (cv < CV) {
        float d[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * C + cv * 8)), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (xv[v][k] - mu) * rstd;
          xv[v][k] = xh;
          dg[v][k] = fmaf(d[k], xh, dg[v][k]);
          db[v][k] += d[k];
          const float g = d[k] * __ldg(gamma + cv * 8 + k);
          gv[v][k] = g;
          sg += g;
          sgx = fmaf(g, xh, sgx);
        }
      }
    }
    const float c1 = warp_sum(sg) * inv_c, c2 = warp_sum(sgx) * inv_c;
#pragma unroll
    for (int v = 0; v < LNB_MAXV; ++v) {
      const int cv = lane + 32 * v;
      if (cv < CV) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rstd * (gv[v][k] - c1 - xv[v][k] * c2);
        if (dres) {
          float rr[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(dres + r * C + cv * 8)), rr);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += rr[k];
        }
        *reinterpret_cast<uint4*>(dx + r * C + cv * 8) = pack8(o);
      }
    }
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    __syncthreads();
#pragma unroll
    for (int v = 0; v < LNB_MAXV; ++v) {
      const int cv = lane + 32 * v;
      if (cv < CV) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp][cv * 8 + k] = which ? db[v][k] : dg[v][k];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      part[((long long)blockIdx.x * 2 + which) * C + c] = t;
    }
  }
}

// dgamma[c] += sum_b part[b][0][c], dbeta[c] += sum_b part[b][1][c]   (fixed order)
__global__ void layernorm_bwd_finalize_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ dgamma,
                                              float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < nblk; ++k) {
    a += part[((long long)k * 2) * C + c];
    b += part[((long long)k * 2 + 1) * C + c];
  }
  if (dgamma) dgamma[c] += a;
  if (dbeta) dbeta[c] += b;
}

// ------------------------------------------------------------------------------------------ window attention backward
// One CTA per (window, head), 256 threads, N = ws^2 <= 196 tokens of head dim 32.  Forward: s_ij = scale q_i.k_j + bias[h][i][j],
// P = softmax_j(s), o_i = sum_j P_ij v_j.  Backward with do_i: dP_ij = do_i.v_j, D_i = sum_j P_ij dP_ij, dS_ij = P_ij (dP_ij - D_i),
// dq_i = scale sum_j dS_ij k_j, dk_j = scale sum_i dS_ij q_i, dv_j = sum_i P_ij do_i, dbias = sum over windows of dS.
// Phase 1: thread i (row): row max, row sum, D_i, then dq_i and the dS row (written to global in fp32: the bias gradient is its
// sum over thousands of windows, and rounding every term to bf16 first cost three digits of it).  Phase 2: thread j (key):
// dk_j, dv_j with P_ij recomputed from the saved row statistics.  Tiles live in shared memory as fp32 [N][33].
constexpr int WB_HD = 32, WB_LD = WB_HD + 1;

struct WinBwdArgs {
  const bf16* qkv;    // [B*H*W, 3*C]: head h at columns [96 h, +32) q | [+32, +64) k | [+64, +96) v
  const bf16* dout;   // [B*H*W, C]: head h at columns [32 h, +32)
  const float* bias;  // [heads][N][N]
  bf16* dqkv;         // like qkv
  float* dS;          // [B * nWin][ldS] fp32, row = [heads][N][N] (ldS >= heads N N); summed over the windows by es3_colsum_f32
  long long ldS;
  int H, W, C, ws, nWx, nWin, N;
  float scale;
};

__global__ void __launch_bounds__(256) win_attn_bias_bwd_kernel(const WinBwdArgs a) {
  extern __shared__ float wb_smem[];
  const int N = a.N;
  float* s_q = wb_smem;
  float* s_k = s_q + N * WB_LD;
  float* s_v = s_k + N * WB_LD;
  float* s_do = s_v + N * WB_LD;
  float* s_m = s_do + N * WB_LD;      // row max
  float* s_l = s_m + N;               // 1 / row sum
  float* s_D = s_l + N;               // D_i
  const int tid = threadIdx.x;
  const int head = blockIdx.y;
  const int b = blockIdx.x / a.nWin, wi = blockIdx.x % a.nWin;
  const int wy = wi / a.nWx, wx = wi % a.nWx;
  const int ld = 3 * a.C;
  // gather the window's q, k, v, do rows
  for (int i = tid; i < N * WB_HD; i += 256) {
    const int r = i / WB_HD, c = i % WB_HD;
    const long long tok = (long long)b * a.H * a.W + (long long)(wy * a.ws + r / a.ws) * a.W + (wx * a.ws + r % a.ws);
    const bf16* row = a.qkv + tok * ld + head * 96;
    s_q[r * WB_LD + c] = __bfloat162float(row[c]);
    s_k[r * WB_LD + c] = __bfloat162float(row[32 + c]);
    s_v[r * WB_LD + c] = __bfloat162float(row[64 + c]);
    s_do[r * WB_LD + c] = __bfloat162float(a.dout[tok * a.C + head * 32 + c]);
  }
  __syncthreads();
  const float* bias_h = a.bias + (long long)head * N * N;
  if (tid < N) {
    const int i = tid;
    float q[WB_HD], d_o[WB_HD];
#pragma unroll
    for (int c = 0; c < WB_HD; ++c) { q[c] = s_q[i * WB_LD + c]; d_o[c] = s_do[i * WB_LD + c]; }
    const float* bi = bias_h + (long long)i * N;
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < WB_HD; ++c) s = fmaf(q[c], s_k[j * WB_LD + c], s);
      mx = fmaxf(mx, fmaf(s, a.scale, bi[j]));
    }
    float l = 0.f, D = 0.f;
    for (int j = 0; j < N; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < WB_HD; ++c) {
        s = fmaf(q[c], s_k[j * WB_LD + c], s);
        dp = fmaf(d_o[c], s_v[j * WB_LD + c], dp);
      }
      const float e = __expf(fmaf(s, a.scale, bi[j]) - mx);
      l += e;
      D = fmaf(e, dp, D);
    }
    const float inv_l = 1.f / l;
    D *= inv_l;
    s_m[i] = mx; s_l[i] = inv_l; s_D[i] = D;
    float dq[WB_HD];
#pragma unroll
    for (int c = 0; c < WB_HD; ++c) dq[c] = 0.f;
    float* dS_row = a.dS + (long long)blockIdx.x * a.ldS + ((long long)head * N + i) * N;
    for (int j = 0; j < N; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < WB_HD; ++c) {
        s = fmaf(q[c], s_k[j * WB_LD + c], s);
        dp = fmaf(d_o[c], s_v[j * WB_LD + c], dp);
      }
      const float p = __expf(fmaf(s, a.scale, bi[j]) - mx) * inv_l;
      const float ds = p * (dp - D);
      dS_row[j] = ds;
#pragma unroll
      for (int c = 0; c < WB_HD; ++c) dq[c] = fmaf(ds, s_k[j * WB_LD + c], dq[c]);
    }
    const long long tok = (long long)b * a.H * a.W + (long long)(wy * a.ws + i / a.ws) * a.W + (wx * a.ws + i % a.ws);
    bf16* o = a.dqkv + tok * ld + head * 96;
#pragma unroll
    for (int c = 0; c < WB_HD; ++c) o[c] = __float2bfloat16(dq[c] * a.scale);
  }
  __syncthreads();
  if (tid < N) {
    const int j = tid;
    float k[WB_HD], v[WB_HD], dk[WB_HD], dv[WB_HD];
#pragma unroll
    for (int c = 0; c < WB_HD; ++c) { k[c] = s_k[j * WB_LD + c]; v[c] = s_v[j * WB_LD + c]; dk[c] = 0.f; dv[c] = 0.f; }
    for (int i = 0; i < N; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < WB_HD; ++c) {
        s = fmaf(s_q[i * WB_LD + c], k[c], s);
        dp = fmaf(s_do[i * WB_LD + c], v[c], dp);
      }
      const float p = __expf(fmaf(s, a.scale, bias_h[(long long)i * N + j]) - s_m[i]) * s_l[i];
      const float ds = p * (dp - s_D[i]);
#pragma unroll
      for (int c = 0; c < WB_HD; ++c) {
        dk[c] = fmaf(ds, s_q[i * WB_LD + c], dk[c]);
        dv[c] = fmaf(p, s_do[i * WB_LD + c], dv[c]);
      }
    }
    const long long tok = (long long)b * a.H * a.W + (long long)(wy * a.ws + j / a.ws) * a.W + (wx * a.ws + j % a.ws);
    bf16* o = a.dqkv + tok * ld + head * 96;
#pragma unroll
    for (int c = 0; c < WB_HD; ++c) {
      o[32 + c] = __float2bfloat16(dk[c] * a.scale);
      o[64 + c] = __float2bfloat16(dv[c]);
    }
  }
}


// ------------------------------------------------------------------------------------------ fp32 column sums
// out[c] += sum_r src[r][c], two stages in a fixed order (no atomics): block (x = 128-column slab, y = row split) sums rows
// y, y + nsplit, ... with coalesced 512-byte row reads, then one thread per column adds the nsplit partials.
__global__ void __launch_bounds__(128) colsum_f32_kernel(const float* __restrict__ src, long long ld, long long M, int L, int nsplit,
                                                         float* __restrict__ part) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= L) return;
  float a0 = 0.f, a1 = 0.f;
  long long r = blockIdx.y;
  for (; r + nsplit < M; r += 2 * (long long)nsplit) {
    a0 += src[r * ld + c];
    a1 += src[(r + nsplit) * ld + c];
  }
  if (r < M) a0 += src[r * ld + c];
  part[(long long)blockIdx.y * L + c] = a0 + a1;
}
__global__ void __launch_bounds__(128) colsum_f32_finalize_kernel(const float* __restrict__ part, int nsplit, int L, float* __restrict__ out) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= L) return;
  float a = 0.f;
  for (int k = 0; k < nsplit; ++k) a += part[(long long)k * L + c];
  out[c] += a;
}
}  // namespace
}  // namespace es3

using namespace es3;

static int lnb_blocks(long long M, long long* rpb) {
  long long nblk = (M + 63) / 64;                 // >= 64 rows (8 per warp) per block
  if (nblk > 592) nblk = 592;
  if (nblk < 1) nblk = 1;
  *rpb = (M + nblk - 1) / nblk;
  return (int)((M + *rpb - 1) / *rpb);
}

extern "C" long long es3_layernorm_bwd_ws_floats(long long M, int C) {
  long long rpb;
  return (long long)lnb_blocks(M, &rpb) * 2 * C;
}

/* dx = d LayerNorm(x) (+ dres); dgamma += sum_rows dy xhat, dbeta += sum_rows dy.  x, dy, dres, dx: [M][C] bf16 contiguous. */
extern "C" int es3_layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* dres, float eps, void* dx, long long M,
                                 int C, float* ws, float* dgamma, float* dbeta, void* stream) {
  ES3_REQUIRE(M > 0 && C % 8 == 0 && C <= 1024, "es3_layernorm_bwd: C=%d must be a multiple of 8 and <= 1024", C);
  long long rpb;
  const int nblk = lnb_blocks(M, &rpb);
  cudaStream_t st = (cudaStream_t)stream;
  layernorm_bwd_kernel<<<nblk, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, gamma, (const bf16*)dres, eps, (bf16*)dx, M, C, rpb, ws);
  ES3_LAUNCH_CHECK("layernorm_bwd_kernel");
  layernorm_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, nblk, C, dgamma, dbeta);
  ES3_LAUNCH_CHECK("layernorm_bwd_finalize_kernel");
  return 0;
}

/* Backward of es3_win_attn_bias_bf16 on a map whose H and W are multiples of ws (the training graph pads the token map itself).
 * dS: [B * (H/ws) * (W/ws)][ldS] fp32, a row holding [heads][ws^2][ws^2] (ldS >= heads ws^4); the bias gradient is its column sum
 * (es3_colsum_f32). */
extern "C" int es3_win_attn_bias_bwd(const void* qkv, const void* dout, const float* bias, void* dqkv, void* dS, long long ldS, int B, int H,
                                     int W, int C, int num_heads, int ws, float scale, void* stream) {
  ES3_REQUIRE(C == num_heads * WB_HD, "es3_win_attn_bias_bwd: head_dim must be 32 (C=%d heads=%d)", C, num_heads);
  ES3_REQUIRE(ws > 0 && H % ws == 0 && W % ws == 0, "es3_win_attn_bias_bwd: the map (%d x %d) must be a multiple of the window %d", H, W, ws);
  const int N = ws * ws;
  ES3_REQUIRE(N <= 256, "es3_win_attn_bias_bwd: window %d too large (N <= 256)", ws);
  ES3_REQUIRE(ldS >= (long long)num_heads * N * N, "es3_win_attn_bias_bwd: ldS=%lld too small", ldS);
  WinBwdArgs a;
  a.qkv = (const bf16*)qkv; a.dout = (const bf16*)dout; a.bias = bias; a.dqkv = (bf16*)dqkv; a.dS = (float*)dS; a.ldS = ldS;
  a.H = H; a.W = W; a.C = C; a.ws = ws; a.nWx = W / ws; a.nWin = (H / ws) * (W / ws); a.N = N; a.scale = scale;
  const int smem = (4 * N * WB_LD + 3 * N) * (int)sizeof(float);
  // function attributes are per device: set unconditionally (a per-process cache would leave a second device at the 48 KB default)
  ES3_CHECK_CUDA(cudaFuncSetAttribute(win_attn_bias_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  win_attn_bias_bwd_kernel<<<dim3(B * a.nWin, num_heads), 256, smem, (cudaStream_t)stream>>>(a);
  ES3_LAUNCH_CHECK("win_attn_bias_bwd_kernel");
  return 0;
}

static int colsum_split(long long M, int L) {
  const long long slabs = (L + 127) / 128;
  long long ns = (148LL * 8 + slabs - 1) / slabs;     // ~8 CTAs per SM in flight
  if (ns > M) ns = M;
  if (ns > 1024) ns = 1024;
  if (ns < 1) ns = 1;
  return (int)ns;
}

extern "C" long long es3_colsum_f32_ws_floats(long long M, int L) { return (long long)colsum_split(M, L) * L; }

/* out[c] += sum over the M rows of src[r * ld + c], c < L (fp32, deterministic order). */
extern "C" int es3_colsum_f32(const float* src, long long ld, long long M, int L, float* ws, float* out, void* stream) {
  ES3_REQUIRE(M > 0 && L > 0 && ld >= L, "es3_colsum_f32: bad shape M=%lld L=%d ld=%lld", M, L, ld);
  const int ns = colsum_split(M, L);
  cudaStream_t st = (cudaStream_t)stream;
  colsum_f32_kernel<<<dim3(ceil_div(L, 128), ns), 128, 0, st>>>(src, ld, M, L, ns, ws);
  ES3_LAUNCH_CHECK("colsum_f32_kernel");
  colsum_f32_finalize_kernel<<<ceil_div(L, 128), 128, 0, st>>>(ws, ns, L, out);
  ES3_LAUNCH_CHECK("colsum_f32_finalize_kernel");
  return 0;
}
