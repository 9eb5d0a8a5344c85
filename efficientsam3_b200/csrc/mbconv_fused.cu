// Fused MBConv for the high-resolution, narrow stages of EfficientViT (reference
// efficientvit/nn/ops.py:315-367 MBConv, wrapped by ResidualBlock ops.py:740-770):
//
//   y = [x +] BN3(pw2( act(BN2(dw3x3_s( act(BN1(pw1(x))) ))) ))
//
// Unfused, the expanded tensor (4x the channels of x) is written and re-read twice through HBM -- at
// 512^2 / 256^2 / 128^2 that is ~85 % of the block's traffic (profiles/r1_kernel_table_c.md: the stage-1/2
// pointwise GEMMs + depthwise kernels were 4.5 of 11.9 ms).  Here one CTA owns a TH x 16 tile of output
// pixels: the haloed input tile is staged in shared memory once, and for each 64-channel chunk of the
// expanded tensor the CTA runs  expand (mma.sync bf16, halo included) -> depthwise 3x3 (mma.sync with
// diagonal B fragments, bf16 weights, fp32 accumulate) -> project (mma.sync, fp32 accumulators in registers).  The expanded tensor never leaves the SM;
// HBM traffic is x (+halo) in, y out.
//
// These layers are HBM-bound with K = 16..64 contractions; warp-level mma.sync is used on purpose -- a
// tcgen05 pipeline (TMEM alloc, UMMA descriptors over thread-written smem) buys nothing at this
// arithmetic intensity.  The tensor-bound GEMMs (head, stage 3/4, ViT) are in gemm_tc.cu.
#include "common.cuh"

namespace es3 {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// m16n8k8: with a diagonal 8x8 B fragment no structural zeros are multiplied (the k16 form wasted half of every MMA)
__device__ __forceinline__ void mma_bf16_1688(float* d, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int CIN, int MID, int COUT, int STRIDE>
struct MBCfg {
  static constexpr int TH = STRIDE == 1 ? 8 : 4, TW = 16;
  static constexpr int HH = (TH - 1) * STRIDE + 3, HWD = (TW - 1) * STRIDE + 3;
  static constexpr int P_IN = HH * HWD, P_IN_PAD = (P_IN + 15) / 16 * 16;
  static constexpr int P_OUT = TH * TW;
  static constexpr int MC = 64;  // mid-channel chunk
  static constexpr int RS_IN = CIN * 2 + 16, RS_MID = MC * 2 + 16, RS_W1 = CIN * 2 + 16, RS_W3 = MC * 2 + 16;
  static constexpr int RS_OUT = COUT * 2 + 16;
  static constexpr int SZ_IN = P_IN_PAD * RS_IN;
  static constexpr int SZ_MID = cmax(P_IN_PAD * RS_MID, P_OUT * RS_OUT);
  static constexpr int SZ_DW = P_OUT * RS_MID;
  static constexpr int SZ_W1 = MC * RS_W1, SZ_W3 = COUT * RS_W3;
  static constexpr int NBF = 9 * 4 * 2 * 32;             // depthwise B-fragment table (see kernel)
  static constexpr int NF = NBF + 3 * MC + 2 * COUT;    // bfrag | scale1 bias1 bias2 chunk | scale3 bias3
  static constexpr int SMEM = SZ_IN + SZ_MID + SZ_DW + SZ_W1 + SZ_W3 + NF * 4;
};

struct MBArgs {
  const bf16* x;      // [B,H,W,CIN]
  bf16* y;            // [B,Ho,Wo,COUT]
  const bf16* w1;     // [MID][CIN]
  const float* s1;    // [MID] scale after pw1 (BN1 or ones)
  const float* b1;    // [MID]
  const float* wdw;   // [9][MID] fp32, BN2 scale folded
  const float* b2;    // [MID]
  const bf16* w3;     // [COUT][MID]
  const float* s3;    // [COUT]
  const float* b3;    // [COUT]
  int H, W, Ho, Wo, tiles_x;
};

template <int CIN, int MID, int COUT, int STRIDE, bool RES, int ACT>
__global__ void __launch_bounds__(256, 2) mbconv_fused_kernel(const MBArgs a) {
  using C = MBCfg<CIN, MID, COUT, STRIDE>;
  static_assert(CIN % 16 == 0 && MID % 64 == 0 && COUT % 16 == 0, "channel multiples");
  static_assert(!RES || (CIN == COUT && STRIDE == 1), "residual needs matching shapes");
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* s_in = smem;
  uint8_t* s_mid = s_in + C::SZ_IN;
  uint8_t* s_dw = s_mid + C::SZ_MID;
  uint8_t* s_w1 = s_dw + C::SZ_DW;
  uint8_t* s_w3 = s_w1 + C::SZ_W1;
  float* s_f = reinterpret_cast<float*>(s_w3 + C::SZ_W3);
  // Depthwise on tensor cores: for a 16-channel group the 3x3 depthwise is 9 MMAs with DIAGONAL B
  // matrices, D[px][c] += A[px+tap][c] * w[tap][c].  A diagonal 16x16 B fragment has at most one non-zero
  // bf16 per lane, so the table stores the ready-made register: s_bfrag[tap][cg][nt][lane].
  uint32_t* s_bfrag = reinterpret_cast<uint32_t*>(s_f);   // [9][4][2][32]
  float* s_s1 = s_f + C::NBF;         // [64]
  float* s_b1 = s_s1 + C::MC;
  float* s_b2 = s_b1 + C::MC;
  float* s_s3 = s_b2 + C::MC;         // [COUT]
  float* s_b3 = s_s3 + COUT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int oy0 = (tile / a.tiles_x) * C::TH, ox0 = (tile % a.tiles_x) * C::TW;
  const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;

  // ---- stage the haloed input tile (zeros outside the image and in the row padding)
  {
    constexpr int NV = CIN / 8;
    const bf16* xb = a.x + (long long)b * a.H * a.W * CIN;
    for (int i = tid; i < C::P_IN_PAD * NV; i += 256) {
      const int v = i % NV, p = i / NV;
      const int py = p / C::HWD, px = p % C::HWD;
      const int iy = iy0 + py, ix = ix0 + px;
      const bool ok = (p < C::P_IN) && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
      cp16(s_in + p * C::RS_IN + v * 16, ok ? xb + ((long long)iy * a.W + ix) * CIN + v * 8 : xb, ok);
    }
    for (int i = tid; i < COUT; i += 256) { s_s3[i] = a.s3[i]; s_b3[i] = a.b3[i]; }
  }

  // project accumulators: warp -> (m-tile, n range)
  constexpr int M_TILES = C::P_OUT / 16;           // 8 (stride 1) or 4 (stride 2)
  constexpr int N_SPLIT = 8 / M_TILES;             // warps sharing an m-tile split the columns
  constexpr int NT = COUT / 8 / N_SPLIT;           // n-tiles per warp
  static_assert((COUT / 8) % N_SPLIT == 0 && NT % 2 == 0, "COUT split");
  const int pm = warp % M_TILES, pn0 = (warp / M_TILES) * NT;
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  const uint32_t u_in = static_cast<uint32_t>(__cvta_generic_to_shared(s_in));
  const uint32_t u_mid = static_cast<uint32_t>(__cvta_generic_to_shared(s_mid));
  const uint32_t u_dw = static_cast<uint32_t>(__cvta_generic_to_shared(s_dw));
  const uint32_t u_w1 = static_cast<uint32_t>(__cvta_generic_to_shared(s_w1));
  const uint32_t u_w3 = static_cast<uint32_t>(__cvta_generic_to_shared(s_w3));
  // ldmatrix lane addressing: A (x4): row = lane%16, k-half = lane/16;  B (x4 over two n-tiles):
  // n = lane%8 + (lane/16)*8, k-half = (lane/8)%2
  const int a_row = lane & 15, a_kh = lane >> 4;
  const int b_n = (lane & 7) + ((lane >> 4) << 3), b_kh = (lane >> 3) & 1;
  const int g = lane >> 2, t4 = lane & 3;

#pragma unroll 1
  for (int ch = 0; ch < MID / C::MC; ++ch) {
    // ---- chunk weights
    {
      constexpr int NV1 = CIN / 8;
      const bf16* w1c = a.w1 + (long long)ch * C::MC * CIN;
      for (int i = tid; i < C::MC * NV1; i += 256) {
        const int r = i / NV1, v = i % NV1;
        cp16(s_w1 + r * C::RS_W1 + v * 16, w1c + r * CIN + v * 8, true);
      }
      const bf16* w3c = a.w3 + ch * C::MC;
      for (int i = tid; i < COUT * 8; i += 256) {
        const int r = i >> 3, v = i & 7;
        cp16(s_w3 + r * C::RS_W3 + v * 16, w3c + (long long)r * MID + v * 8, true);
      }
      for (int i = tid; i < C::NBF; i += 256) {
        const int ln = i & 31, nt = (i >> 5) & 1, cg = (i >> 6) & 3, tap = i >> 8;
        const int gg = ln >> 2, tt = ln & 3;
        const float wv = a.wdw[tap * MID + ch * C::MC + cg * 16 + nt * 8 + gg];
        uint32_t r = 0u;
        if ((gg >> 1) == tt) {
          const uint32_t bits = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(wv));
          r = (gg & 1) ? (bits << 16) : bits;
        }
        s_bfrag[i] = r;
      }
      for (int i = tid; i < C::MC; i += 256) {
        s_s1[i] = a.s1[ch * C::MC + i];
        s_b1[i] = a.b1[ch * C::MC + i];
        s_b2[i] = a.b2[ch * C::MC + i];
      }
      cp_wait_all();
      __syncthreads();
    }

    // ---- expand: s_mid[p][0..63] = act(s1 * (x[p] . W1[c]) + b1), zero outside the image
#pragma unroll 1
    for (int mt = warp; mt < C::P_IN_PAD / 16; mt += 8) {
      float e[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { e[i][0] = e[i][1] = e[i][2] = e[i][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < CIN / 16; ++ks) {
        uint32_t af[4];
        ldsm_x4(u_in + (mt * 16 + a_row) * C::RS_IN + (ks * 16 + a_kh * 8) * 2, af[0], af[1], af[2], af[3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(u_w1 + (np * 16 + b_n) * C::RS_W1 + (ks * 16 + b_kh * 8) * 2, b0, b1, b2, b3);
          mma_bf16_16816(e[2 * np], af, b0, b1);
          mma_bf16_16816(e[2 * np + 1], af, b2, b3);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int p = mt * 16 + g + half * 8;
        const int py = p / C::HWD, px = p - py * C::HWD;
        const int iy = iy0 + py, ix = ix0 + px;
        const bool in = (p < C::P_IN) && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int c = nt * 8 + t4 * 2;
          float v0 = es3_act_t<ACT>(fmaf(e[nt][half * 2 + 0], s_s1[c], s_b1[c]));
          float v1 = es3_act_t<ACT>(fmaf(e[nt][half * 2 + 1], s_s1[c + 1], s_b1[c + 1]));
          if (!in) { v0 = 0.f; v1 = 0.f; }
          *reinterpret_cast<uint32_t*>(s_mid + p * C::RS_MID + c * 2) = pack_bf16x2(v0, v1);
        }
      }
    }
    __syncthreads();

    // ---- depthwise 3x3 (stride STRIDE) on the chunk, on tensor cores (diagonal-B MMAs).
    // warp -> channel group cg = warp % 4 (16 channels), output rows mt = warp / 4, +2, ... (one m-tile
    // = the 16 output pixels of a tile row)
    {
      const int cg = warp & 3;
      const uint32_t* bf = s_bfrag + cg * 64 + lane;
      constexpr int MT_PER_WARP = C::TH / 2;   // m-tiles (tile rows) per warp: rows warp/4, +2, +4, ...
      // taps outer / m-tiles inner: MT_PER_WARP x 2 independent accumulator chains keep the tensor pipe busy
      // (a single m-tile is a 9-deep dependent HMMA chain; profiles/r1_mbconv_ncu.md)
      float d[MT_PER_WARP][2][4];
#pragma unroll
      for (int m = 0; m < MT_PER_WARP; ++m)
#pragma unroll
        for (int i = 0; i < 2; ++i) { d[m][i][0] = d[m][i][1] = d[m][i][2] = d[m][i][3] = 0.f; }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t b_lo = bf[(ky * 3 + kx) * 256], b_hi = bf[(ky * 3 + kx) * 256 + 32];
#pragma unroll
          for (int m = 0; m < MT_PER_WARP; ++m) {
            const int mt = (warp >> 2) + 2 * m;
            uint32_t af[4];
            ldsm_x4(u_mid + ((mt * STRIDE + ky) * C::HWD + a_row * STRIDE + kx) * C::RS_MID + (cg * 16 + a_kh * 8) * 2,
                    af[0], af[1], af[2], af[3]);
            mma_bf16_1688(d[m][0], af[0], af[1], b_lo);
            mma_bf16_1688(d[m][1], af[2], af[3], b_hi);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < MT_PER_WARP; ++m) {
        const int mt = (warp >> 2) + 2 * m;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int p = mt * C::TW + g + half * 8;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int c = cg * 16 + nt * 8 + t4 * 2;
            const float v0 = es3_act_t<ACT>(d[m][nt][half * 2 + 0] + s_b2[c]);
            const float v1 = es3_act_t<ACT>(d[m][nt][half * 2 + 1] + s_b2[c + 1]);
            *reinterpret_cast<uint32_t*>(s_dw + p * C::RS_MID + c * 2) = pack_bf16x2(v0, v1);
          }
        }
      }
    }
    __syncthreads();

    // ---- project: acc[pm rows][pn0..] += s_dw[rows][0..63] . W3[n][chunk]
#pragma unroll
    for (int ks = 0; ks < C::MC / 16; ++ks) {
      uint32_t af[4];
      ldsm_x4(u_dw + (pm * 16 + a_row) * C::RS_MID + (ks * 16 + a_kh * 8) * 2, af[0], af[1], af[2], af[3]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(u_w3 + ((pn0 + np * 2) * 8 + b_n) * C::RS_W3 + (ks * 16 + b_kh * 8) * 2, b0, b1, b2, b3);
        mma_bf16_16816(acc[2 * np], af, b0, b1);
        mma_bf16_16816(acc[2 * np + 1], af, b2, b3);
      }
    }
    __syncthreads();  // s_mid / s_dw / weights are overwritten by the next chunk (or the epilogue)
  }

  // ---- epilogue: BN3 (+ residual) -> bf16 tile in smem -> coalesced 16-byte stores
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = pm * 16 + g + half * 8;  // output pixel inside the tile
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c = (pn0 + nt) * 8 + t4 * 2;
      float v0 = fmaf(acc[nt][half * 2 + 0], s_s3[c], s_b3[c]);
      float v1 = fmaf(acc[nt][half * 2 + 1], s_s3[c + 1], s_b3[c + 1]);
      if (RES) {
        const int sy = r / C::TW, sx = r % C::TW;
        const float2 xr = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(s_in + ((sy + 1) * C::HWD + sx + 1) * C::RS_IN + c * 2));
        v0 += xr.x; v1 += xr.y;
      }
      *reinterpret_cast<uint32_t*>(s_mid + r * C::RS_OUT + c * 2) = pack_bf16x2(v0, v1);
    }
  }
  __syncthreads();
  {
    constexpr int NVO = COUT / 8;
    bf16* yb = a.y + (long long)b * a.Ho * a.Wo * COUT;
    for (int i = tid; i < C::P_OUT * NVO; i += 256) {
      const int p = i / NVO, v = i % NVO;
      const int oy = oy0 + p / C::TW, ox = ox0 + p % C::TW;
      if (oy < a.Ho && ox < a.Wo)
        *reinterpret_cast<uint4*>(yb + ((long long)oy * a.Wo + ox) * COUT + v * 8) =
            *reinterpret_cast<const uint4*>(s_mid + p * C::RS_OUT + v * 16);
    }
  }
}

template <int CIN, int MID, int COUT, int STRIDE, bool RES>
static int launch_mbconv(const MBArgs& a, int B, int act, cudaStream_t st) {
  using C = MBCfg<CIN, MID, COUT, STRIDE>;
  ES3_REQUIRE(act == ACT_HSWISH, "es3_mbconv_fused_bf16: only hardswish is instantiated (got act=%d)", act);
  auto kern = mbconv_fused_kernel<CIN, MID, COUT, STRIDE, RES, ACT_HSWISH>;
  static bool configured = false;
  if (!configured) {
    ES3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  MBArgs aa = a;
  aa.tiles_x = ceil_div(a.Wo, C::TW);
  dim3 grid(aa.tiles_x * ceil_div(a.Ho, C::TH), B);
  kern<<<grid, 256, C::SMEM, st>>>(aa);
  ES3_LAUNCH_CHECK("mbconv_fused_kernel");
  return 0;
}

}  // namespace es3

using namespace es3;

// Returns 0 on success, -1 if this (CIN, MID, COUT, stride, residual) shape is not instantiated (the caller
// then uses the unfused gemm_tc + dwconv path), > 0 on error.
extern "C" int es3_mbconv_fused_bf16(const void* x, void* y, const void* w1, const float* s1, const float* b1,
                                     const float* wdw, const float* b2, const void* w3, const float* s3,
                                     const float* b3, int B, int H, int W, int Cin, int Mid, int Cout, int stride,
                                     int residual, int act, void* stream) {
  MBArgs a;
  a.x = (const bf16*)x; a.y = (bf16*)y; a.w1 = (const bf16*)w1; a.s1 = s1; a.b1 = b1; a.wdw = wdw; a.b2 = b2;
  a.w3 = (const bf16*)w3; a.s3 = s3; a.b3 = b3;
  a.H = H; a.W = W; a.Ho = (H - 1) / stride + 1; a.Wo = (W - 1) / stride + 1; a.tiles_x = 0;
  cudaStream_t st = (cudaStream_t)stream;
#define ES3_MB(CI, MI, CO, S, R) \
  if (Cin == CI && Mid == MI && Cout == CO && stride == S && (residual != 0) == R) return launch_mbconv<CI, MI, CO, S, R>(a, B, act, st);
  ES3_MB(16, 64, 32, 2, false)    // efficientvit_b1 stages.0.op_list.0
  ES3_MB(32, 128, 32, 1, true)    // stages.0.op_list.1
  ES3_MB(32, 128, 64, 2, false)   // stages.1.op_list.0
  ES3_MB(64, 256, 64, 1, true)    // stages.1.op_list.1-2
  ES3_MB(64, 256, 128, 2, false)  // stages.2.op_list.0
#undef ES3_MB
  return -1;
}
