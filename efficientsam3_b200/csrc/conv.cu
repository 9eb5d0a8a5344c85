// Bandwidth-bound convolution kernels (NHWC bf16 activations, fp32 math):
//   * stem 3x3 stride-2 conv reading the NCHW fp32 image (EfficientViT input_stem op 0:
//     efficientvit/backbone.py:49-57; RepViT / TinyViT patch-embed first conv)
//   * depthwise k x k conv (+folded BN / bias, + activation)  (efficientvit/nn/ops.py:273-367 depth_conv)
//   * fused "DSConv residual" block of the EfficientViT stem (dw3x3+BN+act -> pw+BN, + x)
//   * bilinear resize NHWC bf16 -> NCHW fp32 (stage1/model.py:204-210 F.interpolate, align_corners=False)
// None of these is a dense contraction worth tensor cores; they are sized for HBM/L2 streaming:
// 16-byte vector accesses, one thread per (pixel, 8-channel group).
#include "common.cuh"

namespace es3 {

// ------------------------------------------------------------------------------------------ stem
// x: [B,3,H,W] fp32 NCHW.  w: [27][COUT] fp32 (tap-major: (ci*9 + ky*3 + kx), BN scale pre-folded).
// out: [B,Ho,Wo,COUT] bf16.  One thread per output pixel, COUT accumulators.
template <int COUT, int ACT>
__global__ void stem_conv3x3_s2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                       const float* __restrict__ bias, bf16* __restrict__ out, int B, int H, int W,
                                       int Ho, int Wo) {
  __shared__ float sw[27 * COUT];
  __shared__ float sb[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long total = (long long)B * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int ox = (int)(p % Wo);
  const int oy = (int)((p / Wo) % Ho);
  const int b = (int)(p / ((long long)Wo * Ho));
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = sb[c];
  const float* xb = x + (long long)b * 3 * H * W;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const float v = __ldg(xb + ((long long)ci * H + iy) * W + ix);
        const float* wp = sw + (ci * 9 + ky * 3 + kx) * COUT;
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v, wp[c], acc[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = es3_act_t<ACT>(acc[c]);
  uint4* op = reinterpret_cast<uint4*>(out + p * COUT);
#pragma unroll
  for (int j = 0; j < COUT / 8; ++j) op[j] = pack8(acc + 8 * j);
}

// ------------------------------------------------------------------------------------- depthwise
// x: [B,H,W,C] bf16 (pixel stride ldx elements), w: [KS*KS][C] fp32 (tap-major, BN scale pre-folded),
// bias: [C] fp32 or null, out: [B,Ho,Wo,C] bf16 (pixel stride ldo).  pad = KS/2.
template <int KS, int STRIDE, int ACT>
__global__ void dwconv_kernel(const bf16* __restrict__ x, long long ldx, const float* __restrict__ w,
                              const float* __restrict__ bias, bf16* __restrict__ out, long long ldo, int B, int H,
                              int W, int C, int Ho, int Wo) {
  const int cg = C >> 3;  // 8-channel groups
  const long long total = (long long)B * Ho * Wo * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % cg);
  const long long p = idx / cg;
  const int ox = (int)(p % Wo);
  const int oy = (int)((p / Wo) % Ho);
  const int b = (int)(p / ((long long)Wo * Ho));
  const int c0 = g * 8;
  constexpr int PAD = KS / 2;
  float acc[8];
  if (bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  }
#pragma unroll
  for (int ky = 0; ky < KS; ++ky) {
    const int iy = oy * STRIDE - PAD + ky;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const int ix = ox * STRIDE - PAD + kx;
      if (ix < 0 || ix >= W) continue;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * ldx + c0));
      float f[8];
      unpack8(u, f);
      const float* wp = w + (ky * KS + kx) * C + c0;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]);
      acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
      acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]);
      acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = es3_act_t<ACT>(acc[e]);
  *reinterpret_cast<uint4*>(out + p * ldo + c0) = pack8(acc);
}

// ------------------------------------------------------------------------------------- DSConv res
// EfficientViT stem block (backbone.py:58-67 with expand_ratio=1 -> DSConv, ops.py:273-312):
//   y = x + BN2(pw(act(BN1(dw3x3(x)))))       C == 16 (b1) / 8 (b0) / 24 (b2) / 32 (b3)
// One thread per pixel, all C channels in registers; pw weights [C][C] in shared memory.
template <int C, int ACT>
__global__ void dsconv_res_kernel(const bf16* __restrict__ x, const float* __restrict__ wdw /*[9][C]*/,
                                  const float* __restrict__ bdw, const float* __restrict__ wpw /*[Cout][Cin]*/,
                                  const float* __restrict__ bpw, bf16* __restrict__ out, int B, int H, int W) {
  __shared__ float s_wdw[9 * C];
  __shared__ float s_wpw[C * C];
  __shared__ float s_bdw[C];
  __shared__ float s_bpw[C];
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_wdw[i] = wdw[i];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_wpw[i] = wpw[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_bdw[i] = bdw ? bdw[i] : 0.f; s_bpw[i] = bpw ? bpw[i] : 0.f; }
  __syncthreads();
  const long long total = (long long)B * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int ox = (int)(p % W);
  const int oy = (int)((p / W) % H);
  const int b = (int)(p / ((long long)W * H));
  float mid[C];
  float ctr[C];
#pragma unroll
  for (int c = 0; c < C; ++c) mid[c] = s_bdw[c];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy - 1 + ky;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox - 1 + kx;
      const bool in = (iy >= 0 && iy < H && ix >= 0 && ix < W);
      const uint4* xp = reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * C);
#pragma unroll
      for (int j = 0; j < C / 8; ++j) {
        float f[8];
        if (in) unpack8(__ldg(xp + j), f);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          mid[j * 8 + e] = fmaf(f[e], s_wdw[(ky * 3 + kx) * C + j * 8 + e], mid[j * 8 + e]);
          if (ky == 1 && kx == 1) ctr[j * 8 + e] = f[e];
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    // the reference materialises the dw output in the activation dtype; round like the unfused path would
    mid[c] = __bfloat162float(__float2bfloat16(es3_act_t<ACT>(mid[c])));
  }
  float o[C];
#pragma unroll
  for (int n = 0; n < C; ++n) {
    float a = s_bpw[n];
#pragma unroll
    for (int c = 0; c < C; ++c) a = fmaf(mid[c], s_wpw[n * C + c], a);
    o[n] = a + ctr[n];
  }
  uint4* op = reinterpret_cast<uint4*>(out + p * C);
#pragma unroll
  for (int j = 0; j < C / 8; ++j) op[j] = pack8(o + 8 * j);
}

// ------------------------------------------------------------------------------------- bilinear
// in: [B,Hi,Wi,C] bf16 NHWC -> out: [B,C,Ho,Wo] fp32 NCHW, PyTorch bilinear, align_corners=False,
// antialias=False (aten upsample_bilinear2d: src = (dst+0.5)*in/out - 0.5 clamped at 0).
// grid (Ho, C/64, B), block 256: stage the two source rows of a 64-channel slab in smem, then write
// rows of Wo contiguous floats.
// One CTA = one image x 16 channels: the whole Hi x Wi x 16 source slab is staged in smem (fp32), then each
// output plane [Ho][Wo] is written front to back -- 16 fully sequential Ho*Wo*4-byte streams per CTA (the first
// version wrote one 256-byte row per plane per CTA and reached only 1.4-1.6 TB/s, profiles/r1_kernel_table_h.md).
constexpr int BL_CB = 16, BL_PS = 20;   // pixel stride in floats: 16 channels + 4 pad -> 16-byte aligned, conflict-free LDS.128
__global__ void __launch_bounds__(256, 2) bilinear_nhwc_to_nchw_kernel(const bf16* __restrict__ in, float* __restrict__ out,
                                                                    int Hi, int Wi, int C, int Ho, int Wo, float sy,
                                                                    float sx) {
  extern __shared__ __align__(16) float ssrc[];  // [Hi*Wi][BL_PS]
  const int npix = Hi * Wi;
  const int c0 = blockIdx.x * BL_CB, b = blockIdx.y;
  for (int i = threadIdx.x; i < npix * 2; i += 256) {
    const int px = i >> 1, v = i & 1;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((long long)b * npix + px) * C + c0 + v * 8));
    float f[8];
    unpack8(u, f);
    float4* d = reinterpret_cast<float4*>(ssrc + (long long)px * BL_PS + v * 8);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  // thread -> one output pixel column-quad q (fixed: its x taps / weights live in registers) and rows oy = r, r + RS, ...;
  // every tap is read once as 4 x LDS.128 (16 channels) and feeds 16 planes: ~1 shared load and 4 FMAs per output value
  // (the per-plane loop this replaces spent ~7 LDS per output; profiles/r1_kernel_table_i.md: 2.1 TB/s).
  const int nq = (Wo + 3) >> 2;
  int QT = 1;
  while (QT < nq && QT < 256) QT <<= 1;
  const int RS = 256 / QT;
  const int r = threadIdx.x / QT;
  float* obase = out + ((long long)b * C + c0) * Ho * Wo;
  const long long plane = (long long)Ho * Wo;
  for (int q = threadIdx.x % QT; q < nq; q += QT) {
    int x0[4], x1[4];
    float lx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ox = min(q * 4 + e, Wo - 1);
      float f = (ox + 0.5f) * sx - 0.5f;
      if (f < 0.f) f = 0.f;
      const int i0 = min((int)f, Wi - 1);
      x0[e] = i0 * BL_PS; x1[e] = min(i0 + 1, Wi - 1) * BL_PS; lx[e] = f - (float)i0;
    }
    for (int oy = r; oy < Ho; oy += RS) {
      float fy = (oy + 0.5f) * sy - 0.5f;
      if (fy < 0.f) fy = 0.f;
      const int y0 = min((int)fy, Hi - 1), y1 = min(y0 + 1, Hi - 1);
      const float ly = fy - (float)y0, hy = 1.f - ly;
      const float* r0 = ssrc + (long long)y0 * Wi * BL_PS;
      const float* r1 = ssrc + (long long)y1 * Wi * BL_PS;
      float* orow = obase + (long long)oy * Wo + q * 4;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {             // two passes of 8 channels keep the accumulators at 32 registers
        float v[8][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float hx = 1.f - lx[e];
#pragma unroll
          for (int j2 = 0; j2 < 2; ++j2) {
            const int j = jh * 2 + j2;
            const float4 a = *reinterpret_cast<const float4*>(r0 + x0[e] + 4 * j);
            const float4 bb = *reinterpret_cast<const float4*>(r0 + x1[e] + 4 * j);
            const float4 c = *reinterpret_cast<const float4*>(r1 + x0[e] + 4 * j);
            const float4 d = *reinterpret_cast<const float4*>(r1 + x1[e] + 4 * j);
            // same association as aten's upsample_bilinear2d: hy * (hx a + lx b) + ly * (hx c + lx d)
            v[4 * j2 + 0][e] = hy * (hx * a.x + lx[e] * bb.x) + ly * (hx * c.x + lx[e] * d.x);
            v[4 * j2 + 1][e] = hy * (hx * a.y + lx[e] * bb.y) + ly * (hx * c.y + lx[e] * d.y);
            v[4 * j2 + 2][e] = hy * (hx * a.z + lx[e] * bb.z) + ly * (hx * c.z + lx[e] * d.z);
            v[4 * j2 + 3][e] = hy * (hx * a.w + lx[e] * bb.w) + ly * (hx * c.w + lx[e] * d.w);
          }
        }
        float* o8 = orow + (long long)(jh * 8) * plane;
        if ((Wo & 3) == 0) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(o8 + c * plane) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (q * 4 + e < Wo) o8[c * plane + e] = v[c][e];
        }
      }
    }
  }
}

// NHWC bf16 -> NCHW fp32 without resampling (used when the head output already has the embed size,
// and to hand intermediate feature maps back to PyTorch callers).
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ in, float* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < HW && c < C) ? __bfloat162float(in[((long long)b * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (p < HW && c < C) out[((long long)b * C + c) * HW + p] = tile[tx][i];
  }
}

// NCHW fp32 -> NHWC bf16 (entry conversion for callers that hand the native layer fp32 feature maps).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, bf16* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (p < HW && c < C) ? in[((long long)b * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) out[((long long)b * HW + p) * C + c] = __float2bfloat16(tile[tx][i]);
  }
}

}  // namespace es3

using namespace es3;

extern "C" int es3_stem_conv3x3_s2(const float* x, const float* w, const float* bias, void* out, int B, int H, int W,
                                   int Cout, int act, void* stream) {
  ES3_REQUIRE(B > 0 && H > 0 && W > 0, "es3_stem_conv3x3_s2: bad shape");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)B * Ho * Wo;
  const int threads = 128;
  const unsigned blocks = (unsigned)ceil_div(total, threads);
  cudaStream_t st = (cudaStream_t)stream;
  ES3_DISPATCH_ACT(act, A, {
    switch (Cout) {
      case 8: stem_conv3x3_s2_kernel<8, A><<<blocks, threads, 0, st>>>(x, w, bias, (bf16*)out, B, H, W, Ho, Wo); break;
      case 16: stem_conv3x3_s2_kernel<16, A><<<blocks, threads, 0, st>>>(x, w, bias, (bf16*)out, B, H, W, Ho, Wo); break;
      case 24: stem_conv3x3_s2_kernel<24, A><<<blocks, threads, 0, st>>>(x, w, bias, (bf16*)out, B, H, W, Ho, Wo); break;
      case 32: stem_conv3x3_s2_kernel<32, A><<<blocks, threads, 0, st>>>(x, w, bias, (bf16*)out, B, H, W, Ho, Wo); break;
      case 48: stem_conv3x3_s2_kernel<48, A><<<blocks, threads, 0, st>>>(x, w, bias, (bf16*)out, B, H, W, Ho, Wo); break;
      default: ES3_REQUIRE(false, "es3_stem_conv3x3_s2: unsupported Cout=%d (8/16/24/32/48)", Cout);
    }
  })
  ES3_LAUNCH_CHECK("stem_conv3x3_s2_kernel");
  return 0;
}

extern "C" int es3_dwconv_bf16(const void* x, long long ldx, const float* w, const float* bias, void* out,
                               long long ldo, int B, int H, int W, int C, int ks, int stride, int act, void* stream) {
  ES3_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "es3_dwconv_bf16: C/ldx/ldo must be multiples of 8 (C=%d)", C);
  ES3_REQUIRE((ks == 3 || ks == 5) && (stride == 1 || stride == 2), "es3_dwconv_bf16: unsupported ks=%d stride=%d", ks, stride);
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  const long long total = (long long)B * Ho * Wo * (C / 8);
  const int threads = 256;
  const unsigned blocks = (unsigned)ceil_div(total, threads);
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* xi = (const bf16*)x;
  bf16* o = (bf16*)out;
  ES3_DISPATCH_ACT(act, A, {
    if (ks == 3 && stride == 1) dwconv_kernel<3, 1, A><<<blocks, threads, 0, st>>>(xi, ldx, w, bias, o, ldo, B, H, W, C, Ho, Wo);
    else if (ks == 3 && stride == 2) dwconv_kernel<3, 2, A><<<blocks, threads, 0, st>>>(xi, ldx, w, bias, o, ldo, B, H, W, C, Ho, Wo);
    else if (ks == 5 && stride == 1) dwconv_kernel<5, 1, A><<<blocks, threads, 0, st>>>(xi, ldx, w, bias, o, ldo, B, H, W, C, Ho, Wo);
    else dwconv_kernel<5, 2, A><<<blocks, threads, 0, st>>>(xi, ldx, w, bias, o, ldo, B, H, W, C, Ho, Wo);
  })
  ES3_LAUNCH_CHECK("dwconv_kernel");
  return 0;
}

extern "C" int es3_dsconv_res_bf16(const void* x, const float* wdw, const float* bdw, const float* wpw,
                                   const float* bpw, void* out, int B, int H, int W, int C, int act, void* stream) {
  const long long total = (long long)B * H * W;
  const int threads = 128;
  const unsigned blocks = (unsigned)ceil_div(total, threads);
  cudaStream_t st = (cudaStream_t)stream;
  ES3_DISPATCH_ACT(act, A, {
    switch (C) {
      case 8: dsconv_res_kernel<8, A><<<blocks, threads, 0, st>>>((const bf16*)x, wdw, bdw, wpw, bpw, (bf16*)out, B, H, W); break;
      case 16: dsconv_res_kernel<16, A><<<blocks, threads, 0, st>>>((const bf16*)x, wdw, bdw, wpw, bpw, (bf16*)out, B, H, W); break;
      case 24: dsconv_res_kernel<24, A><<<blocks, threads, 0, st>>>((const bf16*)x, wdw, bdw, wpw, bpw, (bf16*)out, B, H, W); break;
      case 32: dsconv_res_kernel<32, A><<<blocks, threads, 0, st>>>((const bf16*)x, wdw, bdw, wpw, bpw, (bf16*)out, B, H, W); break;
      default: ES3_REQUIRE(false, "es3_dsconv_res_bf16: unsupported C=%d (8/16/24/32)", C);
    }
  })
  ES3_LAUNCH_CHECK("dsconv_res_kernel");
  return 0;
}

extern "C" int es3_bilinear_nhwc_to_nchw(const void* in, float* out, int B, int Hi, int Wi, int C, int Ho, int Wo,
                                         void* stream) {
  ES3_REQUIRE(C % BL_CB == 0, "es3_bilinear_nhwc_to_nchw: C=%d must be a multiple of %d", C, BL_CB);
  const size_t smem = (size_t)Hi * Wi * BL_PS * sizeof(float);
  ES3_REQUIRE(smem <= 200 * 1024, "es3_bilinear_nhwc_to_nchw: %dx%d source too large for the smem slab", Hi, Wi);
  static bool configured = false;
  if (!configured) {
    ES3_CHECK_CUDA(cudaFuncSetAttribute(bilinear_nhwc_to_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid(C / BL_CB, B);
  bilinear_nhwc_to_nchw_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)in, out, Hi, Wi, C, Ho, Wo,
                                                                           (float)Hi / (float)Ho, (float)Wi / (float)Wo);
  ES3_LAUNCH_CHECK("bilinear_nhwc_to_nchw_kernel");
  return 0;
}

extern "C" int es3_nhwc_to_nchw_f32(const void* in, float* out, int B, int HW, int C, void* stream) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B);
  nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, out, HW, C);
  ES3_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}

extern "C" int es3_nchw_f32_to_nhwc(const float* in, void* out, int B, int HW, int C, void* stream) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (bf16*)out, HW, C);
  ES3_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

// MaxPool 2x2 stride 2 on NHWC bf16 (FPN 0.5x level, necks.py:64-69).  Thread = (output pixel, 8 channels).
namespace es3 {
__global__ void maxpool2x2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, cg = C / 8;
  const long long total = (long long)B * Ho * Wo * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % cg);
  const long long p = idx / cg;
  const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho);
  const int b = (int)(p / ((long long)Wo * Ho));
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + oy * 2 + dy) * W + ox * 2 + dx) * C + g * 8)), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], f[e]);
    }
  *reinterpret_cast<uint4*>(out + p * C + g * 8) = pack8(m);
}
}  // namespace es3

extern "C" int es3_maxpool2x2_bf16(const void* x, void* out, int B, int H, int W, int C, void* stream) {
  ES3_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2, "es3_maxpool2x2_bf16: bad shape");
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  es3::maxpool2x2_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, B, H, W, C);
  ES3_LAUNCH_CHECK("maxpool2x2_kernel");
  return 0;
}
