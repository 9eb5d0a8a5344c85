// Small kernels of the SAM-style prompt encoder / TwoWayTransformer / MaskDecoder path
// (sam3/sam3/sam/{prompt_encoder,transformer,mask_decoder}.py, glue sam3_tracker_base.py:220-389).
// The dense contractions over the 5184 image tokens run on gemm_tc.cu; everything here is either tiny
// (8 prompt/output tokens per image: latency-bound) or a streaming pass over the image tokens / masks.
// Token-side math is fp32 end to end (it feeds the mask logits through the hypernetwork).
#include "common.cuh"

namespace es3 {

// ------------------------------------------------------------------------------------ positional enc.
// PositionEmbeddingRandom (prompt_encoder.py:200-243): c = 2*coords01 - 1; c @ G[2,F]; *2pi; [sin | cos].
__device__ __forceinline__ void pe_pair(const float* __restrict__ gauss, int F, float x01, float y01, int f, float* s,
                                        float* c) {
  const float cx = 2.f * x01 - 1.f, cy = 2.f * y01 - 1.f;
  const float a = 6.283185307179586f * (cx * gauss[f] + cy * gauss[F + f]);
  sincosf(a, s, c);
}

// Dense PE for an h x w grid, token-major: out [h*w][2F] fp32 (get_dense_pe, prompt_encoder.py:61-69).
__global__ void dense_pe_kernel(const float* __restrict__ gauss, int F, int h, int w, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= h * w * F) return;
  const int f = idx % F, p = idx / F;
  const int y = p / w, x = p % w;
  float s, c;
  pe_pair(gauss, F, (x + 0.5f) / w, (y + 0.5f) / h, f, &s, &c);
  out[(long long)p * 2 * F + f] = s;
  out[(long long)p * 2 * F + F + f] = c;
}

// Sparse point embeddings with the trailing padding point (_embed_points with pad=True, :71-117).
// coords [B,P,2] (x,y) pixels, labels [B,P] int32.  out [B,P+1,2F] fp32.
// tables: not_a_point [2F], point_emb [4][2F].
__global__ void point_embed_kernel(const float* __restrict__ coords, const int* __restrict__ labels,
                                   const float* __restrict__ gauss, const float* __restrict__ not_a_point,
                                   const float* __restrict__ point_emb, int F, int P, float img_w, float img_h,
                                   float* __restrict__ out) {
  const int bp = blockIdx.x;  // b*(P+1) + p
  const int b = bp / (P + 1), p = bp % (P + 1);
  float x = 0.f, y = 0.f;
  int lab = -1;
  if (p < P) {
    x = coords[((long long)b * P + p) * 2 + 0];
    y = coords[((long long)b * P + p) * 2 + 1];
    lab = labels[(long long)b * P + p];
  }
  x = (x + 0.5f) / img_w;
  y = (y + 0.5f) / img_h;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s, c;
    pe_pair(gauss, F, x, y, f, &s, &c);
    if (lab == -1) { s = not_a_point[f]; c = not_a_point[F + f]; }
    else if (lab >= 0 && lab < 4) { s += point_emb[lab * 2 * F + f]; c += point_emb[lab * 2 * F + F + f]; }
    out[(long long)bp * 2 * F + f] = s;
    out[(long long)bp * 2 * F + F + f] = c;
  }
}

// ------------------------------------------------------------------------------------ elementwise
// y[m][c] = x[m][c] + add[(m % R)][c]; writes bf16 and/or fp32.  C % 4 == 0.
__global__ void add_rows_kernel(const float* __restrict__ x, const float* __restrict__ add, long long M, int C, int R,
                                bf16* __restrict__ y_bf16, float* __restrict__ y_f32) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = C >> 2;
  if (i4 >= M * c4) return;
  const long long m = i4 / c4;
  const int c = (int)(i4 % c4);
  float4 v = reinterpret_cast<const float4*>(x)[i4];
  if (add != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(add) + (m % R) * c4 + c);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (y_f32) reinterpret_cast<float4*>(y_f32)[i4] = v;
  if (y_bf16) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(y_bf16)[i4] = u;
  }
}

// [B,C,HW] fp32 -> [B,HW,C] fp32 (+ optional per-channel vector add, e.g. no_mask_embed), also bf16 copy.
__global__ void nchw_to_tokens_kernel(const float* __restrict__ in, const float* __restrict__ addc, float* __restrict__ out_f32,
                                      bf16* __restrict__ out_bf16, int HW, int C) {
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
    if (p < HW && c < C) {
      const float v = tile[tx][i] + (addc ? addc[c] : 0.f);
      if (out_f32) out_f32[((long long)b * HW + p) * C + c] = v;
      if (out_bf16) out_bf16[((long long)b * HW + p) * C + c] = __float2bfloat16(v);
    }
  }
}

// ------------------------------------------------------------------------------------ attention
__device__ __forceinline__ float ld_f(const float* p) { return *p; }
__device__ __forceinline__ float ld_f(const bf16* p) { return __bfloat162float(*p); }

// Few queries (<= 32 per image), many keys: one warp per (query, head, image); lanes stride the keys with a
// private online-softmax state that is merged by shuffles.  q [B,Tq,ldq] fp32, k/v [B,Tk,ld] (bf16|fp32),
// out [B,Tq,ldo] fp32, heads laid out [h*HD, (h+1)*HD).   Attention(transformer.py:185-264) core.
template <int HD, typename TKV>
__global__ void attn_few_queries_kernel(const float* __restrict__ q, long long ldq, const TKV* __restrict__ k,
                                        const TKV* __restrict__ v, long long ldkv, float* __restrict__ out, long long ldo,
                                        int Tq, int Tk, float scale) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int h = blockIdx.y, b = blockIdx.z;
  if (qi >= Tq) return;
  float qr[HD];
  const float* qp = q + ((long long)b * Tq + qi) * ldq + h * HD;
#pragma unroll
  for (int d = 0; d < HD; ++d) qr[d] = qp[d] * scale;
  float m = -INFINITY, l = 0.f, acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  const TKV* kb = k + (long long)b * Tk * ldkv + h * HD;
  const TKV* vb = v + (long long)b * Tk * ldkv + h * HD;
  for (int j = lane; j < Tk; j += 32) {
    const TKV* kp = kb + (long long)j * ldkv;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(qr[d], ld_f(kp + d), s);
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), p = __expf(s - mn);
    l = l * corr + p;
    const TKV* vp = vb + (long long)j * ldkv;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, ld_f(vp + d), acc[d] * corr);
    m = mn;
  }
  // merge the 32 partial states
  float mw = warp_max(m);
  const float f = (m == -INFINITY) ? 0.f : __expf(m - mw);
  l = warp_sum(l * f);
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = warp_sum(acc[d] * f);
  if (lane == 0) {
    float* op = out + ((long long)b * Tq + qi) * ldo + h * HD;
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) op[d] = acc[d] * inv;
  }
}

// Many queries (image tokens), few keys (<= 16 tokens): one thread per (query, head).
// q [B,Nq,ldq] bf16, k/v [B,Tk,ldkv] fp32, out [B,Nq,ldo] bf16.  cross_attn_image_to_token core.
template <int HD>
__global__ void attn_few_keys_kernel(const bf16* __restrict__ q, long long ldq, const float* __restrict__ k,
                                     const float* __restrict__ v, long long ldkv, bf16* __restrict__ out, long long ldo,
                                     int Nq, int Tk, int H, float scale) {
  extern __shared__ float skv[];  // [2][Tk][H*HD]
  const int b = blockIdx.y;
  const int D = H * HD;
  for (int i = threadIdx.x; i < Tk * D; i += blockDim.x) {
    const int t = i / D, c = i % D;
    skv[i] = k[((long long)b * Tk + t) * ldkv + c];
    skv[Tk * D + i] = v[((long long)b * Tk + t) * ldkv + c];
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Nq * H) return;
  const int h = idx % H, n = idx / H;
  float qr[HD];
  const bf16* qp = q + ((long long)b * Nq + n) * ldq + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 8) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(qp + d), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) qr[d + e] = f[e] * scale;
  }
  float mx = -INFINITY, l = 0.f, o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  for (int t = 0; t < Tk; ++t) {   // online softmax over the few keys
    const float* kp = skv + t * D + h * HD;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) a = fmaf(qr[d], kp[d], a);
    const float mn = fmaxf(mx, a);
    const float corr = __expf(mx - mn), p = __expf(a - mn);
    l = l * corr + p;
    const float* vp = skv + Tk * D + t * D + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = fmaf(p, vp[d], o[d] * corr);
    mx = mn;
  }
  const float inv = 1.f / l;
  bf16* op = out + ((long long)b * Nq + n) * ldo + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 8) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = o[d + e] * inv;
    *reinterpret_cast<uint4*>(op + d) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------ LayerNorm2d + GELU
// Rows of C <= 128 channels (C % 32 == 0): y = gelu(LN(x) * w + b) -> bf16.  One warp per row.
// LayerNorm2d (sam/common.py:27-39, eps 1e-6) over the channel dim of an NHWC map, fused with the activation of
// MaskDecoder.output_upscaling (mask_decoder.py:59-70, 213-216).
__global__ void ln_rows_gelu_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                    float eps, bf16* __restrict__ y, long long M, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int per = C >> 5;  // 1..4
  float v[4];
  float s = 0.f;
  for (int i = 0; i < per; ++i) { v[i] = x[row * C + lane + i * 32]; s += v[i]; }
  const float mean = warp_sum(s) / C;
  float qv = 0.f;
  for (int i = 0; i < per; ++i) { const float d = v[i] - mean; qv += d * d; }
  const float rstd = rsqrtf(warp_sum(qv) / C + eps);
  for (int i = 0; i < per; ++i) {
    const int c = lane + i * 32;
    const float o = (v[i] - mean) * rstd * w[c] + bias[c];
    y[row * C + c] = __float2bfloat16(es3_act_t<ACT_GELU>(o));
  }
}

// ------------------------------------------------------------------------------------ mask tail
// masks[b][k][p] = sum_c hyper[b][k][c] * up[b][p][c]   (mask_decoder.py:225-226), optional object gating
// (sam3_tracker_base.py:344-353: where(obj_logit > 0, masks, NO_OBJ_SCORE)).  up fp32 [B,HW,CU], CU == 32.
__global__ void hyper_masks_kernel(const float* __restrict__ up, const float* __restrict__ hyper,
                                   const float* __restrict__ obj_logits, float no_obj, float* __restrict__ masks, int HW,
                                   int Ktot, int K, int k_off) {
  __shared__ float sh[8 * 32];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < K * 32; i += blockDim.x) sh[i] = hyper[((long long)b * Ktot + k_off) * 32 + i];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float4* u4 = reinterpret_cast<const float4*>(up + ((long long)b * HW + p) * 32);
  float u[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = u4[j];
    u[4 * j] = t.x; u[4 * j + 1] = t.y; u[4 * j + 2] = t.z; u[4 * j + 3] = t.w;
  }
  const bool gated = obj_logits != nullptr && !(obj_logits[b] > 0.f);
  for (int k = 0; k < K; ++k) {
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) a = fmaf(sh[k * 32 + c], u[c], a);
    masks[((long long)b * K + k) * HW + p] = gated ? no_obj : a;
  }
}

// Bilinear (align_corners=False) NCHW fp32 -> NCHW fp32; optional uint8 binarisation (x > thr).
__global__ void bilinear_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, uint8_t* __restrict__ bin,
                                     float thr, int Hi, int Wi, int Ho, int Wo, float sy, float sx, long long planes) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = planes * Ho * Wo;
  if (idx >= total) return;
  const int ox = (int)(idx % Wo);
  const int oy = (int)((idx / Wo) % Ho);
  const long long pl = idx / ((long long)Wo * Ho);
  float fy = (oy + 0.5f) * sy - 0.5f, fx = (ox + 0.5f) * sx - 0.5f;
  if (fy < 0.f) fy = 0.f;
  if (fx < 0.f) fx = 0.f;
  const int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* p = in + pl * Hi * Wi;
  const float v = hy * (hx * __ldg(p + y0 * Wi + x0) + lx * __ldg(p + y0 * Wi + x1)) +
                  ly * (hx * __ldg(p + y1 * Wi + x0) + lx * __ldg(p + y1 * Wi + x1));
  if (out) out[idx] = v;
  if (bin) bin[idx] = v > thr ? 1 : 0;
}

}  // namespace es3

using namespace es3;

extern "C" int es3_dense_pe(const float* gauss, int F, int h, int w, float* out, void* stream) {
  const int total = h * w * F;
  dense_pe_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(gauss, F, h, w, out);
  ES3_LAUNCH_CHECK("dense_pe_kernel");
  return 0;
}

extern "C" int es3_point_embed(const float* coords, const int* labels, const float* gauss, const float* not_a_point,
                               const float* point_emb, int F, int B, int P, float img_w, float img_h, float* out,
                               void* stream) {
  point_embed_kernel<<<B * (P + 1), 128, 0, (cudaStream_t)stream>>>(coords, labels, gauss, not_a_point, point_emb, F, P, img_w,
                                                                  img_h, out);
  ES3_LAUNCH_CHECK("point_embed_kernel");
  return 0;
}

extern "C" int es3_add_rows(const float* x, const float* add, long long M, int C, int R, void* y_bf16, float* y_f32,
                            void* stream) {
  ES3_REQUIRE(C % 4 == 0 && (add == nullptr || R > 0), "es3_add_rows: bad C=%d R=%d", C, R);
  const long long n4 = M * (C / 4);
  add_rows_kernel<<<(unsigned)ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(x, add, M, C, R > 0 ? R : 1, (bf16*)y_bf16, y_f32);
  ES3_LAUNCH_CHECK("add_rows_kernel");
  return 0;
}

extern "C" int es3_nchw_f32_to_tokens(const float* in, const float* addc, float* out_f32, void* out_bf16, int B, int HW,
                                      int C, void* stream) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B);
  nchw_to_tokens_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, addc, out_f32, (bf16*)out_bf16, HW, C);
  ES3_LAUNCH_CHECK("nchw_to_tokens_kernel");
  return 0;
}

extern "C" int es3_attn_few_queries(const float* q, long long ldq, const void* k, const void* v, long long ldkv, int kv_f32,
                                    float* out, long long ldo, int B, int H, int head_dim, int Tq, int Tk, float scale,
                                    void* stream) {
  ES3_REQUIRE(head_dim == 16 || head_dim == 32, "es3_attn_few_queries: head_dim %d not in {16,32}", head_dim);
  dim3 grid(ceil_div(Tq, 4), H, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (head_dim == 16) {
    if (kv_f32) attn_few_queries_kernel<16, float><<<grid, 128, 0, st>>>(q, ldq, (const float*)k, (const float*)v, ldkv, out, ldo, Tq, Tk, scale);
    else attn_few_queries_kernel<16, bf16><<<grid, 128, 0, st>>>(q, ldq, (const bf16*)k, (const bf16*)v, ldkv, out, ldo, Tq, Tk, scale);
  } else {
    if (kv_f32) attn_few_queries_kernel<32, float><<<grid, 128, 0, st>>>(q, ldq, (const float*)k, (const float*)v, ldkv, out, ldo, Tq, Tk, scale);
    else attn_few_queries_kernel<32, bf16><<<grid, 128, 0, st>>>(q, ldq, (const bf16*)k, (const bf16*)v, ldkv, out, ldo, Tq, Tk, scale);
  }
  ES3_LAUNCH_CHECK("attn_few_queries_kernel");
  return 0;
}

extern "C" int es3_attn_few_keys(const void* q, long long ldq, const float* k, const float* v, long long ldkv, void* out,
                                 long long ldo, int B, int H, int head_dim, int Nq, int Tk, float scale, void* stream) {
  ES3_REQUIRE(head_dim == 16 && Tk <= 16, "es3_attn_few_keys: head_dim must be 16 and Tk <= 16 (got %d, %d)", head_dim, Tk);
  const size_t smem = (size_t)2 * Tk * H * head_dim * sizeof(float);
  dim3 grid(ceil_div((long long)Nq * H, 256), B);
  attn_few_keys_kernel<16><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)q, ldq, k, v, ldkv, (bf16*)out, ldo, Nq, Tk, H, scale);
  ES3_LAUNCH_CHECK("attn_few_keys_kernel");
  return 0;
}

extern "C" int es3_ln_rows_gelu(const float* x, const float* w, const float* bias, float eps, void* y, long long M, int C,
                                void* stream) {
  ES3_REQUIRE(C % 32 == 0 && C <= 128, "es3_ln_rows_gelu: C=%d must be a multiple of 32 and <= 128", C);
  ln_rows_gelu_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, w, bias, eps, (bf16*)y, M, C);
  ES3_LAUNCH_CHECK("ln_rows_gelu_kernel");
  return 0;
}

extern "C" int es3_hyper_masks(const float* up, const float* hyper, const float* obj_logits, float no_obj, float* masks,
                               int B, int HW, int CU, int Ktot, int K, int k_off, void* stream) {
  ES3_REQUIRE(CU == 32 && K <= 8 && k_off + K <= Ktot, "es3_hyper_masks: CU must be 32, K <= 8, k_off + K <= Ktot");
  dim3 grid(ceil_div(HW, 256), B);
  hyper_masks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(up, hyper, obj_logits, no_obj, masks, HW, Ktot, K, k_off);
  ES3_LAUNCH_CHECK("hyper_masks_kernel");
  return 0;
}

extern "C" int es3_bilinear_nchw_f32(const float* in, float* out, void* bin, float thr, long long planes, int Hi, int Wi,
                                     int Ho, int Wo, void* stream) {
  const long long total = planes * Ho * Wo;
  bilinear_nchw_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      in, out, (uint8_t*)bin, thr, Hi, Wi, Ho, Wo, (float)Hi / Ho, (float)Wi / Wo, planes);
  ES3_LAUNCH_CHECK("bilinear_nchw_kernel");
  return 0;
}
