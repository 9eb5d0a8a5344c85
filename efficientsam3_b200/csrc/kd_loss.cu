// Stage-1 distillation loss, forward (stage1/train_image_encoder_stage1.py:271-307):
//   mask   = bilinear_{img->E}(1[y < h_i, x < w_i]) > 0.5                         (build_valid_mask)
//   mse_i  = sum_{c,p} ((pred - teacher) * mask)^2 / max(sum_p mask, 1)           (masked_mse; note: mask COUNT, not x C)
//   cos_i  = sum_p (1 - cos_C(pred_p, teacher_p)) * mask_p / max(sum_p mask, 1)   (masked_cosine_loss)
//   loss   = mean_i mse_i + cosine_weight * mean_i cos_i                          (:205-210, DISTILL.COSINE = 1.0)
// One streaming pass over preds / teacher ([B,C,E,E] fp32 NCHW, the modules' output layout): thread = pixel,
// loop over channels (coalesced along pixels), deterministic two-stage reduction (no atomics).
#include "common.cuh"

namespace es3 {

__device__ __forceinline__ float valid_mask_at(int oy, int ox, int E, int img, int h, int w) {
  const float s = (float)img / (float)E;
  float fy = (oy + 0.5f) * s - 0.5f, fx = (ox + 0.5f) * s - 0.5f;
  if (fy < 0.f) fy = 0.f;
  if (fx < 0.f) fx = 0.f;
  const int y0 = min((int)fy, img - 1), x0 = min((int)fx, img - 1);
  const int y1 = min(y0 + 1, img - 1), x1 = min(x0 + 1, img - 1);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const float i00 = (y0 < h && x0 < w) ? 1.f : 0.f, i01 = (y0 < h && x1 < w) ? 1.f : 0.f;
  const float i10 = (y1 < h && x0 < w) ? 1.f : 0.f, i11 = (y1 < h && x1 < w) ? 1.f : 0.f;
  const float v = hy * (hx * i00 + lx * i01) + ly * (hx * i10 + lx * i11);
  return v > 0.5f ? 1.f : 0.f;
}

// grid (ceil(E*E/256), B).  part: [B][gridDim.x][3] = (sum sq, sum (1-cos)*mask, mask count)
__global__ void kd_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ teach,
                                       const int* __restrict__ sizes_hw, int C, int E, int img, float* __restrict__ part) {
  __shared__ float red[3][8];
  const int b = blockIdx.y, HW = E * E;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.f, cl = 0.f, mk = 0.f;
  if (p < HW) {
    const float m = valid_mask_at(p / E, p % E, E, img, sizes_hw[b * 2], sizes_hw[b * 2 + 1]);
    if (m > 0.f) {
      const float* pp = pred + (long long)b * C * HW + p;
      const float* tp = teach + (long long)b * C * HW + p;
      float dot = 0.f, np2 = 0.f, nt2 = 0.f, d2 = 0.f;
#pragma unroll 4
      for (int c = 0; c < C; ++c) {
        const float a = __ldg(pp + (long long)c * HW), t = __ldg(tp + (long long)c * HW);
        dot = fmaf(a, t, dot); np2 = fmaf(a, a, np2); nt2 = fmaf(t, t, nt2);
        const float d = a - t;
        d2 = fmaf(d, d, d2);
      }
      const float cosv = dot / (fmaxf(sqrtf(np2), 1e-8f) * fmaxf(sqrtf(nt2), 1e-8f));  // F.cosine_similarity, eps 1e-8
      sq = d2; cl = 1.f - cosv; mk = 1.f;
    }
  }
  sq = warp_sum(sq); cl = warp_sum(cl); mk = warp_sum(mk);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = sq; red[1][w] = cl; red[2][w] = mk; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[threadIdx.x][i];
    part[((long long)b * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
  }
}

// out[0] = loss, out[1] = mse term, out[2] = cosine term; per_sample [B][3] (mse_i, cos_i, mask count) optional
__global__ void kd_loss_final_kernel(const float* __restrict__ part, int B, int nblk, float cosine_w, float* __restrict__ out,
                                     float* __restrict__ per_sample) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mse = 0.f, cosl = 0.f;
  for (int b = 0; b < B; ++b) {
    float sq = 0.f, cl = 0.f, mk = 0.f;
    for (int i = 0; i < nblk; ++i) {
      sq += part[((long long)b * nblk + i) * 3 + 0];
      cl += part[((long long)b * nblk + i) * 3 + 1];
      mk += part[((long long)b * nblk + i) * 3 + 2];
    }
    const float den = fmaxf(mk, 1.f);
    mse += sq / den;
    cosl += cl / den;
    if (per_sample) { per_sample[b * 3] = sq / den; per_sample[b * 3 + 1] = cl / den; per_sample[b * 3 + 2] = mk; }
  }
  mse /= B; cosl /= B;
  out[0] = mse + cosine_w * cosl; out[1] = mse; out[2] = cosl;
}

}  // namespace es3

using namespace es3;

// ws: B * ceil(E*E/256) * 3 floats.  sizes_hw: int32 [B][2] = (h, w) of each image before padding.
extern "C" int es3_kd_loss_fwd(const float* preds, const float* teacher, const int* sizes_hw, int B, int C, int E, int img_size,
                               float cosine_weight, float* ws, float* out3, float* per_sample, void* stream) {
  ES3_REQUIRE(B > 0 && C > 0 && E > 0 && img_size > 0, "es3_kd_loss_fwd: bad shape");
  const int nblk = ceil_div(E * E, 256);
  cudaStream_t st = (cudaStream_t)stream;
  kd_loss_partial_kernel<<<dim3(nblk, B), 256, 0, st>>>(preds, teacher, sizes_hw, C, E, img_size, ws);
  ES3_LAUNCH_CHECK("kd_loss_partial_kernel");
  kd_loss_final_kernel<<<1, 32, 0, st>>>(ws, B, nblk, cosine_weight, out3, per_sample);
  ES3_LAUNCH_CHECK("kd_loss_final_kernel");
  return 0;
}
