// Weight gradient of the pointwise (1x1) convolutions on tcgen05:  dW[n][k] += sum_m dz[m][n] * x[m][k]
// (autograd of nn.Conv2d(k=1) / nn.Linear inside the student's training step, stage1/train_image_encoder_stage1.py:199-227).
//
// It is a dense contraction over the PIXEL index m (33 K .. 8 M rows) with small outputs (N, K <= 1024): a split-K GEMM whose two
// operands both have the contraction index as their slow dimension.  The round-1 kernel (es3_wgrad_pw, mma.sync + ldmatrix.trans)
// ran at 44-62 TFLOP/s on the wide layers (profiles/r1_train_step_c.md).  Here:
//
//   TMA    box {64 channels, 64 pixels} of dz and of x  ->  128B-swizzled tiles [64 px][128 B]: exactly the MN-major ("transposed")
//          shared-memory operand layout of a UMMA -- no transpose pass, no ldmatrix
//   UMMA   D[128 (n) x KT (k)] += A^T B with a_major = b_major = MN (instruction-descriptor bits 15 / 16), K = 16 pixels per
//          instruction, accumulating over this CTA's pixel range in TMEM (KT <= 256 fp32 columns)
//   split  grid = (n-tiles x k-tiles, nsplit): each CTA owns one output tile and one contiguous pixel range; partial tiles go to a
//          workspace and a second kernel adds them into dW in a fixed order (deterministic: no atomics)
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + UMMA issuer, warps 2-5 = epilogue (one TMEM lane quarter each).
// Channel counts below a 64-wide tile read zeros (TMA out-of-bounds fill) or neighbouring columns of a strided view; rows / columns
// past N / K are never stored.  Shapes with N % 64 != 0 or K % 64 != 0 stay on es3_wgrad_pw (the entry point returns -1).
#include <cuda.h>

#include "ptx.cuh"

namespace es3 {

int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b, const uint32_t* box);

constexpr int WG_PX = 64;                 // pixels per stage (4 UMMA k-steps)
constexpr int WG_TILE = WG_PX * 128;      // one [64 px][64 ch] tile: 8 KB
constexpr int WG_STAGES = 4;
constexpr int WG_THREADS = 192;

template <int KT>
struct WGSmem {
  static constexpr int A_BYTES = 2 * WG_TILE;               // 128 dz channels
  static constexpr int B_BYTES = (KT / 64) * WG_TILE;       // KT x channels
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int TOTAL = WG_STAGES * STAGE;
};

struct WGArgs {
  long long M;
  int N, K, tiles_k, nsplit;
  long long px_per_split;     // multiple of WG_PX
  float* ws;                  // [nsplit][N][K]
};

// MN-major SW128 operand: 8-pixel atoms of 1024 B along the contraction (SBO), 64-channel blocks WG_TILE bytes apart (LBO).
__device__ __forceinline__ uint64_t wg_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(WG_TILE >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KT>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_x, const WGArgs a) {
  using L = WGSmem<KT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[WG_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = ((int)blockIdx.x / a.tiles_k) * 128, k0 = ((int)blockIdx.x % a.tiles_k) * KT;
  const int split = blockIdx.y;
  const long long px0 = (long long)split * a.px_per_split;
  long long px1 = px0 + a.px_per_split;
  if (px1 > a.M) px1 = a.M;
  const int nblk = px1 > px0 ? (int)((px1 - px0 + WG_PX - 1) / WG_PX) : 0;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) { printf("es3: wgrad_tc dynamic smem base not 1024-byte aligned\n"); __trap(); }
    ptx::prefetch_tmap(&tm_dz); ptx::prefetch_tmap(&tm_x);
#pragma unroll
    for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    ptx::mbar_init(&acc_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_holder, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE;
        uint8_t* sb = sa + L::A_BYTES;
        const int px = (int)(px0 + (long long)b * WG_PX);          // M < 2^31 (checked on the host)
        ptx::mbar_arrive_expect_tx(&full_bar[stage], L::STAGE);
        ptx::tma_load_2d(&tm_dz, &full_bar[stage], sa, n0, px);
        ptx::tma_load_2d(&tm_dz, &full_bar[stage], sa + WG_TILE, n0 + 64, px);
#pragma unroll
        for (int c = 0; c < KT / 64; ++c) ptx::tma_load_2d(&tm_x, &full_bar[stage], sb + c * WG_TILE, k0 + c * 64, px);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, KT) | (1u << 15) | (1u << 16);   // A and B MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int b = 0; b < nblk; ++b) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * L::STAGE);
        const uint64_t da = wg_desc_mn(sa), db = wg_desc_mn(sa + L::A_BYTES);
#pragma unroll
        for (int ks = 0; ks < WG_PX / 16; ++ks)     // 16 pixels = 2 atoms of 1024 B = +128 in the >>4 address field
          ptx::umma_f16(tmem, da + (uint64_t)(ks * 128), db + (uint64_t)(ks * 128), idesc, (b | ks) != 0);
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(&acc_bar);
    }
  } else {
    // ---- epilogue: TMEM -> workspace partial tile (rows n0 + 32 q + lane, columns k0 ...)
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    float* dst = a.ws + ((long long)split * a.N + n) * a.K + k0;
    if (nblk > 0) {
      ptx::mbar_wait(&acc_bar, 0);
      ptx::tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < KT / 32; ++c) {
      if (k0 + c * 32 >= a.K) break;
      uint32_t v[32];
      if (nblk > 0) {
        ptx::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (n < a.N) {
        float4* o = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

// dW[n * ldn + k] += sum over the splits, in split order
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ ws, int nsplit, int N, int K, float* __restrict__ dW, long long ldn) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  float acc = 0.f;
  for (int s = 0; s < nsplit; ++s) acc += ws[(long long)s * N * K + i];
  const int n = (int)(i / K), k = (int)(i % K);
  dW[(long long)n * ldn + k] += acc;
}

static void wgrad_tc_plan(long long M, int N, int K, int* kt, int* tiles, int* nsplit, long long* pps) {
  *kt = K >= 256 ? 256 : (K >= 128 ? 128 : 64);
  const int tn = (N + 127) / 128, tk = (K + *kt - 1) / *kt;
  *tiles = tn * tk;
  long long blocks = (M + WG_PX - 1) / WG_PX;
  int ns = 148 / *tiles;                       // one CTA per SM (192 KB of operand stages)
  if (ns < 1) ns = 1;
  if (ns > blocks) ns = (int)blocks;
  long long bps = (blocks + ns - 1) / ns;      // pixel blocks per split
  *nsplit = (int)((blocks + bps - 1) / bps);
  *pps = bps * WG_PX;
}

template <int KT>
static int launch_wgrad_tc(const CUtensorMap& tm_dz, const CUtensorMap& tm_x, const WGArgs& a, int tiles, cudaStream_t st) {
  using L = WGSmem<KT>;
  ES3_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  wgrad_tc_kernel<KT><<<dim3(tiles, a.nsplit), WG_THREADS, L::TOTAL, st>>>(tm_dz, tm_x, a);
  ES3_LAUNCH_CHECK("wgrad_tc_kernel");
  return 0;
}

}  // namespace es3

using namespace es3;

extern "C" long long es3_wgrad_tc_ws_floats(long long M, int N, int K) {
  int kt, tiles, nsplit;
  long long pps;
  wgrad_tc_plan(M, N, K, &kt, &tiles, &nsplit, &pps);
  return (long long)nsplit * N * K;
}

/* dW[n * ldn + k] += sum_m dz[m * lddz + n] * x[m * ldx + k], dz / x bf16 with unit channel stride, dW fp32.
 * Returns -1 (no error set) when the shape is not on this kernel (N or K not a multiple of 64, unaligned strides): the caller then
 * uses es3_wgrad_pw.  ws: es3_wgrad_tc_ws_floats(M, N, K) floats. */
extern "C" int es3_wgrad_tc(const void* dz, long long lddz, const void* x, long long ldx, long long M, int N, int K, float* ws,
                            float* dW, long long ldn, void* stream) {
  if (N % 64 != 0 || K % 64 != 0 || N < 64 || K < 64 || lddz % 8 != 0 || ldx % 8 != 0 || M < WG_PX || M >= (1LL << 31) ||
      (((uintptr_t)dz | (uintptr_t)x) & 15) != 0)
    return -1;
  int kt, tiles, nsplit;
  long long pps;
  wgrad_tc_plan(M, N, K, &kt, &tiles, &nsplit, &pps);
  CUtensorMap tm_dz, tm_x;
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lddz * 2};
    uint32_t box[2] = {64u, (uint32_t)WG_PX};
    if (encode_map(&tm_dz, dz, 2, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {64u, (uint32_t)WG_PX};
    if (encode_map(&tm_x, x, 2, dims, str, box)) return 1;
  }
  WGArgs a;
  a.M = M; a.N = N; a.K = K; a.tiles_k = (K + kt - 1) / kt; a.nsplit = nsplit; a.px_per_split = pps; a.ws = ws;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (kt == 256) rc = launch_wgrad_tc<256>(tm_dz, tm_x, a, tiles, st);
  else if (kt == 128) rc = launch_wgrad_tc<128>(tm_dz, tm_x, a, tiles, st);
  else rc = launch_wgrad_tc<64>(tm_dz, tm_x, a, tiles, st);
  if (rc) return rc;
  const long long total = (long long)N * K;
  wgrad_tc_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, nsplit, N, K, dW, ldn);
  ES3_LAUNCH_CHECK("wgrad_tc_reduce_kernel");
  return 0;
}
