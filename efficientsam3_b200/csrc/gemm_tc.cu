// tcgen05 / TMEM / TMA GEMM for sm_100a with a fused per-channel epilogue.
//
//   out[m, n] = act(scale[n] * sum_k A[m, k] * W[n, k] + bias[n]) (+ residual[m, n])
//
// A: bf16 activations, K contiguous (NHWC pixels x channels, or tokens x features)
// W: bf16 weights [N][K], K contiguous (nn.Linear / 1x1 nn.Conv2d weight layout)
// accumulation: fp32 in tensor memory; out: bf16 or fp32.
//
// This one kernel serves every dense contraction on the hot path:
//   * pointwise (1x1) convs of EfficientViT / RepViT / TinyViT MBConv + the student head
//     (reference: efficientvit/nn/ops.py:39-80 ConvLayer with kernel_size=1, stage1/model.py:194-199)
//   * nn.Linear of the ViT teacher / TinyViT / TwoWayTransformer (vitdet.py:466-515, timm Mlp)
//   * dense 3x3 conv as an implicit GEMM (head.3: stage1/model.py:198; FPN neck necks.py) -- the A tile
//     of tap (dy, dx) is a 4-D TMA box at coordinate (c, w0+dx-1, h0+dy-1, b); TMA zero-fills the halo.
//
// Structure (persistent CTAs, 128 x BN output tiles, 320 threads):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor into a STAGES-deep 128B-swizzled smem ring
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16)
//   warps 2-9: epilogue       -- tcgen05.ld 32x32b -> registers -> scale/bias/act/rope/residual -> global
// Pipelines: full[s]/empty[s] mbarriers (TMA <-> MMA), tmem_full[2]/tmem_empty[2] (MMA <-> epilogue) over a
// double-buffered TMEM accumulator, so the epilogue of one tile overlaps the main loop of the next.
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

namespace es3 {

struct GemmArgs {
  int M, N;
  int num_kb;      // total K iterations (64 elements each); conv: 9 * kb_per_tap
  int kb_per_tap;  // conv: C / 64 rounded up
  int tiles_n;
  int conv;  // 0: plain [M,K] A map (2-D); 1: 3x3 implicit GEMM, A map is 4-D (C, W, H, B)
  int H, W, TW, TH, tiles_w, tiles_h;
  int Ctap;  // conv: channels per tap (K offset of tap t in W is t * Ctap)
  const float* scale;
  const float* bias;
  int act;
  const bf16* residual;
  long long ldr;
  void* out;
  long long ldo;
  int out_f32;
  int res_f32;          // residual is fp32 (the ViT residual stream) instead of bf16
  // 2-D axial RoPE fused into the QKV projection epilogue (vitdet.py:68-90, 421-457): columns
  // [0, rope_cols) are (q | k) heads of 64 dims = 32 complex pairs; rope: [positions][32] (cos, sin).
  const float2* rope;
  int rope_cols, rope_H, rope_W, rope_win;
  // ConvTranspose2d(k=2, s=2) as a GEMM with N = 4*Cout (n = (dy*2+dx)*Cout + co): the epilogue scatters row
  // (b,h,w) / column n to output pixel (b, 2h+dy, 2w+dx), channel co (depth-to-space).  0 = off.
  int ct_cout, ct_H, ct_W;
  int act_after_res;    // apply the activation after the residual add (MaskDecoder: gelu(dc2(x) + feat_s0))
  int dbg;              // ES3_GEMM_DBG bits (bottleneck bisection only; 0 in production): 1 no stores, 2 no tmem loads, 4 epilogue = hand-back only, 8 loads pinned to tile 0
};

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row

constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane quarter, splitting the column chunks
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue
// CPR = 16-byte chunks per row of the per-warp transposition tile (see epi_store_rows): 4 -> 32 rows x 64 B = 2 KB per warp,
// 2 -> 32 rows x 32 B = 1 KB per warp (the deep-pipeline BN = 128 configuration, which must keep two CTAs per SM: the first
// version of this epilogue spent 114688 B and the runtime reported ONE resident CTA -- gpurun_out/r2g_bisect.log)
template <int BN, int STAGES, int CPR>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES;
  static constexpr int EPI_TILE_BYTES = 32 * 16 * CPR;
  static constexpr int TOTAL = OFF_EPI + EPI_WARPS * EPI_TILE_BYTES;   // no alignment slack: the buffer is declared __align__(1024)
};

// ---- coalescing the epilogue's global traffic ------------------------------------------------------------------------------
// tcgen05.ld 32x32b hands lane r the 32 columns of ROW r.  Stored straight from there, one STG.128 touches 32 different rows
// (32 cache-line wavefronts of 16 B each): ncu showed the short-K GEMMs of the students sitting on exactly that
// (lg_throttle 3.9 / issue, 17 % of the samples on the four STG.128, profiles/r2c_*: 2048 sixteen-byte write requests per
// 128 x 128 tile).  Each epilogue warp therefore owns a 2 KB shared-memory tile of 32 rows x 64 B: the row owner writes its four
// 16-byte chunks (XOR-swizzled by (row >> 1) & 3: conflict-free both ways), then lane l moves chunk (l & 3) of row 8 j + (l >> 2),
// j = 0..3 -- one instruction now covers 8 rows x 64 contiguous bytes (8 wavefronts of two full sectors).  The same tile runs
// backwards for the residual operand.  Row offsets / validity travel by shuffle, so every row mapping (plain, implicit-GEMM conv
// tiles, ConvTranspose depth-to-space) shares the code.
template <int CPR>
__device__ __forceinline__ uint8_t* epi_chunk(uint8_t* tile, int r, int c) {
  if constexpr (CPR == 4) return tile + r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
  else return tile + r * 32 + ((c ^ ((r >> 2) & 1)) << 4);
}

// `mine`: this lane's row, 64 bytes = 4 chunks; moved to global in 64 / (16 CPR) passes through the tile
template <int CPR>
__device__ __forceinline__ void epi_store_rows(uint8_t* tile, const uint4* mine, uint8_t* gbase, long long my_off_bytes,
                                               bool my_valid, int lane) {
  constexpr int RPI = 32 / CPR;           // rows per store instruction
#pragma unroll
  for (int p = 0; p < 4 / CPR; ++p) {
#pragma unroll
    for (int j = 0; j < CPR; ++j) *reinterpret_cast<uint4*>(epi_chunk<CPR>(tile, lane, j)) = mine[p * CPR + j];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CPR; ++j) {
      const int r = j * RPI + lane / CPR, c = lane % CPR;
      const long long off = __shfl_sync(0xffffffffu, my_off_bytes, r);
      const int ok = __shfl_sync(0xffffffffu, (int)my_valid, r);
      const uint4 u = *reinterpret_cast<const uint4*>(epi_chunk<CPR>(tile, r, c));
      if (ok) *reinterpret_cast<uint4*>(gbase + off + (p * CPR + c) * 16) = u;
    }
    __syncwarp();
  }
}

template <int CPR>
__device__ __forceinline__ void epi_load_rows(uint8_t* tile, uint4* mine, const uint8_t* gbase, long long my_off_bytes,
                                              bool my_valid, int lane) {
  constexpr int RPI = 32 / CPR;
#pragma unroll
  for (int p = 0; p < 4 / CPR; ++p) {
#pragma unroll
    for (int j = 0; j < CPR; ++j) {
      const int r = j * RPI + lane / CPR, c = lane % CPR;
      const long long off = __shfl_sync(0xffffffffu, my_off_bytes, r);
      const int ok = __shfl_sync(0xffffffffu, (int)my_valid, r);
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (ok) u = __ldg(reinterpret_cast<const uint4*>(gbase + off + (p * CPR + c) * 16));
      *reinterpret_cast<uint4*>(epi_chunk<CPR>(tile, r, c)) = u;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CPR; ++j) mine[p * CPR + j] = *reinterpret_cast<const uint4*>(epi_chunk<CPR>(tile, lane, j));
    __syncwarp();
  }
}

// Persistent kernel: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so CTAs that
// run together share the A row-block in L2).  The accumulator is double-buffered in TMEM (2 x BN columns):
// the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
// (min 2 CTAs/SM for BN <= 128: the GELU instantiations grew to 130-138 registers with the staged epilogue and silently dropped to one
//  resident CTA -- TinyViT's fc1 GEMMs lost 13-30 %, gpurun_out/r2k_table_tvm.md)
template <int BN, int STAGES, int ACT, int CPR>
__global__ void __launch_bounds__(GEMM_THREADS, BN >= 256 ? 1 : 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const GemmArgs args, const int num_tiles) {
  // (no integer round trip on the pointer: the compiler keeps shared-space addressing for the staging tiles)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_holder;

  using L = GemmSmem<BN, STAGES, CPR>;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) { printf("es3: gemm_tc dynamic smem base not 1024-byte aligned\n"); __trap(); }
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full_bar[s], 1);
      ptx::mbar_init(&tmem_empty_bar[s], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_base_holder, 2 * BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  const int per_img = args.tiles_w * args.tiles_h;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (args.dbg & 8) ? 0 : (tile % args.tiles_n) * BN;
        const int m_tile = (args.dbg & 8) ? 0 : tile / args.tiles_n;
        int img = 0, h0 = 0, w0 = 0;
        if (args.conv) {
          img = m_tile / per_img;
          const int t = m_tile % per_img;
          h0 = (t / args.tiles_w) * args.TH;
          w0 = (t % args.tiles_w) * args.TW;
        }
        for (int kb = 0; kb < args.num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (args.conv) {
            const int tap = kb / args.kb_per_tap;
            const int kc = kb - tap * args.kb_per_tap;
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            ptx::tma_load_4d(&tmA, &full_bar[stage], sa, kc * BK, w0 + dx, h0 + dy, img);
            ptx::tma_load_2d(&tmB, &full_bar[stage], sb, tap * args.Ctap + kc * BK, n0);
          } else {
            ptx::tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m_tile * BM);
            ptx::tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < args.num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          const uint64_t da = ptx::make_desc_sw128(sa);
          const uint64_t db = ptx::make_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the >>4 address field
            ptx::umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which of the two warps of the quarter: interleaved column chunks
    const int r = q * 32 + lane;            // row inside the tile
    uint8_t* epi_tile = smem + L::OFF_EPI + (warp - 2) * L::EPI_TILE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n0 = (tile % args.tiles_n) * BN;
      const int m_tile = tile / args.tiles_n;
      bool valid;
      long long row_off;
      const bool dbg_nostore = (args.dbg & 1) != 0;
      if (args.conv) {
        const int img = m_tile / per_img;
        const int t = m_tile % per_img;
        const int h0 = (t / args.tiles_w) * args.TH, w0 = (t % args.tiles_w) * args.TW;
        const int dy = r / args.TW, dx = r - dy * args.TW;
        const int h = h0 + dy, w = w0 + dx;
        valid = (h < args.H) && (w < args.W);
        row_off = ((long long)img * args.H + h) * args.W + w;
      } else {
        const long long m = (long long)m_tile * BM + r;
        valid = m < args.M;
        row_off = m;
      }

      if (dbg_nostore) valid = false;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();

#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        if (args.dbg & 4) break;
        uint32_t v[32];
        if (args.dbg & 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(r + j));
        } else {
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
          ptx::tmem_ld_wait();
        }
        const int nb = n0 + c * 32;
        if (valid && nb < args.N && nb + 32 > args.N) {
          // ragged last chunk (N % 32 != 0, N % 8 == 0): per-column path, plain GEMM epilogue only
          const int ncols = args.N - nb;
          const long long out_off = row_off * args.ldo + nb, res_off = row_off * args.ldr + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < ncols) {
              float x = __uint_as_float(v[j]);
              if (args.scale != nullptr) x *= __ldg(args.scale + nb + j);
              if (args.bias != nullptr) x += __ldg(args.bias + nb + j);
              if (!args.act_after_res) x = es3_act_t<ACT>(x);
              if (args.residual != nullptr)
                x += args.res_f32 ? __ldg(reinterpret_cast<const float*>(args.residual) + res_off + j)
                                  : __bfloat162float(args.residual[res_off + j]);
              if (args.act_after_res) x = es3_act_t<ACT>(x);
              if (args.out_f32) reinterpret_cast<float*>(args.out)[out_off + j] = x;
              else reinterpret_cast<bf16*>(args.out)[out_off + j] = __float2bfloat16(x);
            }
          }
        } else if (nb + 32 <= args.N) {
          // every lane of the warp comes through here together (nb, N are warp-uniform): the loads / stores below are
          // cooperative; rows past M (or outside the conv tile) compute on whatever TMEM holds and are masked at the store
          float f[32];
          if (args.scale != nullptr) {
            const float4* sc4 = reinterpret_cast<const float4*>(args.scale + nb);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sc = __ldg(sc4 + j);
              f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) * sc.x;
              f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) * sc.y;
              f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) * sc.z;
              f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) * sc.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          }
          if (args.bias != nullptr) {
            const float4* bi4 = reinterpret_cast<const float4*>(args.bias + nb);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bi = __ldg(bi4 + j);
              f[4 * j + 0] += bi.x; f[4 * j + 1] += bi.y; f[4 * j + 2] += bi.z; f[4 * j + 3] += bi.w;
            }
          }
          long long out_off = row_off * args.ldo + nb, res_off = row_off * args.ldr + nb;
          if (args.ct_cout) {
            const int pos = nb / args.ct_cout, co = nb - pos * args.ct_cout;
            const long long hw = (long long)args.ct_H * args.ct_W;
            const long long bi = row_off / hw;
            const int rem = (int)(row_off - bi * hw);
            const int hh = rem / args.ct_W, ww = rem - hh * args.ct_W;
            const long long orow = (bi * 2 * args.ct_H + 2 * hh + (pos >> 1)) * (2LL * args.ct_W) + 2 * ww + (pos & 1);
            out_off = orow * args.ldo + co;
            res_off = orow * args.ldr + co;
          }
          if (!args.act_after_res) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = es3_act_t<ACT>(f[j]);
          }
          if (args.rope != nullptr && nb < args.rope_cols) {
            const int t = (int)(row_off % ((long long)args.rope_H * args.rope_W));
            const int h = t / args.rope_W, w = t - h * args.rope_W;
            const int pidx = args.rope_win ? (h % args.rope_win) * args.rope_win + (w % args.rope_win) : t;
            const float4* tp = reinterpret_cast<const float4*>(args.rope + (long long)pidx * 32 + ((nb & 63) >> 1));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 cs = __ldg(tp + j);  // (cos, sin) of two consecutive pairs
              const float x0 = f[4 * j], x1 = f[4 * j + 1], x2 = f[4 * j + 2], x3 = f[4 * j + 3];
              f[4 * j] = x0 * cs.x - x1 * cs.y;
              f[4 * j + 1] = x0 * cs.y + x1 * cs.x;
              f[4 * j + 2] = x2 * cs.z - x3 * cs.w;
              f[4 * j + 3] = x2 * cs.w + x3 * cs.z;
            }
          }
          if (args.residual != nullptr && args.res_f32) {     // 32 fp32 = two 64-byte halves through the staging tile
            const uint8_t* rbase = reinterpret_cast<const uint8_t*>(args.residual);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint4 rr[4];
              epi_load_rows<CPR>(epi_tile, rr, rbase, res_off * 4 + hf * 64, valid, lane);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f[hf * 16 + 4 * j + 0] += __uint_as_float(rr[j].x); f[hf * 16 + 4 * j + 1] += __uint_as_float(rr[j].y);
                f[hf * 16 + 4 * j + 2] += __uint_as_float(rr[j].z); f[hf * 16 + 4 * j + 3] += __uint_as_float(rr[j].w);
              }
            }
          } else if (args.residual != nullptr) {
            uint4 rr[4];
            epi_load_rows<CPR>(epi_tile, rr, reinterpret_cast<const uint8_t*>(args.residual), res_off * 2, valid, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float rf[8];
              unpack8(rr[j], rf);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[j * 8 + e] += rf[e];
            }
          }
          if (args.act_after_res) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = es3_act_t<ACT>(f[j]);
          }
          if (args.out_f32) {
            uint8_t* obase = reinterpret_cast<uint8_t*>(args.out);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint4 oo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                oo[j] = make_uint4(__float_as_uint(f[hf * 16 + 4 * j]), __float_as_uint(f[hf * 16 + 4 * j + 1]),
                                   __float_as_uint(f[hf * 16 + 4 * j + 2]), __float_as_uint(f[hf * 16 + 4 * j + 3]));
              epi_store_rows<CPR>(epi_tile, oo, obase, out_off * 4 + hf * 64, valid, lane);
            }
          } else {
            uint4 oo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) oo[j] = pack8(f + 8 * j);
            epi_store_rows<CPR>(epi_tile, oo, reinterpret_cast<uint8_t*>(args.out), out_off * 2, valid, lane);
          }
        }
      }
      // this warp is done reading the accumulator: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// rank-N bf16 tensor map, 128B swizzle, zero OOB fill. dims/strides innermost first; strides in BYTES for dims 1..
int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_b[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,..] box=[%u,%u,..]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return 1;
  }
  return 0;
}

template <int BN, int STAGES, int ACT, int CPR>
static int launch_act(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args, int tiles_m,
                      cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, CPR>;
  static int resident = 0;     // CTAs of this instantiation one SM holds (asked of the runtime: smem is sized to the last KB)
  static int sm_count = 0;
  if (resident == 0) {
    ES3_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, ACT, CPR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        L::TOTAL));
    // ask for the full shared-memory carveout: with the default preference the runtime reported ONE resident CTA even for 80 KB
    // blocks (gpurun_out/r2h_bisect.log) and a grid sized from that left half of every SM's slots empty
    ES3_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, ACT, CPR>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    int dev = 0, occ = 0;
    ES3_CHECK_CUDA(cudaGetDevice(&dev));
    ES3_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    ES3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_tc_kernel<BN, STAGES, ACT, CPR>, GEMM_THREADS, L::TOTAL));
    ES3_REQUIRE(occ >= 1, "gemm_tc_kernel<%d,%d>: does not fit an SM (%d B of dynamic shared memory)", BN, STAGES, L::TOTAL);
    // slots per SM: TMEM (2 * BN of 512 columns) and shared memory (228 KB minus 1 KB per CTA) allow two CTAs for BN <= 128;
    // the grid is a persistent tile loop, so an over-estimate only costs a second wave
    resident = (BN >= 256 || 2 * (L::TOTAL + 2048) > 233472) ? 1 : 2;
    if (getenv("ES3_DEBUG_OCCUPANCY"))
      fprintf(stderr, "es3: gemm_tc_kernel<BN=%d,STAGES=%d,ACT=%d,CPR=%d> %d B smem -> %d CTA/SM (runtime occupancy %d)\n", BN, STAGES, ACT,
              CPR, L::TOTAL, resident, occ);
  }
  const int num_tiles = tiles_m * args.tiles_n;
  // persistent grid: every resident slot of every SM, BN = 256 uses all 512 TMEM columns and ~209 KB of smem (one CTA per SM)
  const int ctas = sm_count * resident;
  dim3 grid((unsigned)(num_tiles < ctas ? num_tiles : ctas));
  gemm_tc_kernel<BN, STAGES, ACT, CPR><<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, args, num_tiles);
  ES3_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

template <int BN, int STAGES, int CPR>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args, int tiles_m,
                  cudaStream_t stream) {
  switch (args.act) {
    case ACT_NONE: return launch_act<BN, STAGES, ACT_NONE, CPR>(tmA, tmB, args, tiles_m, stream);
    case ACT_RELU: return launch_act<BN, STAGES, ACT_RELU, CPR>(tmA, tmB, args, tiles_m, stream);
    case ACT_HSWISH: return launch_act<BN, STAGES, ACT_HSWISH, CPR>(tmA, tmB, args, tiles_m, stream);
    case ACT_GELU: return launch_act<BN, STAGES, ACT_GELU, CPR>(tmA, tmB, args, tiles_m, stream);
    default: set_error("gemm_tc: activation code %d not instantiated (0,1,2,3)", args.act); return 1;
  }
}

// Tile width: N need not be a multiple of BN (TMA zero-fills weight rows >= N, the epilogue masks 32-column
// chunks), so take 256 whenever the padding waste is small -- it halves the smem operand traffic per MMA.
static int pick_bn(int N, int bn_hint, int K = 1 << 30, int act = ACT_NONE) {
  if (bn_hint == 32 || bn_hint == 64 || bn_hint == 128 || bn_hint == 256) return bn_hint;
  // short-K GEMMs (the pointwise convs of the students: K = 128 / 256) are epilogue- and store-bound, and narrower tiles
  // (two CTAs per SM, more tiles in flight) win there: measured with scripts/gemm_bn_sweep.py on B200, e.g. K=128 N=512
  // 104 -> 85 us, K=256 N=1024 56 -> 49 us, K=256 N=128 (+residual) 48 -> 43 us, K=256 N=1024 + GELU 68 -> 61 us
  if (K < 512 && N >= 64) {
    if (act == ACT_GELU) return 64;
    return N >= 384 ? 128 : 64;
  }
  if (N >= 512 && (ceil_div(N, 256) * 256 - N) * 16 <= N) return 256;   // <= 6.25 % padded columns
  if (N % 128 == 0 || N > 1024) return 128;
  if (N % 64 == 0) return 64;
  return N >= 128 ? 128 : 32;
}

static int dispatch(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args, int tiles_m,
                    cudaStream_t stream) {
  switch (bn) {
    // (shared memory per CTA; two CTAs per SM need <= ~110 KB each)
    case 256: return launch<256, 4, 4>(tmA, tmB, args, tiles_m, stream);                        // 208 KB, one CTA per SM
    case 128: return args.num_kb <= 4 ? launch<128, 2, 4>(tmA, tmB, args, tiles_m, stream)      //  80 KB: short K, a tile is <= 4 k-blocks
                                      : launch<128, 3, 2>(tmA, tmB, args, tiles_m, stream);     // 104 KB: deep pipeline, 1 KB staging tiles
    case 64: return launch<64, 3, 4>(tmA, tmB, args, tiles_m, stream);                          //  88 KB
    default: return launch<32, 4, 4>(tmA, tmB, args, tiles_m, stream);                          //  96 KB
  }
}

}  // namespace es3

using namespace es3;

extern "C" int es3_gemm_bf16_ex(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                                int out_f32, int M, int N, int K, const float* scale, const float* bias, int act,
                                const void* residual, long long ldr, int res_f32, const float* rope, int rope_cols,
                                int rope_H, int rope_W, int rope_win, int act_after_res, int bn_hint, void* stream);

extern "C" int es3_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                             int out_f32, int M, int N, int K, const float* scale, const float* bias, int act,
                             const void* residual, long long ldr, int bn_hint, void* stream) {
  return es3_gemm_bf16_ex(A, lda, W, ldw, out, ldo, out_f32, M, N, K, scale, bias, act, residual, ldr, 0, nullptr, 0, 0,
                          0, 0, 0, bn_hint, stream);
}

extern "C" int es3_gemm_bf16_ex(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                                int out_f32, int M, int N, int K, const float* scale, const float* bias, int act,
                                const void* residual, long long ldr, int res_f32, const float* rope, int rope_cols,
                                int rope_H, int rope_W, int rope_win, int act_after_res, int bn_hint, void* stream) {
  ES3_REQUIRE(M > 0 && N > 0 && K > 0, "es3_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  ES3_REQUIRE(rope == nullptr || (rope_cols % 64 == 0 && rope_cols <= N && rope_H > 0 && rope_W > 0 &&
                                  M % (rope_H * rope_W) == 0 && ((uintptr_t)rope & 15) == 0),
              "es3_gemm_bf16_ex: bad rope arguments");
  ES3_REQUIRE(N % 8 == 0, "es3_gemm_bf16: N=%d must be a multiple of 8", N);
  ES3_REQUIRE(N % 32 == 0 || rope == nullptr, "es3_gemm_bf16_ex: the rope epilogue needs N %% 32 == 0 (N=%d)", N);
  ES3_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "es3_gemm_bf16: K/lda/ldw must be multiples of 8 (16-byte TMA strides)");
  ES3_REQUIRE(ldo % 8 == 0 && (residual == nullptr || ldr % 8 == 0), "es3_gemm_bf16: ldo/ldr must be multiples of 8");
  ES3_REQUIRE(rope == nullptr || act == ACT_NONE, "es3_gemm_bf16_ex: rope epilogue expects act = none");
  ES3_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0,
              "es3_gemm_bf16: pointers must be 16-byte aligned");
  // the RoPE epilogue is latency-heavy (table gathers): two narrower CTAs per SM hide it better (K=1024 N=3072: 809 -> 863 TF/s)
  const int bn = (rope != nullptr && bn_hint == 0 && N >= 128) ? 128 : pick_bn(N, bn_hint, K, act);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (encode_map(&tmA, A, 2, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)bn};
    if (encode_map(&tmB, W, 2, dims, str, box)) return 1;
  }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  { const char* e = getenv("ES3_GEMM_DBG"); if (e) a.dbg = atoi(e); }
  a.M = M; a.N = N;
  a.num_kb = ceil_div(K, BK);
  a.kb_per_tap = a.num_kb;
  a.tiles_n = ceil_div(N, bn);
  a.conv = 0;
  a.scale = scale; a.bias = bias; a.act = act;
  a.residual = (const bf16*)residual; a.ldr = ldr; a.res_f32 = res_f32;
  a.rope = (const float2*)rope; a.rope_cols = rope_cols; a.rope_H = rope_H; a.rope_W = rope_W; a.rope_win = rope_win;
  a.act_after_res = act_after_res;
  a.out = out; a.ldo = ldo; a.out_f32 = out_f32;
  return dispatch(bn, tmA, tmB, a, ceil_div(M, BM), (cudaStream_t)stream);
}

// Dense 3x3, stride 1, pad 1 conv on NHWC bf16 as an implicit GEMM.  x: [B,H,Wd,C]; W: [N][9*C] with
// k = (ky*3+kx)*C + c (repacked from the nn.Conv2d [N,C,3,3] weight by the Python side); out: [B,H,Wd,N].
extern "C" int es3_conv3x3_bf16(const void* x, const void* W, void* out, int out_f32, int B, int H, int Wd, int C,
                                int N, const float* scale, const float* bias, int act, const void* residual,
                                int bn_hint, void* stream) {
  ES3_REQUIRE(B > 0 && H > 0 && Wd > 0, "es3_conv3x3_bf16: bad shape");
  ES3_REQUIRE(C % 8 == 0 && N % 32 == 0, "es3_conv3x3_bf16: need C %% 8 == 0 and N %% 32 == 0 (C=%d N=%d)", C, N);
  const int bn = pick_bn(N, bn_hint);
  int TW = 8;
  if (Wd % 32 == 0) TW = 32; else if (Wd % 16 == 0) TW = 16;
  const int TH = BM / TW;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wd, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)Wd * C * 2, (uint64_t)H * Wd * C * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)TW, (uint32_t)TH, 1u};
    if (encode_map(&tmA, x, 4, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)9 * C * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)bn};
    if (encode_map(&tmB, W, 2, dims, str, box)) return 1;
  }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  { const char* e = getenv("ES3_GEMM_DBG"); if (e) a.dbg = atoi(e); }
  a.M = B * H * Wd; a.N = N;
  a.kb_per_tap = ceil_div(C, BK);
  a.num_kb = 9 * a.kb_per_tap;
  a.tiles_n = ceil_div(N, bn);
  a.conv = 1;
  a.H = H; a.W = Wd; a.TW = TW; a.TH = TH;
  a.tiles_w = ceil_div(Wd, TW); a.tiles_h = ceil_div(H, TH);
  a.Ctap = C;
  a.scale = scale; a.bias = bias; a.act = act;
  a.residual = (const bf16*)residual; a.ldr = N;
  a.out = out; a.ldo = N; a.out_f32 = out_f32;
  // NOTE: when C is not a multiple of 64 the last K block of a tap would read the next tap's weights
  // from W; the A side is zero-filled by TMA (channel coordinate beyond C), so the product is still 0.
  return dispatch(bn, tmA, tmB, a, B * a.tiles_w * a.tiles_h, (cudaStream_t)stream);
}

// ConvTranspose2d(kernel 2, stride 2) on NHWC: x [B,H,W,Cin] bf16, Wt [4*Cout][Cin] bf16 with
// Wt[(dy*2+dx)*Cout + co][ci] = w[ci][co][dy][dx]; bias4 [4*Cout] (the conv bias repeated per position) or NULL;
// out [B,2H,2W,Cout] bf16|fp32; residual (same layout as out, bf16|fp32) optional; act applied before the
// residual unless act_after_res.  Replaces nn.ConvTranspose2d in MaskDecoder.output_upscaling
// (mask_decoder.py:59-70, 213-216) and the FPN neck (necks.py).  Cout % 32 == 0.
extern "C" int es3_convt2x2_bf16(const void* x, const void* Wt, void* out, int out_f32, int B, int H, int Wd, int Cin,
                                 int Cout, const float* bias4, int act, const void* residual, int res_f32,
                                 int act_after_res, void* stream) {
  ES3_REQUIRE(Cout % 32 == 0 && Cin % 8 == 0, "es3_convt2x2_bf16: need Cout %% 32 == 0 and Cin %% 8 == 0 (Cin=%d Cout=%d)", Cin, Cout);
  const int M = B * H * Wd, N = 4 * Cout, K = Cin;
  const int bn = pick_bn(N, 0);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (encode_map(&tmA, x, 2, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)bn};
    if (encode_map(&tmB, Wt, 2, dims, str, box)) return 1;
  }
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  { const char* e = getenv("ES3_GEMM_DBG"); if (e) a.dbg = atoi(e); }
  a.M = M; a.N = N;
  a.num_kb = ceil_div(K, BK);
  a.kb_per_tap = a.num_kb;
  a.tiles_n = ceil_div(N, bn);
  a.bias = bias4; a.act = act;
  a.residual = (const bf16*)residual; a.ldr = Cout; a.res_f32 = res_f32;
  a.out = out; a.ldo = Cout; a.out_f32 = out_f32;
  a.ct_cout = Cout; a.ct_H = H; a.ct_W = Wd; a.act_after_res = act_after_res;
  return dispatch(bn, tmA, tmB, a, ceil_div(M, BM), (cudaStream_t)stream);
}
