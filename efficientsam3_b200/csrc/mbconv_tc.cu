// Fused MBConv on tcgen05 (stride 1, residual): y = x + BN3(pw2(act(BN2(dw3x3(act(BN1(pw1(x)))))))) with the 4x-expanded
// tensor kept on the SM (reference efficientvit/nn/ops.py:315-367 MBConv inside ResidualBlock :740-770).
//
// Why a second fused kernel (profiles/r1_mbconv_ncu.md): in mbconv_fused.cu all three contractions run on mma.sync, whose
// issue rate on sm_100a (~0.5 HMMA.16816 / cycle / SM) made the tensor pipe 35 % of the block's time, with every phase
// separated by __syncthreads.  Here the two pointwise GEMMs are UMMAs issued by one thread:
//
//   TMA (4-D map, halo + zero fill)  ->  s_in  [10x18 px][64 ch]  128B-swizzled K-major A operand
//   per 64-channel chunk of the expanded tensor:
//     expand   D_exp[2][128 x 64] (TMEM) = s_in (2 x M=128 rows) x W1c^T          tcgen05.mma, SS
//     epilogue tcgen05.ld -> BN1 + act (+ zero outside the image = the depthwise's padding) -> bf16 s_mid (pixel-major)
//     dw3x3    9 diagonal-B mma.sync per (16 px, 16 ch) from s_mid -> +bias, act -> bf16 s_dw, written in the
//              128B-swizzled K-major layout a UMMA A operand needs (fence.proxy.async before handing it over)
//     project  D_proj[128 x COUT] (TMEM) += s_dw x W3c^T                           tcgen05.mma, SS, accumulating over chunks
//   final     tcgen05.ld D_proj -> BN3 + residual (x re-read from global / L2) -> global
//
// Warps 0-7 do the elementwise / depthwise work, warp 8 lane 0 issues TMA and UMMAs; chunk weights sit in a 2-stage ring so
// expand(c+1) and the weight loads overlap the depthwise of chunk c.  TMEM: 128 (D_exp) + COUT (D_proj) <= 256 columns, so
// two CTAs share an SM.  CTAs are persistent: parameters, TMEM and barriers are set up once, the weight ring runs across
// tiles and the next tile's input TMA is issued as soon as the current tile's last expand retires.
#include <cuda.h>

#include "ptx.cuh"

namespace es3 {

int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b, const uint32_t* box);

namespace {
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// m16n8k8: a diagonal 8x8 B has no structural zeros to multiply (the k16 form wasted half of every MMA)
__device__ __forceinline__ void mma_1688(float* d, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace

constexpr int MT_TH = 8, MT_TW = 16;                    // output tile
constexpr int MT_HH = MT_TH + 2, MT_HW = MT_TW + 2;      // haloed input tile: 10 x 18
constexpr int MT_PIN = MT_HH * MT_HW;                    // 180 pixels
constexpr int MT_MC = 64;                                // expanded-channel chunk
constexpr int MT_RS_MID = MT_MC * 2 + 16;                // s_mid row stride (bytes)
constexpr int MT_THREADS = 288;                          // 8 compute warps + 1 control warp

template <int MID, int COUT>
struct MTSmem {
  static constexpr int IN = (MT_PIN * 128 + 1023) / 1024 * 1024;   // rows 180..255 of the 2 x M=128 A operand alias what follows
  static constexpr int W1 = 2 * MT_MC * 128;
  static constexpr int W3 = 2 * COUT * 128;
  static constexpr int DW = 128 * 128;
  static constexpr int MIDB = MT_PIN * MT_RS_MID;
  static constexpr int OFF_W1 = IN, OFF_W3 = OFF_W1 + W1, OFF_DW = OFF_W3 + W3, OFF_MID = OFF_DW + DW;
  static constexpr int OFF_WDW = OFF_MID + (MIDB + 15) / 16 * 16;   // bf16 [MID/64][9][64]
  static constexpr int OFF_PAR = OFF_WDW + 9 * MID * 2;              // fp32 s1[MID] b1[MID] b2[MID] s3[COUT] b3[COUT]
  static constexpr int OFF_BAR = OFF_PAR + (3 * MID + 2 * COUT) * 4;
  static constexpr int TOTAL = OFF_BAR + 128;   // 10 mbarriers
  static_assert(OFF_W1 + 76 * 128 <= OFF_MID, "the aliased tail of the A operand must stay inside the operand buffers");
};

struct MTArgs {
  const bf16* x;       // [B,H,W,CIN] (the TMA map reads it; the final epilogue re-reads the residual)
  bf16* y;             // [B,H,W,COUT]
  const float* s1;     // [MID]
  const float* b1;
  const float* wdw;    // [9][MID] fp32 (BN2 scale folded)
  const float* b2;     // [MID]
  const float* s3;     // [COUT]
  const float* b3;
  int H, W, tiles_x, tiles_y, total_tiles;
};

template <int CIN, int MID, int COUT, int ACT>
__global__ void __launch_bounds__(MT_THREADS, 2)
mbconv_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w3, const MTArgs a) {
  using L = MTSmem<MID, COUT>;
  constexpr int NC = MID / MT_MC;
  static_assert(CIN == COUT && CIN % 16 == 0 && CIN <= 64 && MID % 64 == 0 && NC >= 2 && (COUT == 32 || COUT == 64), "shape");
  // (no integer round trip on the pointer: the compiler must keep seeing shared-space addresses, or every access below
  //  turns into a generic LD/ST -- the first version of this kernel spent its time in long-scoreboard stalls on those)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_in = smem;
  uint8_t* s_w1 = smem + L::OFF_W1;
  uint8_t* s_w3 = smem + L::OFF_W3;
  uint8_t* s_dw = smem + L::OFF_DW;
  uint8_t* s_mid = smem + L::OFF_MID;
  const bf16* s_wdw = reinterpret_cast<const bf16*>(smem + L::OFF_WDW);
  float* s_par = reinterpret_cast<float*>(smem + L::OFF_PAR);
  float *s_s1 = s_par, *s_b1 = s_par + MID, *s_b2 = s_par + 2 * MID, *s_s3 = s_par + 3 * MID, *s_b3 = s_s3 + COUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t *bar_in = bars, *bar_w1 = bars + 1, *bar_w3 = bars + 3, *bar_exp = bars + 5, *bar_expfree = bars + 6,
           *bar_dw = bars + 7, *bar_proj = bars + 8, *bar_projfree = bars + 9;
  __shared__ uint32_t tmem_holder;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // persistent: this CTA owns tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the B * tiles_y * tiles_x tiles
  const int my_tiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_total = my_tiles * NC;                      // chunks this CTA will run (the weight ring spans tiles)

  if (tid == 0) {
    if (ptx::smem_u32(smem) & 1023u) { printf("es3: mbconv_tc dynamic smem base not 1024-byte aligned\n"); __trap(); }
    ptx::prefetch_tmap(&tm_in); ptx::prefetch_tmap(&tm_w1); ptx::prefetch_tmap(&tm_w3);
    ptx::mbar_init(bar_in, 1);
    ptx::mbar_init(bar_w1, 1); ptx::mbar_init(bar_w1 + 1, 1);
    ptx::mbar_init(bar_w3, 1); ptx::mbar_init(bar_w3 + 1, 1);
    ptx::mbar_init(bar_exp, 1);
    ptx::mbar_init(bar_expfree, 8);
    ptx::mbar_init(bar_dw, 8);
    ptx::mbar_init(bar_proj, 1);
    ptx::mbar_init(bar_projfree, 8);
    ptx::fence_mbar_init();
  }
  if (warp == 8) ptx::tmem_alloc(&tmem_holder, 256);
  // per-channel parameters and the depthwise weights (bf16, [chunk][tap][64]) -- once per CTA
  for (int i = tid; i < MID; i += MT_THREADS) { s_s1[i] = a.s1[i]; s_b1[i] = a.b1[i]; s_b2[i] = a.b2[i]; }
  for (int i = tid; i < COUT; i += MT_THREADS) { s_s3[i] = a.s3[i]; s_b3[i] = a.b3[i]; }
  for (int i = tid; i < 9 * MID; i += MT_THREADS) {
    const int c = i % 64, tap = (i / 64) % 9, ch = i / (64 * 9);
    const_cast<bf16*>(s_wdw)[i] = __float2bfloat16(a.wdw[tap * MID + ch * 64 + c]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t t_exp = tmem, t_proj = tmem + 128;
  const int tiles_per_img = a.tiles_x * a.tiles_y;

  if (warp == 8) {
    // ------------------------------------------------------------------------------------ control: TMA + UMMA issue
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t W1_BYTES = MT_MC * 128, W3_BYTES = COUT * 128;
      constexpr uint32_t idesc_exp = ptx::make_idesc_bf16_f32(128, MT_MC);
      constexpr uint32_t idesc_proj = ptx::make_idesc_bf16_f32(128, COUT);
      const uint32_t u_in = ptx::smem_u32(s_in), u_dw = ptx::smem_u32(s_dw);
      auto load_in = [&](int t) {
        const int bb = t / tiles_per_img, r = t % tiles_per_img;
        ptx::mbar_arrive_expect_tx(bar_in, MT_PIN * 128);
        ptx::tma_load_4d(&tm_in, bar_in, s_in, 0, (r % a.tiles_x) * MT_TW - 1, (r / a.tiles_x) * MT_TH - 1, bb);
      };
      auto load_w1 = [&](int gc) {
        const int s = gc & 1;
        ptx::mbar_arrive_expect_tx(bar_w1 + s, W1_BYTES);
        ptx::tma_load_2d(&tm_w1, bar_w1 + s, s_w1 + s * W1_BYTES, 0, (gc % NC) * MT_MC);
      };
      auto load_w3 = [&](int gc) {
        const int s = gc & 1;
        ptx::mbar_arrive_expect_tx(bar_w3 + s, W3_BYTES);
        ptx::tma_load_2d(&tm_w3, bar_w3 + s, s_w3 + s * W3_BYTES, (gc % NC) * MT_MC, 0);
      };
      load_in((int)blockIdx.x);
      load_w1(0); load_w3(0); load_w1(1); load_w3(1);
      // expand(g): needs the tile's input (first chunk), the W1 stage and a drained D_exp (the caller has waited for that)
      auto issue_expand = [&](int g) {
        const int it = g / NC, c = g % NC, st = g & 1;
        if (c == 0) ptx::mbar_wait(bar_in, (uint32_t)(it & 1));
        ptx::mbar_wait(bar_w1 + st, (uint32_t)((g >> 1) & 1));
        ptx::tc_fence_after();
        const uint64_t db1 = ptx::make_desc_sw128(ptx::smem_u32(s_w1 + st * W1_BYTES));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t da = ptx::make_desc_sw128(u_in + half * 128 * 128);
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k)
            ptx::umma_f16(t_exp + half * MT_MC, da + (uint64_t)(k * 2), db1 + (uint64_t)(k * 2), idesc_exp, k != 0);
        }
        ptx::umma_commit(bar_exp);
        // the W1 stage of chunk g+1 was last read by expand(g-1), which retired before epilogue(g-1) ran: refill it now
        if (g >= 1 && g + 1 < n_total) load_w1(g + 1);
        if (c == NC - 1 && it + 1 < my_tiles) {
          // last expand of this tile: once it retires s_in is free -> prefetch the next tile's input under the remaining
          // depthwise / project / final epilogue (the residual is re-read from global, not from s_in)
          ptx::mbar_wait(bar_exp, (uint32_t)(g & 1));
          load_in((int)blockIdx.x + (it + 1) * (int)gridDim.x);
        }
      };
      issue_expand(0);
#pragma unroll 1
      for (int g = 0; g < n_total; ++g) {
        const int it = g / NC, c = g % NC, st = g & 1;
        if (g + 1 < n_total) {
          // expand runs ONE chunk ahead: as soon as epilogue(g) has drained D_exp, expand(g+1) is issued, so it overlaps the
          // depthwise of chunk g instead of sitting behind project(g) (first version: 15 % of samples waiting on bar_exp)
          ptx::mbar_wait(bar_expfree, (uint32_t)(g & 1));
          issue_expand(g + 1);
        }
        ptx::mbar_wait(bar_w3 + st, (uint32_t)((g >> 1) & 1));
        ptx::mbar_wait(bar_dw, (uint32_t)(g & 1));        // s_dw(g) written (and, transitively, project(g-1) retired)
        if (g >= 1 && g + 1 < n_total) load_w3(g + 1);    // ... so the W3 stage of chunk g+1 is free as well
        if (c == 0 && it > 0) ptx::mbar_wait(bar_projfree, (uint32_t)((it - 1) & 1));   // D_proj drained by the last epilogue
        ptx::tc_fence_after();
        const uint64_t da = ptx::make_desc_sw128(u_dw);
        const uint64_t db3 = ptx::make_desc_sw128(ptx::smem_u32(s_w3 + st * W3_BYTES));
#pragma unroll
        for (int k = 0; k < MT_MC / 16; ++k)
          ptx::umma_f16(t_proj, da + (uint64_t)(k * 2), db3 + (uint64_t)(k * 2), idesc_proj, (c | k) != 0);
        ptx::umma_commit(bar_proj);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ compute warps 0..7
    const int q = warp & 3, hsel = warp >> 2;            // TMEM lane quarter; column half (expand epilogue) / m-tile parity (dw)
    const int g = lane >> 2, t4 = lane & 3;
    const int a_row = lane & 15, a_kh = lane >> 4;
    const uint32_t u_mid = ptx::smem_u32(s_mid);
    const uint32_t dshift = (g & 1) ? 16u : 0u;
    const bool dvalid = (g >> 1) == t4;
    int gc = 0;

#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = t / tiles_per_img, tr = t % tiles_per_img;
      const int oy0 = (tr / a.tiles_x) * MT_TH, ox0 = (tr % a.tiles_x) * MT_TW;
      const int iy0 = oy0 - 1, ix0 = ox0 - 1;
#pragma unroll 1
      for (int c = 0; c < NC; ++c, ++gc) {
        ptx::mbar_wait(bar_exp, (uint32_t)(gc & 1));
        ptx::tc_fence_after();
        compute_bar_sync();                              // every warp is done reading s_mid for the previous chunk
        // ---- expand epilogue: BN1 + act, zero outside the image, bf16 -> s_mid[row][hsel*32 .. +32)
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          if (d == 1 && q >= 2) break;                   // rows 192..255 of the second M=128 block are padding (warp-uniform)
          const int row = d * 128 + q * 32 + lane;
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_exp + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * MT_MC + hsel * 32), v);
          ptx::tmem_ld_wait();
          if (row < MT_PIN) {
            const int iy = iy0 + row / MT_HW, ix = ix0 + row % MT_HW;
            const bool in = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
            const float4* sc = reinterpret_cast<const float4*>(s_s1 + c * MT_MC + hsel * 32);   // warp-uniform: LDS.128 broadcasts
            const float4* bi = reinterpret_cast<const float4*>(s_b1 + c * MT_MC + hsel * 32);
            uint4* dst = reinterpret_cast<uint4*>(s_mid + row * MT_RS_MID + hsel * 64);
            if (in) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 s0 = sc[2 * j], s1v = sc[2 * j + 1], b0 = bi[2 * j], b1v = bi[2 * j + 1];
                const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1v.x, s1v.y, s1v.z, s1v.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = es3_act_t<ACT>(fmaf(__uint_as_float(v[j * 8 + e]), sv[e], bv[e]));
                dst[j] = pack8(f);
              }
            } else {                                       // outside the image: the depthwise's zero padding (border tiles only,
#pragma unroll                                             // so interior tiles carry no per-element select)
              for (int j = 0; j < 4; ++j) dst[j] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_expfree);
        compute_bar_sync();                              // s_mid complete
        if (c > 0) ptx::mbar_wait(bar_proj, (uint32_t)((gc - 1) & 1));   // project(gc-1) has finished reading s_dw

        // ---- depthwise 3x3 on tensor cores (diagonal-B MMAs): warp -> channel group cg (16 ch), output rows hsel*4 .. hsel*4+3.
        // Four CONSECUTIVE output rows share their input rows: per kx the warp loads the 6 input-row fragments once and feeds
        // them to the 3 ky taps of each output row -- 18 ldmatrix.x4 per chunk instead of 36 (the kernel's busiest unit is the
        // shared-memory pipe, 54 % LSU + 9 % UMMA wavefronts, profiles/r2s_ncu_evm_forward.md; ldmatrix was 2/3 of its loads).
        {
          const int cg = q;
          const bf16* wd = s_wdw + c * 9 * 64 + cg * 16 + g;
          float dacc[4][2][4];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int i = 0; i < 2; ++i) { dacc[m][i][0] = dacc[m][i][1] = dacc[m][i][2] = dacc[m][i][3] = 0.f; }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            uint32_t af[6][4];
#pragma unroll
            for (int r = 0; r < 6; ++r)
              ldsm_x4(u_mid + ((hsel * 4 + r) * MT_HW + a_row + kx) * MT_RS_MID + (cg * 16 + a_kh * 8) * 2, af[r][0], af[r][1], af[r][2], af[r][3]);
            // two taps per MMA: m16n8k16 issues at the rate of m16n8k8 (0.46 / clk / SM, scripts/mma_bench.cu), so the taps (0, kx) and
            // (1, kx) share one instruction -- A = [tap-0 fragment | tap-1 fragment] along k, B = [diag(w0) ; diag(w1)] -- and (2, kx)
            // stays a k8 MMA: 12 MMAs per 16 pixels x 16 channels instead of 18
            uint32_t b_lo[3], b_hi[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t w_lo = (uint32_t)__bfloat16_as_ushort(wd[(ky * 3 + kx) * 64]);
              const uint32_t w_hi = (uint32_t)__bfloat16_as_ushort(wd[(ky * 3 + kx) * 64 + 8]);
              b_lo[ky] = dvalid ? (w_lo << dshift) : 0u;
              b_hi[ky] = dvalid ? (w_hi << dshift) : 0u;
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const uint32_t a_lo[4] = {af[m][0], af[m][1], af[m + 1][0], af[m + 1][1]};
              const uint32_t a_hi[4] = {af[m][2], af[m][3], af[m + 1][2], af[m + 1][3]};
              mma_16816(dacc[m][0], a_lo, b_lo[0], b_lo[1]);                   // channels cg*16 + 0..7, taps ky = 0, 1
              mma_16816(dacc[m][1], a_hi, b_hi[0], b_hi[1]);                   // channels cg*16 + 8..15
              mma_1688(dacc[m][0], af[m + 2][0], af[m + 2][1], b_lo[2]);       // tap ky = 2
              mma_1688(dacc[m][1], af[m + 2][2], af[m + 2][3], b_hi[2]);
            }
          }
          const float* b2 = s_b2 + c * MT_MC;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int mt = hsel * 4 + m;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int p = mt * MT_TW + g + half * 8;     // output pixel = A-operand row of the project UMMA
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                const int ch = cg * 16 + nt * 8 + t4 * 2;
                const float2 bb = *reinterpret_cast<const float2*>(b2 + ch);
                const float v0 = es3_act_t<ACT>(dacc[m][nt][half * 2 + 0] + bb.x);
                const float v1 = es3_act_t<ACT>(dacc[m][nt][half * 2 + 1] + bb.y);
                const int j = cg * 2 + nt;                  // 16-byte chunk inside the 128-byte row; XOR-swizzled by row % 8
                *reinterpret_cast<uint32_t*>(s_dw + p * 128 + ((j ^ (p & 7)) << 4) + t4 * 4) = pack_bf16x2(v0, v1);
              }
            }
          }
        }
        ptx::fence_proxy_async();                         // generic-proxy writes -> visible to the UMMA (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_dw);
      }

      // ---- final epilogue: BN3 + residual -> global.  Thread = output pixel q*32+lane, channel half hsel (COUT 64) or all (COUT 32).
      const bool active = (COUT == 64) || (hsel == 0);
      const int col0 = (COUT == 64) ? hsel * 32 : 0;
      const int r = q * 32 + lane;
      const int oy = oy0 + r / MT_TW, ox = ox0 + r % MT_TW;
      const bool inb = active && oy < a.H && ox < a.W;
      const long long pix = (((long long)b * a.H + oy) * a.W + ox) * COUT + col0;
      uint4 xres[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)                         // residual x (L2-resident: TMA just read it), issued before the wait
        xres[j] = inb ? __ldg(reinterpret_cast<const uint4*>(a.x + pix) + j) : make_uint4(0u, 0u, 0u, 0u);
      ptx::mbar_wait(bar_proj, (uint32_t)((gc - 1) & 1));
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_proj + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_projfree);
      if (inb) {
        bf16* dst = a.y + pix;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float xr[8], f[8];
          unpack8(xres[j], xr);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = fmaf(__uint_as_float(v[j * 8 + e]), s_s3[col0 + j * 8 + e], s_b3[col0 + j * 8 + e]) + xr[e];
          reinterpret_cast<uint4*>(dst)[j] = pack8(f);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

template <int CIN, int MID, int COUT>
static int launch_mbconv_tc(const void* x, void* y, const void* w1, const void* w3, const MTArgs& a, int B, cudaStream_t st) {
  using L = MTSmem<MID, COUT>;
  CUtensorMap tm_in, tm_w1, tm_w3;
  {
    uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)CIN * 2, (uint64_t)a.W * CIN * 2, (uint64_t)a.H * a.W * CIN * 2};
    uint32_t box[4] = {64u, (uint32_t)MT_HW, (uint32_t)MT_HH, 1u};
    if (encode_map(&tm_in, x, 4, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)MID};
    uint64_t str[1] = {(uint64_t)CIN * 2};
    uint32_t box[2] = {64u, (uint32_t)MT_MC};
    if (encode_map(&tm_w1, w1, 2, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)MID, (uint64_t)COUT};
    uint64_t str[1] = {(uint64_t)MID * 2};
    uint32_t box[2] = {(uint32_t)MT_MC, (uint32_t)COUT};
    if (encode_map(&tm_w3, w3, 2, dims, str, box)) return 1;
  }
  auto kern = mbconv_tc_kernel<CIN, MID, COUT, ACT_HSWISH>;
  static bool configured = false;
  if (!configured) {
    ES3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    ES3_CHECK_CUDA(cudaGetDevice(&dev));
    ES3_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int ctas = a.total_tiles < 2 * sm_count ? a.total_tiles : 2 * sm_count;   // persistent, two CTAs per SM
  kern<<<ctas, MT_THREADS, L::TOTAL, st>>>(tm_in, tm_w1, tm_w3, a);
  ES3_LAUNCH_CHECK("mbconv_tc_kernel");
  return 0;
}

}  // namespace es3

using namespace es3;

// Same contract as es3_mbconv_fused_bf16 for the stride-1 residual blocks (Cin == Cout in {32, 64}, Mid = 4 Cin, hardswish).
// Returns -1 (no error set) for any other shape.
extern "C" int es3_mbconv_tc_bf16(const void* x, void* y, const void* w1, const float* s1, const float* b1, const float* wdw,
                                  const float* b2, const void* w3, const float* s3, const float* b3, int B, int H, int W, int Cin,
                                  int Mid, int Cout, int stride, int residual, int act, void* stream) {
  if (!(stride == 1 && residual && act == ACT_HSWISH && Cin == Cout && Mid == 4 * Cin && (Cin == 32 || Cin == 64))) return -1;
  ES3_REQUIRE(B > 0 && H > 0 && W > 0, "es3_mbconv_tc_bf16: bad shape");
  ES3_REQUIRE((((uintptr_t)x | (uintptr_t)w1 | (uintptr_t)w3 | (uintptr_t)y) & 15) == 0, "es3_mbconv_tc_bf16: 16-byte alignment");
  MTArgs a;
  a.x = (const bf16*)x; a.y = (bf16*)y; a.s1 = s1; a.b1 = b1; a.wdw = wdw; a.b2 = b2; a.s3 = s3; a.b3 = b3;
  a.H = H; a.W = W; a.tiles_x = ceil_div(W, MT_TW); a.tiles_y = ceil_div(H, MT_TH);
  a.total_tiles = B * a.tiles_x * a.tiles_y;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32) return launch_mbconv_tc<32, 128, 32>(x, y, w1, w3, a, B, st);
  return launch_mbconv_tc<64, 256, 64>(x, y, w1, w3, a, B, st);
}
