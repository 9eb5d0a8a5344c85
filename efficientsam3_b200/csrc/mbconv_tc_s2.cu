// Fused stride-2 MBConv on tcgen05 (the stage-opening blocks 16->64->32 @512^2, 32->128->64 @256^2, 64->256->128 @128^2 of
// efficientvit_b1: y = BN3(pw2(act(BN2(dw3x3_s2(act(BN1(pw1(x)))))))), no residual; reference efficientvit/nn/ops.py:315-367).
// Same machinery as mbconv_tc.cu (TMA-staged swizzled input tile, UMMA expand into TMEM, BN+act epilogue into a pixel-major
// smem tile, diagonal m16n8k8 depthwise into the swizzled A operand of the projecting UMMA that accumulates over chunks),
// with the geometry of a stride-2 block: a 4 x 16 output tile needs a 9 x 33 input tile = 297 pixels = three M=128 UMMA row
// blocks, so the expanded tensor is processed in 32-channel chunks to keep its tile (297 x 32 bf16) at 24 KB and two CTAs
// on an SM.  Project K-steps per chunk: 2 (32 channels); the W3 box still loads 64 K-columns (128-byte swizzled rows), the
// upper half is simply not referenced.
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
__device__ __forceinline__ void mma_1688(float* d, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace

constexpr int S2_TH = 4, S2_TW = 16;                       // output tile (64 pixels = rows 0..63 of the project UMMA)
constexpr int S2_IH = 2 * S2_TH + 1, S2_IW = 2 * S2_TW + 1;  // 9 x 33 input pixels
constexpr int S2_PIN = S2_IH * S2_IW;                      // 297
constexpr int S2_MC = 32;                                  // expanded-channel chunk
constexpr int S2_RS_MID = S2_MC * 2 + 16;                  // 80-byte rows: conflict-free 16-byte stores / ldmatrix
// s_mid keeps the EVEN and ODD input columns of every input row in separate planes (even column 2j -> slot j, odd column 2j+1 ->
// slot S2_ODD + j).  A stride-2 tap then walks CONSECUTIVE 80-byte rows (8 rows -> 8 distinct 4-bank groups); walking every other
// row of an interleaved tile (160-byte stride) made every ldmatrix 2-way conflicted: 25.8 M of 78 M load wavefronts at 16->64->32
// (profiles/r2s_ncu/r2s_fwd_raw.csv.gz).  S2_ODD = 20 also keeps the expand epilogue's 16-byte stores (lanes alternate planes) apart.
constexpr int S2_ODD = 20, S2_PW = S2_ODD + S2_TW;           // slots per input row: 17 even | 3 unused | 16 odd
constexpr int S2_THREADS = 288;

template <int MID, int COUT, int WSTAGES>
struct S2Smem {
  static constexpr int IN = (S2_PIN * 128 + 1023) / 1024 * 1024;    // 38912; rows 297..383 of the third M block alias what follows
  static constexpr int W1 = 2 * S2_MC * 128;
  static constexpr int W3_STAGE = COUT * 128;
  static constexpr int OFF_W1 = IN, OFF_W3 = OFF_W1 + W1, OFF_DW = OFF_W3 + WSTAGES * W3_STAGE, OFF_MID = OFF_DW + 128 * 128;
  static constexpr int OFF_WDW = OFF_MID + (S2_IH * S2_PW * S2_RS_MID + 15) / 16 * 16;   // bf16 [MID/32][9][32]
  static constexpr int OFF_PAR = OFF_WDW + 9 * MID * 2;
  static constexpr int OFF_BAR = OFF_PAR + (3 * MID + 2 * COUT) * 4;
  static constexpr int TOTAL = OFF_BAR + 128;
  static_assert(3 * 128 * 128 <= OFF_MID, "the aliased tail of the A operand must stay inside the operand buffers");
};

struct S2Args {
  bf16* y;             // [B,Ho,Wo,COUT]
  const float* s1;     // [MID]
  const float* b1;
  const float* wdw;    // [9][MID] fp32 (BN2 scale folded)
  const float* b2;
  const float* s3;     // [COUT]
  const float* b3;
  int H, W, Ho, Wo, tiles_x, tiles_y, total_tiles;
};

template <int CIN, int MID, int COUT, int WSTAGES, int ACT>
__global__ void __launch_bounds__(S2_THREADS, 2)
mbconv_tc_s2_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w1,
                    const __grid_constant__ CUtensorMap tm_w3, const S2Args a) {
  using L = S2Smem<MID, COUT, WSTAGES>;
  constexpr int NC = MID / S2_MC;
  constexpr int TCOLS = (96 + COUT <= 128) ? 128 : 256;    // 3 x 32 expand columns + COUT project columns
  static_assert(CIN % 16 == 0 && CIN <= 64 && MID % 32 == 0 && NC >= 2 && (COUT == 32 || COUT == 64 || COUT == 128), "shape");
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
  const int my_tiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_total = my_tiles * NC;

  if (tid == 0) {
    if (ptx::smem_u32(smem) & 1023u) { printf("es3: mbconv_tc_s2 dynamic smem base not 1024-byte aligned\n"); __trap(); }
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
  if (warp == 8) ptx::tmem_alloc(&tmem_holder, TCOLS);
  for (int i = tid; i < MID; i += S2_THREADS) { s_s1[i] = a.s1[i]; s_b1[i] = a.b1[i]; s_b2[i] = a.b2[i]; }
  for (int i = tid; i < COUT; i += S2_THREADS) { s_s3[i] = a.s3[i]; s_b3[i] = a.b3[i]; }
  for (int i = tid; i < 9 * MID; i += S2_THREADS) {      // bf16 depthwise weights, [chunk][tap][32]
    const int c = i % S2_MC, tap = (i / S2_MC) % 9, ch = i / (S2_MC * 9);
    const_cast<bf16*>(s_wdw)[i] = __float2bfloat16(a.wdw[tap * MID + ch * S2_MC + c]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t t_exp = tmem, t_proj = tmem + 96;
  const int tiles_per_img = a.tiles_x * a.tiles_y;

  if (warp == 8) {
    // ------------------------------------------------------------------------------------ control: TMA + UMMA issue
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t W1_BYTES = S2_MC * 128, W3_BYTES = L::W3_STAGE;
      constexpr uint32_t idesc_exp = ptx::make_idesc_bf16_f32(128, S2_MC);
      constexpr uint32_t idesc_proj = ptx::make_idesc_bf16_f32(128, COUT);
      const uint32_t u_in = ptx::smem_u32(s_in), u_dw = ptx::smem_u32(s_dw);
      auto load_in = [&](int t) {
        const int bb = t / tiles_per_img, r = t % tiles_per_img;
        ptx::mbar_arrive_expect_tx(bar_in, S2_PIN * 128);
        ptx::tma_load_4d(&tm_in, bar_in, s_in, 0, 2 * (r % a.tiles_x) * S2_TW - 1, 2 * (r / a.tiles_x) * S2_TH - 1, bb);
      };
      auto load_w1 = [&](int g) {
        const int s = g & 1;
        ptx::mbar_arrive_expect_tx(bar_w1 + s, W1_BYTES);
        ptx::tma_load_2d(&tm_w1, bar_w1 + s, s_w1 + s * W1_BYTES, 0, (g % NC) * S2_MC);
      };
      auto load_w3 = [&](int g) {
        const int s = (WSTAGES == 2) ? (g & 1) : 0;
        ptx::mbar_arrive_expect_tx(bar_w3 + s, W3_BYTES);
        ptx::tma_load_2d(&tm_w3, bar_w3 + s, s_w3 + s * W3_BYTES, (g % NC) * S2_MC, 0);
      };
      load_in((int)blockIdx.x);
      load_w1(0); load_w3(0); load_w1(1);
      if (WSTAGES == 2) load_w3(1);
      auto issue_expand = [&](int g) {
        const int it = g / NC, c = g % NC, st = g & 1;
        if (c == 0) ptx::mbar_wait(bar_in, (uint32_t)(it & 1));
        ptx::mbar_wait(bar_w1 + st, (uint32_t)((g >> 1) & 1));
        ptx::tc_fence_after();
        const uint64_t db1 = ptx::make_desc_sw128(ptx::smem_u32(s_w1 + st * W1_BYTES));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const uint64_t da = ptx::make_desc_sw128(u_in + d * 128 * 128);
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k)
            ptx::umma_f16(t_exp + d * S2_MC, da + (uint64_t)(k * 2), db1 + (uint64_t)(k * 2), idesc_exp, k != 0);
        }
        ptx::umma_commit(bar_exp);
        if (g >= 1 && g + 1 < n_total) load_w1(g + 1);   // stage last read by expand(g-1), retired before epilogue(g-1) ran
        if (c == NC - 1 && it + 1 < my_tiles) {            // s_in is free once this tile's last expand retires
          ptx::mbar_wait(bar_exp, (uint32_t)(g & 1));
          load_in((int)blockIdx.x + (it + 1) * (int)gridDim.x);
        }
      };
      issue_expand(0);
#pragma unroll 1
      for (int g = 0; g < n_total; ++g) {
        const int it = g / NC, c = g % NC;
        const int ws = (WSTAGES == 2) ? (g & 1) : 0;
        if (g + 1 < n_total) {                             // expand runs one chunk ahead of the depthwise
          ptx::mbar_wait(bar_expfree, (uint32_t)(g & 1));
          issue_expand(g + 1);
        }
        ptx::mbar_wait(bar_w3 + ws, (uint32_t)(WSTAGES == 2 ? ((g >> 1) & 1) : (g & 1)));
        ptx::mbar_wait(bar_dw, (uint32_t)(g & 1));
        if (WSTAGES == 2 && g >= 1 && g + 1 < n_total) load_w3(g + 1);   // project(g-1) retired (dw(g) waited for it)
        if (c == 0 && it > 0) ptx::mbar_wait(bar_projfree, (uint32_t)((it - 1) & 1));
        ptx::tc_fence_after();
        const uint64_t da = ptx::make_desc_sw128(u_dw);
        const uint64_t db3 = ptx::make_desc_sw128(ptx::smem_u32(s_w3 + ws * W3_BYTES));
#pragma unroll
        for (int k = 0; k < S2_MC / 16; ++k)
          ptx::umma_f16(t_proj, da + (uint64_t)(k * 2), db3 + (uint64_t)(k * 2), idesc_proj, (c | k) != 0);
        ptx::umma_commit(bar_proj);
        if (WSTAGES == 1 && g + 1 < n_total) {             // single W3 stage: reload as soon as this project has retired
          ptx::mbar_wait(bar_proj, (uint32_t)(g & 1));
          load_w3(g + 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ compute warps 0..7
    const int q = warp & 3, hsel = warp >> 2;
    const int g4 = lane >> 2, t4 = lane & 3;
    const int a_row = lane & 15, a_kh = lane >> 4;
    const uint32_t u_mid = ptx::smem_u32(s_mid);
    const uint32_t dshift = (g4 & 1) ? 16u : 0u;
    const bool dvalid = (g4 >> 1) == t4;
    int g = 0;

#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = t / tiles_per_img, tr = t % tiles_per_img;
      const int oy0 = (tr / a.tiles_x) * S2_TH, ox0 = (tr % a.tiles_x) * S2_TW;
      const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
#pragma unroll 1
      for (int c = 0; c < NC; ++c, ++g) {
        ptx::mbar_wait(bar_exp, (uint32_t)(g & 1));
        ptx::tc_fence_after();
        compute_bar_sync();                                // every warp is done reading s_mid for the previous chunk
        // ---- expand epilogue over the 297 input pixels: warp half 0 takes M blocks 0 and 2, half 1 takes block 1
#pragma unroll
        for (int dd = 0; dd < 2; ++dd) {
          const int d = hsel == 0 ? dd * 2 : 1;
          if (dd == 1 && (hsel == 1 || q >= 2)) break;     // block 2 holds rows 256..296 only (quarters 0, 1); warp-uniform
          const int row = d * 128 + q * 32 + lane;
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_exp + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * S2_MC), v);
          ptx::tmem_ld_wait();
          if (row < S2_PIN) {
            const int iy = iy0 + row / S2_IW, ix = ix0 + row % S2_IW;
            const bool in = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
            const float4* sc = reinterpret_cast<const float4*>(s_s1 + c * S2_MC);
            const float4* bi = reinterpret_cast<const float4*>(s_b1 + c * S2_MC);
            const int lx = row % S2_IW;
            uint4* dst = reinterpret_cast<uint4*>(s_mid + ((row / S2_IW) * S2_PW + ((lx & 1) ? S2_ODD + (lx >> 1) : (lx >> 1))) * S2_RS_MID);
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
        compute_bar_sync();                                // s_mid complete
        if (c > 0) ptx::mbar_wait(bar_proj, (uint32_t)((g - 1) & 1));   // project(g-1) has finished reading s_dw

        // ---- depthwise 3x3 stride 2 (diagonal m16n8k8 MMAs): warp -> channel group cg (16 of the 32), output row mt
        {
          const int cg = warp & 1, mt = warp >> 1;
          const bf16* wd = s_wdw + c * 9 * S2_MC + cg * 16 + g4;
          float dacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          // taps in pairs (0,1) (2,3) (4,5) (6,7) as m16n8k16 MMAs -- A = [tap-a fragment | tap-b fragment] along k, B = [diag(wa) ; diag(wb)]
          // -- and tap 8 as m16n8k8: 10 MMAs instead of 18 (k16 issues at the rate of k8, scripts/mma_bench.cu)
          auto frag = [&](int tap, uint32_t* af) {
            const int ky = tap / 3, kx = tap - ky * 3;
            // input column 2 * a_row + kx: kx = 0 / 2 -> even slots a_row / a_row + 1, kx = 1 -> odd slot a_row
            ldsm_x4(u_mid + ((2 * mt + ky) * S2_PW + a_row + (kx == 1 ? S2_ODD : (kx >> 1))) * S2_RS_MID + (cg * 16 + a_kh * 8) * 2,
                    af[0], af[1], af[2], af[3]);
          };
          auto bfrag = [&](int tap, uint32_t& b_lo, uint32_t& b_hi) {
            const uint32_t w_lo = (uint32_t)__bfloat16_as_ushort(wd[tap * S2_MC]);
            const uint32_t w_hi = (uint32_t)__bfloat16_as_ushort(wd[tap * S2_MC + 8]);
            b_lo = dvalid ? (w_lo << dshift) : 0u;
            b_hi = dvalid ? (w_hi << dshift) : 0u;
          };
#pragma unroll
          for (int tp = 0; tp < 4; ++tp) {
            uint32_t fa[4], fb[4], la, ha, lb, hb;
            frag(2 * tp, fa); frag(2 * tp + 1, fb);
            bfrag(2 * tp, la, ha); bfrag(2 * tp + 1, lb, hb);
            const uint32_t a_lo[4] = {fa[0], fa[1], fb[0], fb[1]};
            const uint32_t a_hi[4] = {fa[2], fa[3], fb[2], fb[3]};
            mma_16816(dacc[0], a_lo, la, lb);
            mma_16816(dacc[1], a_hi, ha, hb);
          }
          {
            uint32_t fa[4], la, ha;
            frag(8, fa); bfrag(8, la, ha);
            mma_1688(dacc[0], fa[0], fa[1], la);
            mma_1688(dacc[1], fa[2], fa[3], ha);
          }
          const float* b2 = s_b2 + c * S2_MC;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int p = mt * S2_TW + g4 + half * 8;      // output pixel = A-operand row (0..63) of the project UMMA
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const int ch = cg * 16 + nt * 8 + t4 * 2;
              const float2 bb = *reinterpret_cast<const float2*>(b2 + ch);
              const float v0 = es3_act_t<ACT>(dacc[nt][half * 2 + 0] + bb.x);
              const float v1 = es3_act_t<ACT>(dacc[nt][half * 2 + 1] + bb.y);
              const int j = cg * 2 + nt;                    // 16-byte chunk 0..3 of the 128-byte row, XOR-swizzled by row % 8
              *reinterpret_cast<uint32_t*>(s_dw + p * 128 + ((j ^ (p & 7)) << 4) + t4 * 4) = pack_bf16x2(v0, v1);
            }
          }
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_dw);
      }

      // ---- final epilogue: BN3 -> global.  Rows 0..63 of D_proj are the tile's pixels: lane quarters 0 and 1.
      constexpr int CW = COUT >= 64 ? COUT / 2 : COUT;     // columns per warp half
      const bool active = q < 2 && (COUT >= 64 || hsel == 0);
      const int col0 = COUT >= 64 ? hsel * CW : 0;
      const int r = q * 32 + lane;
      const int oy = oy0 + r / S2_TW, ox = ox0 + r % S2_TW;
      const bool inb = active && oy < a.Ho && ox < a.Wo;
      const long long pix = (((long long)b * a.Ho + oy) * a.Wo + ox) * COUT + col0;
      ptx::mbar_wait(bar_proj, (uint32_t)((g - 1) & 1));
      ptx::tc_fence_after();
      if (q < 2) {                                         // warp-uniform
#pragma unroll 1
        for (int cb = 0; cb < CW / 32; ++cb) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_proj + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + cb * 32), v);
          ptx::tmem_ld_wait();
          if (inb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                f[e] = fmaf(__uint_as_float(v[j * 8 + e]), s_s3[col0 + cb * 32 + j * 8 + e], s_b3[col0 + cb * 32 + j * 8 + e]);
              reinterpret_cast<uint4*>(a.y + pix + cb * 32)[j] = pack8(f);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_projfree);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, TCOLS);
  }
}

template <int CIN, int MID, int COUT, int WSTAGES>
static int launch_mbconv_tc_s2(const void* x, const void* w1, const void* w3, const S2Args& a, int B, cudaStream_t st) {
  using L = S2Smem<MID, COUT, WSTAGES>;
  CUtensorMap tm_in, tm_w1, tm_w3;
  {
    uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)CIN * 2, (uint64_t)a.W * CIN * 2, (uint64_t)a.H * a.W * CIN * 2};
    uint32_t box[4] = {64u, (uint32_t)S2_IW, (uint32_t)S2_IH, 1u};
    if (encode_map(&tm_in, x, 4, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)MID};
    uint64_t str[1] = {(uint64_t)CIN * 2};
    uint32_t box[2] = {64u, (uint32_t)S2_MC};
    if (encode_map(&tm_w1, w1, 2, dims, str, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)MID, (uint64_t)COUT};
    uint64_t str[1] = {(uint64_t)MID * 2};
    uint32_t box[2] = {64u, (uint32_t)COUT};
    if (encode_map(&tm_w3, w3, 2, dims, str, box)) return 1;
  }
  auto kern = mbconv_tc_s2_kernel<CIN, MID, COUT, WSTAGES, ACT_HSWISH>;
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
  const int ctas = a.total_tiles < 2 * sm_count ? a.total_tiles : 2 * sm_count;
  kern<<<ctas, S2_THREADS, L::TOTAL, st>>>(tm_in, tm_w1, tm_w3, a);
  ES3_LAUNCH_CHECK("mbconv_tc_s2_kernel");
  return 0;
}

}  // namespace es3

using namespace es3;

// Same contract as es3_mbconv_fused_bf16 for the stride-2, no-residual, hardswish blocks with (Cin, Mid, Cout) in
// {(16,64,32), (32,128,64), (64,256,128)}.  Returns -1 (no error set) for any other shape.
extern "C" int es3_mbconv_tc_s2_bf16(const void* x, void* y, const void* w1, const float* s1, const float* b1, const float* wdw,
                                     const float* b2, const void* w3, const float* s3, const float* b3, int B, int H, int W, int Cin,
                                     int Mid, int Cout, int stride, int residual, int act, void* stream) {
  const bool ok = stride == 2 && !residual && act == ACT_HSWISH &&
                  ((Cin == 16 && Mid == 64 && Cout == 32) || (Cin == 32 && Mid == 128 && Cout == 64) || (Cin == 64 && Mid == 256 && Cout == 128));
  if (!ok) return -1;
  ES3_REQUIRE(B > 0 && H > 0 && W > 0, "es3_mbconv_tc_s2_bf16: bad shape");
  ES3_REQUIRE((((uintptr_t)x | (uintptr_t)w1 | (uintptr_t)w3 | (uintptr_t)y) & 15) == 0, "es3_mbconv_tc_s2_bf16: 16-byte alignment");
  S2Args a;
  a.y = (bf16*)y; a.s1 = s1; a.b1 = b1; a.wdw = wdw; a.b2 = b2; a.s3 = s3; a.b3 = b3;
  a.H = H; a.W = W; a.Ho = (H - 1) / 2 + 1; a.Wo = (W - 1) / 2 + 1;
  a.tiles_x = ceil_div(a.Wo, S2_TW); a.tiles_y = ceil_div(a.Ho, S2_TH);
  a.total_tiles = B * a.tiles_x * a.tiles_y;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 16) return launch_mbconv_tc_s2<16, 64, 32, 2>(x, w1, w3, a, B, st);
  if (Cin == 32) return launch_mbconv_tc_s2<32, 128, 64, 2>(x, w1, w3, a, B, st);
  return launch_mbconv_tc_s2<64, 256, 128, 1>(x, w1, w3, a, B, st);
}
