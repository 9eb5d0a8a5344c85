// Flash attention on tcgen05 / TMEM for the SAM3 ViT trunk (head_dim 64), windowed or global.
//   O = softmax(Q K^T / 8) V            reference: F.scaled_dot_product_attention, vitdet.py:502
//
// One CTA = up to two 128-row query tiles ("slots") of one (image, window, head), ping-ponged so the
// exp-heavy softmax of one slot overlaps the MMAs of the other:
//   warp 0 / warp 10 : K / V producers -- cp.async row gather (windows are gathered in place, no
//                      window_partition copy) into 128B-swizzled smem tiles, 3-stage ring
//   warp 1           : TMEM allocator + single-thread tcgen05.mma issuer
//                        S_q  = Q_q K_j^T        (SS: A = Q smem, B = K smem, both K-major, 128x128x64)
//                        Ot_q = P_q V_j          (TS: A = P in TMEM (bf16), B = V smem MN-major, 128x64x128)
//   warps 2-5 / 6-9  : softmax of slot 0 / 1 -- one thread per query row: tcgen05.ld S (two passes: row max,
//                      then exp2), P written back to TMEM with tcgen05.st, per-tile Ot read from TMEM and folded
//                      into the fp32 running output in registers (o = o * corr + Ot), so O is never rescaled in TMEM.
// TMEM map (512 columns): slot q at q*256: S [0,128) fp32 | P [128,192) packed bf16x2 | Ot [192,256) fp32.
#include <cstdlib>

#include "ptx.cuh"

namespace es3 {

constexpr int FA_BM = 128, FA_D = 64, FA_STAGES = 3;   // (max) KV ring depth   // KV tile rows BN is a template parameter (128 or 96)
constexpr int FA_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct FaArgs {
  const bf16* qkv;
  bf16* out;
  int H, W, C, win, nwx, nwin, L;
  float scale_log2;
};

__device__ __forceinline__ long long fa_token_row(const FaArgs& a, int b, int wi, int l) {
  if (a.win == 0) return (long long)b * a.H * a.W + l;
  const int wy = wi / a.nwx, wx = wi % a.nwx;
  const int i = l / a.win, j = l % a.win;
  return (long long)b * a.H * a.W + (long long)(wy * a.win + i) * a.W + wx * a.win + j;
}

__device__ __forceinline__ void fa_cp16(uint32_t saddr, const void* g, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void fa_cp_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// 128-byte-swizzled placement of 16-byte chunk c of row r inside a 1024-byte-aligned tile (what TMA SWIZZLE_128B
// would write, and what the UMMA SW128 descriptors expect).
__device__ __forceinline__ uint32_t fa_sw(uint32_t tile, int r, int c) { return tile + r * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// V tile as an MN-major B operand: rows = kv (K dim), 64 d contiguous per row (128 B), SW128; 8-row atoms of
// 1024 B along K (SBO); a single 64-wide N chunk (LBO unused).  (cute::UMMA::make_umma_desc<Major::MN>.)
__device__ __forceinline__ uint64_t fa_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(FA_TILE_BYTES >> 4) << 16;  // LBO: next 64-wide N chunk (not used, N = 64)
  d |= (uint64_t)(1024u >> 4) << 32;          // SBO: next 8 kv rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// BN = keys per tile: 128 in general; 96 for 24x24 windows (576 = 6 x 96: no padded columns).
template <int FA_BN, int NS>
__global__ void __launch_bounds__((3 + 4 * NS) * 32, NS == 1 ? 2 : 1) attn_tc_kernel(const FaArgs a) {
  // NS query slots per CTA: 2 (one CTA per SM, the slots ping-pong) for long sequences; 1 (two CTAs per SM, 256 TMEM columns
  // each, 2-stage KV ring) for the 24x24 windows, where a CTA lives for only 6 key tiles and the start-up / drain of
  // a single resident CTA was not hidden (ncu: 31 K cycles per CTA, 21 % of samples waiting on MMA results)
  constexpr int STG = NS == 2 ? 3 : 2, THR = (3 + 4 * NS) * 32, VWARP = 2 + 4 * NS;
  extern __shared__ uint8_t fa_raw[];
  __shared__ __align__(8) uint64_t kv_full[FA_STAGES], kv_empty[FA_STAGES];
  __shared__ __align__(8) uint64_t s_full[2], p_full[2], ot_full[2], ot_free[2];
  __shared__ uint32_t tmem_holder;

  const uint32_t smem0 = (ptx::smem_u32(fa_raw) + 1023u) & ~1023u;
  const uint32_t u_q = smem0;                          // [2][128 x 128 B]
  const uint32_t u_kv = smem0 + NS * FA_TILE_BYTES;    // [stage][K | V]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z / a.nwin, wi = blockIdx.z % a.nwin;
  const int q0 = blockIdx.x * NS * FA_BM;              // first query row of slot 0
  const int nslots = (NS == 2 && q0 + FA_BM < a.L) ? 2 : 1;
  const int ld = 3 * a.C;
  const bf16* qbase = a.qkv + head * FA_D;
  const bf16* kbase = qbase + a.C;
  const bf16* vbase = qbase + 2 * a.C;
  const int ntiles = (a.L + FA_BN - 1) / FA_BN;

  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { ptx::mbar_init(&kv_full[s], 2); ptx::mbar_init(&kv_empty[s], 1); }
    for (int q = 0; q < 2; ++q) {
      ptx::mbar_init(&s_full[q], 1); ptx::mbar_init(&p_full[q], 4);
      ptx::mbar_init(&ot_full[q], 1); ptx::mbar_init(&ot_free[q], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_holder, 256 * NS);
  // Q tiles (both slots), gathered by everyone
  for (int i = tid; i < NS * FA_BM * 8; i += THR) {
    const int c = i & 7, r = (i >> 3) & 127, q = i >> 10;
    const int l = q0 + q * FA_BM + r;
    const bool ok = l < a.L;
    const long long row = ok ? fa_token_row(a, b, wi, l) : 0;
    fa_cp16(fa_sw(u_q + q * FA_TILE_BYTES, r, c), qbase + row * ld + c * 8, ok);
  }
  fa_cp_wait_all();
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0 || warp == VWARP) {
    // ------------------------------------------------------------------ K (warp 0) / V (warp 10) producers
    const bf16* base = (warp == 0) ? kbase : vbase;
    const uint32_t off = (warp == 0) ? 0 : FA_TILE_BYTES;
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < ntiles; ++j) {
      ptx::mbar_wait(&kv_empty[stage], phase ^ 1);
      const uint32_t tile = u_kv + stage * 2 * FA_TILE_BYTES + off;
#pragma unroll
      for (int k = 0; k < FA_BN / 32; ++k) {
        const int r = lane + 32 * k;
        const int l = j * FA_BN + r;
        const bool ok = l < a.L;
        const bf16* src = base + (ok ? fa_token_row(a, b, wi, l) : 0) * ld;
#pragma unroll
        for (int c = 0; c < 8; ++c) fa_cp16(fa_sw(tile, r, c), src + c * 8, ok);
      }
      fa_cp_wait_all();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&kv_full[stage]);
      if (++stage == STG) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16_f32(128, FA_BN);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16_f32(128, 64) | (1u << 16);  // B is MN-major
      auto issue_s = [&](int q, int stage) {
        const uint64_t dq = ptx::make_desc_sw128(u_q + q * FA_TILE_BYTES);
        const uint64_t dk = ptx::make_desc_sw128(u_kv + stage * 2 * FA_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) ptx::umma_f16(tmem + q * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
        ptx::umma_commit(&s_full[q]);
      };
      int stage = 0;
      uint32_t phase = 0;
      ptx::mbar_wait(&kv_full[0], 0);
      ptx::tc_fence_after();
      for (int q = 0; q < nslots; ++q) issue_s(q, 0);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t jp = (uint32_t)(j & 1);
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == STG) { nstage = 0; nphase ^= 1; }
        for (int q = 0; q < nslots; ++q) {
          ptx::mbar_wait(&p_full[q], jp);          // P_q(j) written, S_q(j) fully consumed
          ptx::mbar_wait(&ot_free[q], jp ^ 1);      // Ot_q(j-1) folded into registers
          ptx::tc_fence_after();
          const uint64_t dv = fa_desc_mn(u_kv + stage * 2 * FA_TILE_BYTES + FA_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < FA_BN / 16; ++k)
            umma_f16_ts(tmem + q * 256 + 192, tmem + q * 256 + 128 + k * 8, dv + (uint64_t)(k * 128), idesc_o, k != 0);
          ptx::umma_commit(&ot_full[q]);
          if (j + 1 < ntiles) {
            if (q == 0) { ptx::mbar_wait(&kv_full[nstage], nphase); ptx::tc_fence_after(); }
            issue_s(q, nstage);
          }
        }
        ptx::umma_commit(&kv_empty[stage]);          // K_j / V_j no longer needed once the MMAs above retire
        stage = nstage; phase = nphase;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (2..9)
    const int q = (warp - 2) >> 2;
    const int quarter = warp & 3;
    if (q < nslots) {
      const int row = quarter * 32 + lane;
      const uint32_t t_s = tmem + q * 256 + ((uint32_t)(quarter * 32) << 16);
      const uint32_t t_p = t_s + 128, t_o = t_s + 192;
      float o_reg[FA_D];
#pragma unroll
      for (int d = 0; d < FA_D; ++d) o_reg[d] = 0.f;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t jp = (uint32_t)(j & 1);
        ptx::mbar_wait(&s_full[q], jp);
        ptx::tc_fence_after();
        const int col0 = j * FA_BN;
        const bool full = col0 + FA_BN <= a.L;   // every column of this tile is a real key: no masking work
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < FA_BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_s + c * 32, v);
          ptx::tmem_ld_wait();
          if (full) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float sv = (col0 + c * 32 + e < a.L) ? __uint_as_float(v[e]) : -INFINITY;
              mx = fmaxf(mx, sv);
            }
          }
        }
        const float m_new = fmaxf(m_run, mx * a.scale_log2);
        const float corr = fa_exp2(m_run - m_new);
        float rs = 0.f;
#pragma unroll 1
        for (int c = 0; c < FA_BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_s + c * 32, v);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float p0 = fa_exp2(fmaf(__uint_as_float(v[2 * e]), a.scale_log2, -m_new));
            float p1 = fa_exp2(fmaf(__uint_as_float(v[2 * e + 1]), a.scale_log2, -m_new));
            if (!full) {
              const int cc = col0 + c * 32 + 2 * e;
              if (cc >= a.L) p0 = 0.f;
              if (cc + 1 >= a.L) p1 = 0.f;
            }
            rs += p0 + p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x16(t_p + c * 16, pk);
        }
        tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[q]);
        l_run = l_run * corr + rs;
        m_run = m_new;
        // fold this tile's P V into the running output
        ptx::mbar_wait(&ot_full[q], jp);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_o + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o_reg[c * 32 + e] = fmaf(o_reg[c * 32 + e], corr, __uint_as_float(v[e]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ot_free[q]);
      }
      const int l = q0 + q * FA_BM + row;
      if (l < a.L) {
        const float inv = 1.f / l_run;
        bf16* op = a.out + fa_token_row(a, b, wi, l) * a.C + head * FA_D;
#pragma unroll
        for (int d = 0; d < FA_D; d += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = o_reg[d + e] * inv;
          *reinterpret_cast<uint4*>(op + d) = pack8(f);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256 * NS);
  }
}


// ------------------------------------------------------------------------------------------ 24 x 24 windows, whole window resident
// The windowed blocks of the trunk (28 of 32: L = 576 keys, 4.5 query tiles) ran attn_tc_kernel<96, 1>: one CTA per QUERY TILE, so
// each (window, head) gathered its K and V five times, lived for only six key tiles and had no second slot to overlap softmax with
// the MMAs (round-1 table: 222-260 TFLOP/s against 430 for the global blocks).  Here one CTA owns a (window, head): Q (5 tiles,
// the last half empty), K and V (6 tiles of 96 rows each) are gathered ONCE into 224 KB of shared memory by all warps, then the
// five query tiles run through the two ping-ponged slots (pairs (0,1), (2,3), (4)) against the resident K / V -- no producer warps,
// no ring, the same S / P / Ot protocol with per-slot iteration counters as barrier phases.
constexpr int FW_L = 576, FW_BN = 96, FW_NKT = 6, FW_NQT = 5, FW_KT_BYTES = FW_BN * 128;   // 12288 B per K / V tile (12 atoms)
constexpr int FW_THREADS = 10 * 32;
constexpr int FW_SMEM = 1024 + FW_NQT * FA_TILE_BYTES + 2 * FW_NKT * FW_KT_BYTES;          // 1 KB + 80 KB + 144 KB

__global__ void __launch_bounds__(FW_THREADS, 1) attn_win24_tc_kernel(const FaArgs a) {
  extern __shared__ uint8_t fa_raw[];
  __shared__ __align__(8) uint64_t s_full[2], p_full[2], ot_full[2], ot_free[2];
  __shared__ uint32_t tmem_holder;
  const uint32_t smem0 = (ptx::smem_u32(fa_raw) + 1023u) & ~1023u;
  const uint32_t u_q = smem0;                                   // [5][128 x 128 B]
  const uint32_t u_k = smem0 + FW_NQT * FA_TILE_BYTES;          // [6][96 x 128 B]
  const uint32_t u_v = u_k + FW_NKT * FW_KT_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z / a.nwin, wi = blockIdx.z % a.nwin;
  const int ld = 3 * a.C;
  const bf16* qbase = a.qkv + head * FA_D;

  if (tid == 0) {
    for (int q = 0; q < 2; ++q) {
      ptx::mbar_init(&s_full[q], 1); ptx::mbar_init(&p_full[q], 4);
      ptx::mbar_init(&ot_full[q], 1); ptx::mbar_init(&ot_free[q], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(&tmem_holder, 512);
  // ---- gather the whole window once: Q rows 0..639 (>= 576: zero), K and V rows 0..575; 16-byte chunks, 128B-swizzled tiles
  for (int i = tid; i < (FW_NQT * FA_BM + 2 * FW_L) * 8; i += FW_THREADS) {
    const int c = i & 7, r = i >> 3;
    if (r < FW_NQT * FA_BM) {
      const bool ok = r < FW_L;
      const long long row = ok ? fa_token_row(a, b, wi, r) : 0;
      fa_cp16(fa_sw(u_q + (r >> 7) * FA_TILE_BYTES, r & 127, c), qbase + row * ld + c * 8, ok);
    } else {
      const int rr = r - FW_NQT * FA_BM;
      const int isv = rr >= FW_L;
      const int l = isv ? rr - FW_L : rr;
      const long long row = fa_token_row(a, b, wi, l);
      const int t = l / FW_BN, lr = l % FW_BN;
      fa_cp16(fa_sw((isv ? u_v : u_k) + t * FW_KT_BYTES, lr, c), qbase + (isv ? 2 : 1) * a.C + row * ld + c * 8, true);
    }
  }
  fa_cp_wait_all();
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_holder;
  constexpr int NPAIR = (FW_NQT + 1) / 2;

  if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16_f32(128, FW_BN);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16_f32(128, 64) | (1u << 16);  // B is MN-major
      auto issue_s = [&](int q, int qt, int kt) {
        const uint64_t dq = ptx::make_desc_sw128(u_q + qt * FA_TILE_BYTES);
        const uint64_t dk = ptx::make_desc_sw128(u_k + kt * FW_KT_BYTES);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k) ptx::umma_f16(tmem + q * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
        ptx::umma_commit(&s_full[q]);
      };
      uint32_t it[2] = {0u, 0u};                     // tiles processed per slot: the phase of its four barriers
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int nslots = (2 * pr + 1 < FW_NQT) ? 2 : 1;
        for (int q = 0; q < nslots; ++q) issue_s(q, 2 * pr + q, 0);
        for (int j = 0; j < FW_NKT; ++j) {
          for (int q = 0; q < nslots; ++q) {
            ptx::mbar_wait(&p_full[q], it[q] & 1u);            // P_q written, S_q consumed
            ptx::mbar_wait(&ot_free[q], (it[q] & 1u) ^ 1u);     // previous Ot_q folded into registers
            ptx::tc_fence_after();
            const uint64_t dv = fa_desc_mn(u_v + j * FW_KT_BYTES);
#pragma unroll
            for (int k = 0; k < FW_BN / 16; ++k)
              umma_f16_ts(tmem + q * 256 + 192, tmem + q * 256 + 128 + k * 8, dv + (uint64_t)(k * 128), idesc_o, k != 0);
            ptx::umma_commit(&ot_full[q]);
            if (j + 1 < FW_NKT) issue_s(q, 2 * pr + q, j + 1);
            ++it[q];
          }
        }
      }
    }
  } else if (warp >= 2) {
    const int q = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_s = tmem + q * 256 + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_p = t_s + 128, t_o = t_s + 192;
    uint32_t it = 0u;
    for (int pr = 0; pr < NPAIR; ++pr) {
      const int qt = 2 * pr + q;
      if (qt >= FW_NQT) break;
      float o_reg[FA_D];
#pragma unroll
      for (int d = 0; d < FA_D; ++d) o_reg[d] = 0.f;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < FW_NKT; ++j, ++it) {
        const uint32_t jp = it & 1u;
        ptx::mbar_wait(&s_full[q], jp);
        ptx::tc_fence_after();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < FW_BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_s + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
        }
        const float m_new = fmaxf(m_run, mx * a.scale_log2);
        const float corr = fa_exp2(m_run - m_new);
        float rs = 0.f;
#pragma unroll 1
        for (int c = 0; c < FW_BN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_s + c * 32, v);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = fa_exp2(fmaf(__uint_as_float(v[2 * e]), a.scale_log2, -m_new));
            const float p1 = fa_exp2(fmaf(__uint_as_float(v[2 * e + 1]), a.scale_log2, -m_new));
            rs += p0 + p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x16(t_p + c * 16, pk);
        }
        tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[q]);
        l_run = l_run * corr + rs;
        m_run = m_new;
        ptx::mbar_wait(&ot_full[q], jp);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_o + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o_reg[c * 32 + e] = fmaf(o_reg[c * 32 + e], corr, __uint_as_float(v[e]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ot_free[q]);
      }
      const int l = qt * FA_BM + row;
      if (l < FW_L) {
        const float inv = 1.f / l_run;
        bf16* op = a.out + fa_token_row(a, b, wi, l) * a.C + head * FA_D;
#pragma unroll
        for (int d = 0; d < FA_D; d += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = o_reg[d + e] * inv;
          *reinterpret_cast<uint4*>(op + d) = pack8(f);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace es3

using namespace es3;

// tcgen05 flash attention; same contract as es3_attention_bf16 (which dispatches here for L >= 128).
extern "C" int es3_attention_tc_bf16(const void* qkv, void* out, int B, int H, int W, int C, int num_heads, int win,
                                     float scale, void* stream) {
  ES3_REQUIRE(C == num_heads * FA_D, "es3_attention_tc_bf16: head_dim must be 64 (C=%d heads=%d)", C, num_heads);
  ES3_REQUIRE(win == 0 || (H % win == 0 && W % win == 0), "es3_attention_tc_bf16: H,W must be multiples of the window");
  FaArgs a;
  a.qkv = (const bf16*)qkv; a.out = (bf16*)out; a.H = H; a.W = W; a.C = C; a.win = win;
  a.nwx = win ? W / win : 1;
  a.nwin = win ? (H / win) * (W / win) : 1;
  a.L = win ? win * win : H * W;
  a.scale_log2 = scale * 1.4426950408889634f;
  const int ns = a.L <= 1024 ? 1 : 2;       // windows: one slot per CTA, two CTAs per SM; global attention: two slots
  const int smem = 1024 + ns * FA_TILE_BYTES + (ns == 2 ? 3 : 2) * 2 * FA_TILE_BYTES;
  const int threads = (3 + 4 * ns) * 32;
  static bool configured = false;
  if (!configured) {
    const int s2 = 1024 + 2 * FA_TILE_BYTES + 3 * 2 * FA_TILE_BYTES, s1 = 1024 + FA_TILE_BYTES + 2 * 2 * FA_TILE_BYTES;
    ES3_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    ES3_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<96, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    ES3_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    ES3_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<96, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (win == 24 && a.L == FW_L && !getenv("ES3_ATTN_WIN_TILED")) {     // the trunk's windowed blocks: whole window resident
    static bool cfg_win = false;
    if (!cfg_win) {
      ES3_CHECK_CUDA(cudaFuncSetAttribute(attn_win24_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM));
      cfg_win = true;
    }
    attn_win24_tc_kernel<<<dim3(1, num_heads, B * a.nwin), FW_THREADS, FW_SMEM, st>>>(a);
    ES3_LAUNCH_CHECK("attn_win24_tc_kernel");
    return 0;
  }
  dim3 grid(ceil_div(a.L, ns * FA_BM), num_heads, B * a.nwin);
  // 24x24 windows (L = 576 = 6 x 96) and other multiples of 96 that are not multiples of 128: 96-key tiles, no padding
  const bool bn96 = a.L % 96 == 0 && a.L % 128 != 0;
  if (ns == 1) {
    if (bn96) attn_tc_kernel<96, 1><<<grid, threads, smem, st>>>(a);
    else attn_tc_kernel<128, 1><<<grid, threads, smem, st>>>(a);
  } else {
    if (bn96) attn_tc_kernel<96, 2><<<grid, threads, smem, st>>>(a);
    else attn_tc_kernel<128, 2><<<grid, threads, smem, st>>>(a);
  }
  ES3_LAUNCH_CHECK("attn_tc_kernel");
  return 0;
}
