// Two-query-tile flash attention for sm_100a (tcgen05 + TMEM + TMA): the long-sequence, small-head
// self-attention of the SD1.5 UNet (S = 4096 / 9216, head dim 40), SURVEY.md K5.  Replaces diffusers'
// AttnProcessor2_0 (F.scaled_dot_product_attention) in BasicTransformerBlock.attn1, reached from
// reference `backends/cuda_worker.py:222`.
//
// Why a second kernel: at head dim 40 the 128 x 128 score tile costs 384 tensor-core cycles
// (3 QK^T + 8 PV UMMAs) but 16384 exponentials = 1024 cycles of the 16-lane MUFU pipe, so the softmax
// instruction stream — not the MMA — bounds the kernel.  `attn_tc_kernel` (attention.cu) splits every
// query row over two threads (half-row max exchange through shared memory + a 256-thread named
// barrier per tile) and runs P over the S columns, so S_{j+1} cannot start before the exp pass of
// tile j is over; measured 26 % of the tensor roofline with the XU pipe 54 % busy.
//
// This kernel, NQ = 2 (one CTA per SM, 256 queries per CTA; NQ = 1: 128 queries per CTA, two CTAs per SM):
//   warp 0      TMA producer: Q (2 x 128 rows) once, K_j / V_j (128 keys) through an mbarrier ring,
//               fetched ONCE for 256 query rows
//   warp 1      MMA issuer: S_w = Q_w K_j^T (SS) and O_w += P_w V_j (TS, P from TMEM) for both
//               query tiles w = 0, 1
//   warps 2-5   softmax group 0, warps 6-9 softmax group 1: ONE thread per query row, the whole
//               128-key row of S in registers (one TMEM round trip), FMNMX3 max tree, packed
//               FFMA2 scale/shift, exponentials on MUFU.EX2 with kPoly of every 8 pairs on the FMA
//               pipe (packed Cody-Waite + cubic), P written as packed bf16 into its OWN TMEM
//               columns — so the tensor core computes S_w,j+1 as soon as group w holds S_w,j in
//               registers, i.e. under the exp pass, and a softmax group never waits for S.
//   TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512)
//   NQ = 1: one query tile per CTA, 256 columns (S [0,128) P [128,192) O [192,256)), two CTAs per SM with
//   their own K/V rings: the prologue (TMEM allocation, Q / K / V fetch, first S) and the epilogue of one CTA
//   run under the steady state of the other, and the two tiles are no longer coupled through one issuer.
// The softmax denominator is accumulated by the tensor core through V's ones column (as in
// attention.cu); running max is lazy (moved only when the tile max exceeds it by 2^8).
#include <stdlib.h>

#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int PP_TILE = 128;                      // queries per tile (two per CTA), keys per K/V tile
constexpr int PP_TILE_BYTES = PP_TILE * 128;      // [128 rows x 64 bf16] swizzled block

struct AttnPPParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  long long ldo;
  int sq, skv, d, dh_stride;
  int dv;            // PV MMA N (multiple of 16, <= 64): head dim + ones column, padded
  int l_col;         // O column holding the softmax denominator
  int stages;
  int smem_bytes;    // dynamic shared memory of the launch (checked against the carve-up on the device)
  float scale_log2;
};

__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// KS: QK^T K-steps of 16 (3: head dim 40, 4: head dim <= 64).  kPoly: of every 8 score pairs, this
// many take 2^x from the packed polynomial (FMA / ALU pipes) instead of MUFU.EX2.  NQ: query tiles per CTA.
template <int KS, int kPoly, int NQ>
__global__ void __launch_bounds__(64 + 128 * NQ, 3 - NQ)
attn_pp_kernel(const __grid_constant__ AttnPPParams p) {
  static_assert(NQ == 1 || NQ == 2, "one or two query tiles per CTA");
  constexpr uint32_t TMEM_COLS = 256u * NQ;
  constexpr uint32_t P_COL = 128u * NQ, O_COL = 192u * NQ;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 x 16 KB
  uint8_t* sKV = smem + NQ * PP_TILE_BYTES;             // stages x (K 16 KB | V 16 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + (size_t)p.stages * 2 * PP_TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* k_full = bars + 1;               // [stages]
  uint64_t* k_empty = k_full + p.stages;     // [stages]
  uint64_t* v_full = k_empty + p.stages;     // [stages]
  uint64_t* v_empty = v_full + p.stages;     // [stages]
  uint64_t* s_full = v_empty + p.stages;     // [2]  S_w,j computed
  uint64_t* s_taken = s_full + 2;            // [2]  group w holds S_w,j in registers
  uint64_t* p_ready = s_taken + 2;           // [2]  P_w,j written (and O_w rescaled)
  uint64_t* pv_done = p_ready + 2;           // [2]  O_w += P_w,j V_j retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  // the launch reserves < 1 KB of alignment slack (two CTAs of NQ = 1 must fit one SM): a dynamic segment
  // that starts less aligned than assumed becomes a launch error, not silent corruption
  if (threadIdx.x == 0 && reinterpret_cast<uint8_t*>(tmem_slot + 1) - smem_raw > p.smem_bytes) __trap();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * NQ * PP_TILE;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.skv + PP_TILE - 1) / PP_TILE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int w = 0; w < NQ; ++w) {
      mbar_init(&s_full[w], 1); mbar_init(&s_taken[w], 4);
      mbar_init(&p_ready[w], 4); mbar_init(&pv_done[w], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    const bool issuer = elect_one();
    const int col0 = h * p.dh_stride;
    if (issuer) {
      mbar_expect_tx(q_full, (uint32_t)(NQ * PP_TILE_BYTES));
      tma_load_2d(sQ, &p.tmQ, q_full, col0, b * p.sq + q0);
      if (NQ == 2) tma_load_2d(sQ + PP_TILE_BYTES, &p.tmQ, q_full, col0, b * p.sq + q0 + PP_TILE);
    }
    __syncwarp();
    auto load = [&](int t, bool is_v) {
      const int st = t % p.stages;
      uint64_t* empty = is_v ? &v_empty[st] : &k_empty[st];
      uint64_t* full = is_v ? &v_full[st] : &k_full[st];
      mbar_wait(empty, (uint32_t)(((t / p.stages) & 1) ^ 1));
      if (issuer) {
        uint8_t* dst = sKV + (size_t)st * 2 * PP_TILE_BYTES + (is_v ? PP_TILE_BYTES : 0);
        mbar_expect_tx(full, (uint32_t)PP_TILE_BYTES);
        tma_load_2d(dst, is_v ? &p.tmV : &p.tmK, full, col0, b * p.skv + t * PP_TILE);
      }
      __syncwarp();
    };
    // K runs one tile ahead of V (S_j+1 is issued a whole softmax pass before PV_j)
    load(0, false);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) load(j + 1, false);
      load(j, true);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const bool issuer = elect_one();
    const uint32_t idesc_s = umma_idesc_bf16(128, PP_TILE, 0, 0);          // Q K^T
    const uint32_t idesc_o = umma_idesc_bf16(128, (uint32_t)p.dv, 0, 1);   // P V (B MN-major)
    const uint32_t hi_k = umma_desc_hi_sw128(1024);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ));
    const uint32_t skv_addr = smem_u32(sKV);
    auto issue_s = [&](int w, int stage) {
      const uint32_t k_lo = umma_desc_lo(skv_addr + (uint32_t)(stage * 2 * PP_TILE_BYTES));
      const uint32_t qw_lo = q_lo + (uint32_t)(w * (PP_TILE_BYTES >> 4));
      if (issuer) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)          // 32 B = 2 sixteen-byte units per K-step inside the row
          umma_ss_lohi(tmem_base + (uint32_t)(w * PP_TILE), qw_lo + (uint32_t)(ks * 2),
                       k_lo + (uint32_t)(ks * 2), hi_k, idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(&s_full[w]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    if (NQ == 2) issue_s(1, 0);
    if (issuer) umma_commit(&k_empty[0]);
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_tiles; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
      if (j + 1 < n_tiles) {
        // S_w,j+1 as soon as group w has S_w,j in registers: it runs under the exp pass
        mbar_wait(&k_full[nstage], nphase);
        for (int w = 0; w < NQ; ++w) {
          mbar_wait(&s_taken[w], (uint32_t)(j & 1));
          tc_fence_after();
          issue_s(w, nstage);
        }
        if (issuer) umma_commit(&k_empty[nstage]);
        __syncwarp();
      }
      mbar_wait(&v_full[stage], phase);
      // V is the MN-major B operand: SBO = 8-key groups (1024 B); one 64-wide d chunk, LBO unused
      const uint32_t v_lo = umma_desc_lo(skv_addr + (uint32_t)(stage * 2 * PP_TILE_BYTES + PP_TILE_BYTES),
                                         (uint32_t)PP_TILE_BYTES);
      for (int w = 0; w < NQ; ++w) {
        mbar_wait(&p_ready[w], (uint32_t)(j & 1));
        tc_fence_after();
        if (issuer) {
          const uint32_t o_tmem = tmem_base + O_COL + (uint32_t)(w * 64);
          const uint32_t p_tmem = tmem_base + P_COL + (uint32_t)(w * 64);
#pragma unroll
          for (int ks = 0; ks < PP_TILE / 16; ++ks)   // 16 keys = two 8-row atoms = 2048 B; P: 8 packed columns
            umma_ts_lohi(o_tmem, p_tmem + (uint32_t)(ks * 8), v_lo + (uint32_t)(ks * 128), hi_k, idesc_o,
                         (ks > 0 || j > 0) ? 1u : 0u);
          umma_commit(&pv_done[w]);
        }
        __syncwarp();
      }
      if (issuer) umma_commit(&v_empty[stage]);
      __syncwarp();
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ============================ softmax + epilogue ============================
    const int w = (warp - 2) >> 2;                 // query tile / softmax group
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;                // row of the 128-row tile
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + (uint32_t)(w * PP_TILE);
    const uint32_t p_tmem = tmem_base + lane_off + P_COL + (uint32_t)(w * 64);
    const uint32_t o_tmem = tmem_base + lane_off + O_COL + (uint32_t)(w * 64);
    const float sc = p.scale_log2;
    const uint64_t sc2 = f2_pack(sc, sc);
    float m_run = -INFINITY;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[w], (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_ld32(s_tmem + (uint32_t)(c * 32), reinterpret_cast<uint32_t(&)[32]>(s[c * 32]));
      tmem_ld_wait();
      // S_w,j is in registers: the tensor core may overwrite the S columns with S_w,j+1
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_taken[w]);
      if (j * PP_TILE + PP_TILE > p.skv) {           // ragged last tile (warp-uniform)
        const int kbase = j * PP_TILE;
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kbase + i >= p.skv) s[i] = 0xff800000u;     // -inf: ignored by max, 2^x -> 0
      }
      float mx;
      {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          m0 = fmax3(m0, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          m1 = fmax3(m1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
          m2 = fmax3(m2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
          m3 = fmax3(m3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      // lazy running max: moved only when the tile max exceeds it by more than 2^8, so P <= 256 and
      // the O rescale is rare; the final O / l is unchanged because l (V's ones column) sees the same P
      const float m_tile = mx * sc;
      const float m_new = (m_tile > m_run + 8.0f) ? m_tile : m_run;
      const float corr = fast_exp2(m_run - m_new);
      m_run = m_new;
      const uint64_t nm2 = f2_pack(-m_new, -m_new);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[c * 32 + 2 * i]), __uint_as_float(s[c * 32 + 2 * i + 1])),
                                     sc2, nm2);
          float e0, e1;
          if (kPoly > 0 && (i & 7) >= 8 - kPoly) {
            // packed 2^x on the FMA / ALU pipes (see common.cuh::poly_exp2): x = n + f, |f| <= 0.5
            float x0, x1;
            f2_unpack(x2, x0, x1);
            const uint64_t xc = f2_pack(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
            const uint64_t r2 = f2_add(xc, f2_pack(12582912.0f, 12582912.0f));
            const uint64_t t2 = f2_add(r2, f2_pack(-12582912.0f, -12582912.0f));
            const uint64_t fr = f2_fma(t2, f2_pack(-1.0f, -1.0f), xc);
            uint64_t q2 = f2_fma(f2_pack(0.05508873611688614f, 0.05508873611688614f), fr,
                                 f2_pack(0.242604061961174f, 0.242604061961174f));
            q2 = f2_fma(q2, fr, f2_pack(0.6932762265205383f, 0.6932762265205383f));
            q2 = f2_fma(q2, fr, f2_pack(0.9999289512634277f, 0.9999289512634277f));
            float q0f, q1f, r0f, r1f;
            f2_unpack(q2, q0f, q1f);
            f2_unpack(r2, r0f, r1f);
            e0 = __int_as_float(__float_as_int(q0f) + (__float_as_int(r0f) << 23));
            e1 = __int_as_float(__float_as_int(q1f) + (__float_as_int(r1f) << 23));
          } else {
            float x0, x1;
            f2_unpack(x2, x0, x1);
            e0 = fast_exp2(x0);
            e1 = fast_exp2(x1);
          }
          pk[i] = pack_bf16x2(e0, e1);
        }
        if (c == 0 && j > 0) {
          // PV_w,j-1 reads the P columns this pass overwrites and writes the O it may rescale
          // (issued a whole softmax pass ago: this wait rarely spins)
          mbar_wait(&pv_done[w], (uint32_t)((j - 1) & 1));
          tc_fence_after();
        }
        tmem_st16(p_tmem + (uint32_t)(c * 16), pk);
      }
      if (j > 0 && __any_sync(0xffffffffu, corr != 1.0f)) {
#pragma unroll 1
        for (int c = 0; c < p.dv; c += 16) {
          uint32_t oo[16];
          tmem_ld16(o_tmem + (uint32_t)c, oo);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * corr);
          tmem_st16(o_tmem + (uint32_t)c, oo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[w]);
    }
    // epilogue: O / l  (l accumulated by the tensor core through V's ones column)
    mbar_wait(&pv_done[w], (uint32_t)((n_tiles - 1) & 1));
    tc_fence_after();
    float l_run;
    {
      uint32_t oo[16];
      tmem_ld16(o_tmem + (uint32_t)(p.l_col & ~15), oo);
      tmem_ld_wait();
      float lv = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.l_col & 15)) lv = __uint_as_float(oo[i]);
      l_run = lv;
    }
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + w * PP_TILE + r;
    const bool valid = qrow < p.sq;
    __nv_bfloat16* orow = p.out + ((long long)b * p.sq + qrow) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = 0; c < p.d; c += 16) {
      uint32_t oo[16];
      tmem_ld16(o_tmem + (uint32_t)c, oo);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (c + hh * 8 + 8 <= p.d) {
            uint4 ov;
            ov.x = pack_bf16x2(__uint_as_float(oo[hh * 8 + 0]) * inv_l, __uint_as_float(oo[hh * 8 + 1]) * inv_l);
            ov.y = pack_bf16x2(__uint_as_float(oo[hh * 8 + 2]) * inv_l, __uint_as_float(oo[hh * 8 + 3]) * inv_l);
            ov.z = pack_bf16x2(__uint_as_float(oo[hh * 8 + 4]) * inv_l, __uint_as_float(oo[hh * 8 + 5]) * inv_l);
            ov.w = pack_bf16x2(__uint_as_float(oo[hh * 8 + 6]) * inv_l, __uint_as_float(oo[hh * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + hh * 8) = ov;
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// -> 0 launched, 1 error, -1 shape not covered (caller falls through to attn_tc_kernel)
int attn_pp_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                   int dh_stride, void* out, long long ldo, int batch, int sq, int skv, int heads, int d,
                   float scale, int v_ones, cudaStream_t stream) {
  static int mode = -2;
  if (mode == -2) { const char* e = getenv("DL_ATTN_PP"); mode = e ? atoi(e) : 1; }
  const int d16 = (d + 15) / 16 * 16;
  const int dv = (d + 1 + 15) / 16 * 16;
  // one 64-column chunk per operand: QK^T depth 48 or 64 (compile-time K-steps), O width <= 64
  if (!mode || !v_ones || d % 8 || (d16 != 48 && d16 != 64) || dv > 64 || dh_stride < dv || dh_stride < d16 ||
      skv < 512 || sq < 256)
    return -1;
  if (dh_stride % 8 || ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return -1;
  AttnPPParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.sq = sq; p.skv = skv; p.d = d; p.dh_stride = dh_stride;
  p.dv = dv;
  p.l_col = d;
  // DL_ATTN_PP_NQ = query tiles per CTA: 1 (default: two CTAs per SM, each with its own K/V ring; measured on B200,
  // B = 16, S = 4096, d = 40: 690 us) or 2 (one CTA per SM, K/V fetched once for 256 rows: 716 us)
  static int nq = -2, poly = -2;
  if (nq == -2) { const char* e = getenv("DL_ATTN_PP_NQ"); nq = e ? atoi(e) : 1; }
  // share of the exponentials on the FMA pipe: DL_ATTN_PP_POLY = 0 / 2 / 3 of every 8 pairs
  if (poly == -2) { const char* e = getenv("DL_ATTN_PP_POLY"); poly = e ? atoi(e) : 2; }
  const int NQr = nq == 1 ? 1 : 2;
  const int overhead = 512 + 256;              // alignment slack + barriers
  const int budget = (NQr == 1 ? (228 * 1024) / 2 - 1024 : 200 * 1024) - overhead;
  p.stages = (budget - NQr * PP_TILE_BYTES) / (2 * PP_TILE_BYTES);
  if (p.stages > 4) p.stages = 4;
  DL_CHECK_ARG(p.stages >= 2, "attention(pp): not enough shared memory for 2 K/V stages");
  p.scale_log2 = scale * 1.4426950408889634f;
  const uint32_t box[2] = {64, PP_TILE};
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * sq};
    const uint64_t str[1] = {(uint64_t)ldq * 2};
    if (make_tmap_bf16(&p.tmQ, q, 2, dims, str, box)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * skv};
    const uint64_t str[1] = {(uint64_t)ldk * 2};
    if (make_tmap_bf16(&p.tmK, k, 2, dims, str, box)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * skv};
    const uint64_t str[1] = {(uint64_t)ldv * 2};
    if (make_tmap_bf16(&p.tmV, v, 2, dims, str, box)) return 1;
  }
  const int smem_bytes = NQr * PP_TILE_BYTES + p.stages * 2 * PP_TILE_BYTES + overhead;
  p.smem_bytes = smem_bytes;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaSuccess;
#define DL_PP_ATTR(...)                                                                         \
  if (e == cudaSuccess)                                                                         \
    e = cudaFuncSetAttribute(attn_pp_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
    DL_PP_ATTR(3, 0, 1); DL_PP_ATTR(3, 2, 1); DL_PP_ATTR(3, 3, 1);
    DL_PP_ATTR(4, 0, 1); DL_PP_ATTR(4, 2, 1); DL_PP_ATTR(4, 3, 1);
    DL_PP_ATTR(3, 0, 2); DL_PP_ATTR(3, 2, 2); DL_PP_ATTR(3, 3, 2);
    DL_PP_ATTR(4, 0, 2); DL_PP_ATTR(4, 2, 2); DL_PP_ATTR(4, 3, 2);
#undef DL_PP_ATTR
    if (e != cudaSuccess) { set_error("attention(pp): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set[dev & 63] = true;
  }
  dim3 grid((sq + NQr * PP_TILE - 1) / (NQr * PP_TILE), heads, batch);
  const int ks = d16 / 16;
  const int pl = poly <= 0 ? 0 : (poly >= 3 ? 3 : 2);
#define DL_PP_LAUNCH(KS_, PL_, NQ_) \
  attn_pp_kernel<KS_, PL_, NQ_><<<grid, 64 + 128 * NQ_, smem_bytes, stream>>>(p)
#define DL_PP_PICK(KS_, NQ_)                                                                   \
  do {                                                                                          \
    if (pl == 0) DL_PP_LAUNCH(KS_, 0, NQ_); else if (pl == 2) DL_PP_LAUNCH(KS_, 2, NQ_);        \
    else DL_PP_LAUNCH(KS_, 3, NQ_);                                                             \
  } while (0)
  if (ks == 3) { if (NQr == 1) DL_PP_PICK(3, 1); else DL_PP_PICK(3, 2); }
  else { if (NQr == 1) DL_PP_PICK(4, 1); else DL_PP_PICK(4, 2); }
#undef DL_PP_PICK
#undef DL_PP_LAUNCH
  return check_launch("attention(pp)");
}

}  // namespace dl
