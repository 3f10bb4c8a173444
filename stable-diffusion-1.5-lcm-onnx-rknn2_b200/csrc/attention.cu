// Fused flash-style attention for sm_100a (tcgen05 + TMEM + TMA), SURVEY.md K5/K6.
// Replaces diffusers' Attention / AttnProcessor2_0 (F.scaled_dot_product_attention) inside
// BasicTransformerBlock.attn1 (self) and .attn2 (cross, S_kv = 77), reached from reference
// `backends/cuda_worker.py:222`.
//
// One CTA per (128-query tile, head, batch element); 6 warps:
//   warp 0  TMA producer   Q once, then K_j / V_j tiles (128 keys) through an mbarrier ring
//   warp 1  MMA issuer     S_j = Q K_j^T  (SS, M=128 N=128 K=d16)   -> TMEM S[j&1]
//                          O  += P_j V_j  (TS: A = P_j in TMEM, B = V_j MN-major smem)
//   warps 2-9 softmax      two threads per query row (64 key columns each): tcgen05.ld, online softmax in fp32
//                          (exp2, scale*log2e folded), P_j written back to TMEM as packed bf16
//                          over the S_j columns it was read from, O rescaled in TMEM when the
//                          running max moved; final O / l -> bf16 global.
// S is double-buffered so the tensor core computes S_{j+1} while the softmax warps work on
// S_j.  Head dims that are not a multiple of 16 (SD1.5: 40) are laid out by the QKV projection
// with a zero-padded per-head stride (dh_stride = 48), so the padded K-steps contribute 0.
#include <stdlib.h>

#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

int attn_simt_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                     long long ldv, int dh_stride, void* out, long long ldo, int batch, int sq,
                     int skv, int heads, int d, float scale, int causal, cudaStream_t stream);
// attention_x.cu: short key lists (cross-attention, S_kv <= 128): one CTA per (image, head) and query range;
// -1 = shape not covered
int attn_x_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                  int dh_stride, void* out, long long ldo, int batch, int sq, int skv, int heads, int d,
                  float scale, int v_ones, cudaStream_t stream);
// attention_pp.cu: two query tiles per CTA, one softmax thread per row (long keys, head dim <= 56);
// -1 = shape not covered
int attn_pp_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                   int dh_stride, void* out, long long ldo, int batch, int sq, int skv, int heads, int d,
                   float scale, int v_ones, cudaStream_t stream);

constexpr int AT_THREADS = 320;               // TMA warp, MMA warp, 8 softmax warps
constexpr int AT_TILE = 128;                 // queries per CTA and keys per KV tile
constexpr int AT_CHUNK_BYTES = AT_TILE * 128;   // one [128 rows x 64 bf16] swizzled block

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  long long ldo;
  int sq, skv, d, dh_stride;
  int ksteps;        // QK^T K-steps of 16 (ceil(d/16))
  int nchunk_qk;     // 64-column TMA boxes per Q / K tile
  int dv;            // PV MMA N = O columns (multiple of 16; includes the ones column if any)
  int nchunk_v;      // 64-column TMA boxes per V tile
  int l_col;         // O column that accumulates the softmax denominator (V ones column), or -1
  int kt;            // keys per KV tile (128 / 96 / 64): S columns, PV depth, TMA box rows
  int stages;        // K/V smem ring depth
  int sbuf;          // S/P TMEM buffers (2: S_{j+1} overlaps softmax_j; 1: two CTAs per SM)
  int pcol;          // > 0: P lives in its own TMEM columns [pcol, pcol + kt/2) instead of over S, so
                     // S_{j+1} is computed while the softmax warps are still in the exp pass of tile j
  int tmem_cols;
  long long* trace;  // debug: per-tile clock stamps of CTA (0,0,0), or NULL
  float scale_log2;
};

// kOnes: V carries a ones column, the PV MMA accumulates the softmax denominator (product path).
// A compile-time flag: as a runtime branch the unused row-sum code still cost ~190 predicated
// instructions per thread and tile in the exp loop.
// KS / KT: compile-time QK^T K-steps and key tile (0 = runtime values from the params).  The MMA
// warp shares its scheduler with four busy softmax warps, so every instruction it issues costs
// several cycles: with KS/KT known the issue paths are straight runs of UMMAs with folded
// descriptor offsets instead of 12 / 8 predicated slots with runtime address math (measured
// ~660 / ~330 cycles per tile for the S / PV issue before).
// kPoly: of every 8 score pairs, this many take their exponentials from `poly_exp2` (FMA / ALU
// pipes) instead of MUFU.EX2 — the exp pass of the d = 40 / 64 / 80 heads is bound by the 16-lane XU
// pipe (one MUFU per score), not by the tensor core (opt-in: DL_ATTN_POLY).
template <bool kOnes, bool kBf16Exp, int KS, int KT, int kPoly = 0>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ AttnParams p) {
  constexpr int SW = 8;                              // softmax warps: two threads per query row
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ float xch[2][2][AT_TILE];              // [tile parity][half][row] max / sum exchange
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int q_bytes = p.nchunk_qk * AT_CHUNK_BYTES;
  const int kchunk = p.kt * 128;                          // one [kt rows x 64 bf16] swizzled block
  const int k_bytes = p.nchunk_qk * kchunk;
  const int v_bytes = p.nchunk_v * kchunk;
  const int kv_bytes = k_bytes + v_bytes;
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + (size_t)p.stages * kv_bytes);
  uint64_t* q_full = bars;                 // 1
  // K and V slots are tracked separately: K_j is dead as soon as S_j is computed, V_j only after
  // PV_j — with a shared barrier K_{j+1} could not be fetched before PV_{j-1} had retired
  uint64_t* k_full = bars + 1;             // [stages]
  uint64_t* k_empty = k_full + p.stages;   // [stages]
  uint64_t* v_full = k_empty + p.stages;   // [stages]
  uint64_t* v_empty = v_full + p.stages;   // [stages]
  uint64_t* s_full = v_empty + p.stages;   // [2]
  uint64_t* p_ready = s_full + 2;          // [2]
  uint64_t* pv_done = p_ready + 2;         // 1
  uint64_t* s_taken = pv_done + 1;         // 1: the softmax warps hold S_j in registers (pcol mode)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_taken + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_TILE;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.skv + p.kt - 1) / p.kt;
  // TMEM: S/P buffers first (then the separate P buffer), then O
  const uint32_t o_col = (uint32_t)(p.pcol > 0 ? p.pcol + (p.kt >> 1) : p.sbuf * p.kt);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], SW); }
    mbar_init(pv_done, 1);
    mbar_init(s_taken, SW);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    // warp-uniform loop, one lane issues (see igemm.cu: a loop on a single divergent lane pays
    // ~100 cycles of R2UR round trips per TMA / MMA issue)
    const bool issuer = elect_one();   // elect.sync: the compiler keeps the issue path uniform
    const int col0 = h * p.dh_stride;
    if (issuer) {
      mbar_expect_tx(q_full, (uint32_t)q_bytes);
      for (int c = 0; c < p.nchunk_qk; ++c)
        tma_load_2d(sQ + c * AT_CHUNK_BYTES, &p.tmQ, q_full, col0 + c * 64, b * p.sq + q0);
    }
    __syncwarp();
    // K runs one tile ahead of V: K_{j+1}'s slot is free as soon as S_{j-1} is computed, while
    // V_j's slot waits for PV_{j-2} — in a strictly alternating order every K load queued
    // behind a V wait and arrived a whole tile late (TMA latency under load ~3k cycles)
    auto load_k = [&](int t) {
      const int st = t % p.stages;
      mbar_wait(&k_empty[st], (uint32_t)(((t / p.stages) & 1) ^ 1));
      if (issuer) {
        uint8_t* sK = sKV + (size_t)st * kv_bytes;
        mbar_expect_tx(&k_full[st], (uint32_t)k_bytes);
        for (int c = 0; c < p.nchunk_qk; ++c)
          tma_load_2d(sK + c * kchunk, &p.tmK, &k_full[st], col0 + c * 64, b * p.skv + t * p.kt);
      }
      __syncwarp();
    };
    load_k(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) load_k(j + 1);
      const int st = j % p.stages;
      mbar_wait(&v_empty[st], (uint32_t)(((j / p.stages) & 1) ^ 1));
      if (issuer) {
        uint8_t* sV = sKV + (size_t)st * kv_bytes + k_bytes;
        mbar_expect_tx(&v_full[st], (uint32_t)v_bytes);
        for (int c = 0; c < p.nchunk_v; ++c)
          tma_load_2d(sV + c * kchunk, &p.tmV, &v_full[st], col0 + c * 64, b * p.skv + j * p.kt);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const bool issuer = elect_one();   // elect.sync: the compiler keeps the issue path uniform
    const uint32_t idesc_s = umma_idesc_bf16(128, (uint32_t)p.kt, 0, 0);  // Q K^T
    const uint32_t idesc_o = umma_idesc_bf16(128, (uint32_t)p.dv, 0, 1);  // P V (B MN-major)
    const uint32_t hi_k = umma_desc_hi_sw128(1024);                       // K-major operands
    const uint32_t sq_addr = smem_u32(sQ);
    const uint32_t skv_addr = smem_u32(sKV);
    const int n_ks = KS > 0 ? KS : p.ksteps;
    const int kchunk16 = (KT > 0 ? KT * 128 : kchunk) >> 4;
    const uint32_t q_lo = umma_desc_lo(sq_addr);
    auto issue_s = [&](int j, int stage) {
      const uint32_t sk_addr = skv_addr + (uint32_t)(stage * kv_bytes);
      const int sb = j % p.sbuf;
      const uint32_t d_tmem = tmem_base + (uint32_t)(sb * p.kt);
      const uint32_t k_lo = umma_desc_lo(sk_addr);
      if (issuer) {
#pragma unroll
        for (int ks = 0; ks < (KS > 0 ? KS : 12); ++ks) {   // d <= 192: at most 12 K-steps of 16
          if (ks < n_ks) {
            // next 64-wide chunk every 4 steps (16 KB = 1024 x 16 B), 32 B = 2 units inside a row
            const uint32_t qoff = (uint32_t)((ks >> 2) * (AT_CHUNK_BYTES >> 4) + (ks & 3) * 2);
            const uint32_t koff = (uint32_t)((ks >> 2) * kchunk16 + (ks & 3) * 2);
            umma_ss_lohi(d_tmem, q_lo + qoff, k_lo + koff, hi_k, idesc_s, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(&k_empty[stage]);
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    const bool prefetch_s = (p.sbuf == 2 && p.stages >= 2);
    // single S buffer but >= 2 smem stages: PV_j and S_{j+1} go out as ONE burst of UMMAs right
    // after p_ready (K_{j+1} is waited for, and every operand computed, before that wait) — the
    // in-order tensor pipe keeps S_{j+1} behind PV_j, which reads the P it would overwrite
    const bool fused = (p.sbuf == 1 && p.stages >= 2 && p.pcol == 0);
    // separate P buffer: S_{j+1} goes out as soon as the softmax warps have S_j in registers
    // (s_taken), i.e. it runs under their exp pass; PV_j follows once P_j is written.  The pipe
    // is in order, so PV_j (reads P, V_j) and S_{j+2} (writes S) never race.
    const bool split = (p.pcol > 0);
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_s(0, 0);
    int stage = 0;
    uint32_t phase = 0;
    const int pv_steps = KT > 0 ? KT / 16 : (p.kt >> 4);
    const uint32_t o_tmem = tmem_base + o_col;
    for (int j = 0; j < n_tiles; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
      const bool more = j + 1 < n_tiles;
      if (prefetch_s && more) {
        mbar_wait(&k_full[nstage], nphase);
        tc_fence_after();
        issue_s(j + 1, nstage);
      }
      const int sb = j % p.sbuf;
      if (split && more) {
        const bool tr2 = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 16 && issuer;
        mbar_wait(&k_full[nstage], nphase);
        if (tr2) p.trace[j * 16 + 12] = clock64();
        mbar_wait(s_taken, (uint32_t)(j & 1));
        tc_fence_after();
        if (tr2) p.trace[j * 16 + 2] = clock64();
        issue_s(j + 1, nstage);
        if (tr2) p.trace[j * 16 + 3] = clock64();
      }
      // operands of PV_j (and of S_{j+1} in the fused burst), ready before the wait
      const uint32_t sv_addr = skv_addr + (uint32_t)(stage * kv_bytes + k_bytes);
      const uint32_t p_tmem = tmem_base + (uint32_t)(split ? p.pcol : sb * p.kt);
      const uint32_t s_next = tmem_base + (uint32_t)(sb * p.kt);
      // V is the MN-major B operand: LBO = stride between 64-wide d chunks, SBO = 8-key groups
      const uint32_t v_lo = umma_desc_lo(sv_addr, (uint32_t)(KT > 0 ? KT * 128 : kchunk));
      const uint32_t nk_lo = umma_desc_lo(skv_addr + (uint32_t)(nstage * kv_bytes));
      const uint32_t acc0 = j > 0 ? 1u : 0u;
      if (fused && more) mbar_wait(&k_full[nstage], nphase);
      mbar_wait(&v_full[stage], phase);
      mbar_wait(&p_ready[sb], (uint32_t)((j / p.sbuf) & 1));
      tc_fence_after();
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 16 && issuer;
      if (tr) p.trace[j * 16 + 0] = clock64();
      if (issuer) {
#pragma unroll
        for (int ks = 0; ks < (KT > 0 ? KT / 16 : AT_TILE / 16); ++ks) {
          // 16 keys = two 8-row swizzle atoms = 2048 B (+128 in 16-byte units); P advances 8
          // packed columns
          if (ks < pv_steps)
            umma_ts_lohi(o_tmem, p_tmem + (uint32_t)(ks * 8), v_lo + (uint32_t)(ks * 128), hi_k, idesc_o,
                         ks > 0 ? 1u : acc0);
        }
        umma_commit(&v_empty[stage]);
        umma_commit(pv_done);
        if (fused && more) {
#pragma unroll
          for (int ks = 0; ks < (KS > 0 ? KS : 12); ++ks) {
            if (ks < n_ks) {
              const uint32_t qoff = (uint32_t)((ks >> 2) * (AT_CHUNK_BYTES >> 4) + (ks & 3) * 2);
              const uint32_t koff = (uint32_t)((ks >> 2) * kchunk16 + (ks & 3) * 2);
              umma_ss_lohi(s_next, q_lo + qoff, nk_lo + koff, hi_k, idesc_s, ks > 0 ? 1u : 0u);
            }
          }
          umma_commit(&k_empty[nstage]);
          umma_commit(&s_full[sb]);
        }
      }
      __syncwarp();
      if (tr) p.trace[j * 16 + 1] = clock64();
      if (!prefetch_s && !fused && !split && more) {
        // one smem stage: K_{j+1} can only land after PV_j released the stage
        mbar_wait(&k_full[nstage], nphase);
        tc_fence_after();
        if (tr) p.trace[j * 16 + 2] = clock64();
        issue_s(j + 1, nstage);
        if (tr) p.trace[j * 16 + 3] = clock64();
      }
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ============================ softmax + epilogue ============================
    // 8 warps: two per TMEM lane quadrant; a query row is shared by two threads, each owning
    // 64 of the tile's 128 key columns (halves the serial TMEM-load -> exp chain per tile).
    const int quad = warp & 3;
    const int ch = (warp - 2) >> 2;                  // column half owned by this thread
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + o_col;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j % p.sbuf;
      const uint32_t s_tmem = tmem_base + lane_off + (uint32_t)(sb * p.kt);
      if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 16 && threadIdx.x == 64)
        p.trace[j * 16 + 11] = clock64();
      mbar_wait(&s_full[sb], (uint32_t)((j / p.sbuf) & 1));
      tc_fence_after();
      const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 16 &&
                      threadIdx.x == 64;
      if (tr) p.trace[j * 16 + 4] = clock64();
      // my columns: 64 / 48 / 32 (compile-time when KT is)
      const int hc = KT > 0 ? KT / 2 : (p.kt >> 1);
      const int kbase = j * p.kt + ch * hc;
      const bool need_mask = (j * p.kt + p.kt > p.skv);         // only the last tile (warp-uniform)
      // my S columns -> registers with ONE TMEM round trip (loads in flight together); they
      // serve the max pass and the exp pass, so S is never re-read
      uint32_t sa[32], sb32[32];
      tmem_ld32(s_tmem + (uint32_t)(ch * hc), sa);
      if (hc == 64) {
        tmem_ld32(s_tmem + (uint32_t)(ch * hc + 32), sb32);
      } else if (hc == 48) {
        uint32_t t16[16];
        tmem_ld16(s_tmem + (uint32_t)(ch * hc + 32), t16);
#pragma unroll
        for (int i = 0; i < 16; ++i) { sb32[i] = t16[i]; sb32[16 + i] = 0xff800000u; }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) sb32[i] = 0xff800000u;     // -inf: ignored by max, exp2 -> 0
      }
      tmem_ld_wait();
      if (p.pcol > 0) {
        // S_j is in registers: the tensor core may overwrite the S columns with S_{j+1}
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_taken);
      }
      if (tr) p.trace[j * 16 + 5] = clock64();
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (kbase + i >= p.skv) sa[i] = 0xff800000u;          // -inf: exp2 -> 0, max ignores
          if (kbase + 32 + i >= p.skv) sb32[i] = 0xff800000u;
        }
      }
      float mx;
      {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(sa[i]), __uint_as_float(sb32[i])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(sa[i + 1]), __uint_as_float(sb32[i + 1])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(sa[i + 2]), __uint_as_float(sb32[i + 2])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(sa[i + 3]), __uint_as_float(sb32[i + 3])));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      // exchange the half-row maxima; after this barrier every thread of the CTA has its S
      // values in registers, so P may overwrite the S columns right away
      xch[j & 1][ch][r] = mx;
      if (tr) p.trace[j * 16 + 6] = clock64();
      named_bar_sync(1, 256);
      if (tr) p.trace[j * 16 + 7] = clock64();
      mx = fmaxf(mx, xch[j & 1][ch ^ 1][r]);
      // lazy running max: only move it when the tile max exceeds it by more than 2^8 — P then
      // stays <= 256 (exact in bf16/fp32 range), the O rescale becomes rare instead of
      // per-tile, and the final O / l is unchanged because l sees the same P
      const float m_tile = mx * p.scale_log2;
      const float m_new = (m_tile > m_run + 8.0f) ? m_tile : m_run;
      const float corr = fast_exp2(m_run - m_new);
      float lsum0 = 0.f, lsum1 = 0.f;
      const uint32_t pw_tmem = p.pcol > 0 ? tmem_base + lane_off + (uint32_t)p.pcol : s_tmem;
      if (p.pcol > 0 && j > 0) {
        // PV_{j-1} reads the P buffer this pass is about to overwrite (long finished in practice:
        // it was issued a whole softmax pass ago)
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c == 1 && hc == 32) break;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (c == 1 && hc == 48 && i >= 8) { pk[i] = 0u; continue; }
          const uint32_t u0 = c == 0 ? sa[2 * i] : sb32[2 * i];
          const uint32_t u1 = c == 0 ? sa[2 * i + 1] : sb32[2 * i + 1];
          const float x0 = fmaf(__uint_as_float(u0), p.scale_log2, -m_new);
          const float x1 = fmaf(__uint_as_float(u1), p.scale_log2, -m_new);
          __nv_bfloat162 pb;
          if (kPoly > 0 && (i & 7) >= 8 - kPoly) {
            pb = __floats2bfloat162_rn(poly_exp2(x0), poly_exp2(x1));
          } else if (kBf16Exp) {
            // packed exp: bf16 results without a conversion (still one MUFU per value in SASS); P is
            // rounded to bf16 for the PV MMA anyway, and the denominator comes from the same P through
            // V's ones column
            const __nv_bfloat162 xb = __floats2bfloat162_rn(x0, x1);
            const uint32_t e = ex2_bf16x2(*reinterpret_cast<const uint32_t*>(&xb));
            pb = *reinterpret_cast<const __nv_bfloat162*>(&e);
          } else {
            pb = __floats2bfloat162_rn(fast_exp2(x0), fast_exp2(x1));
          }
          if (!kOnes) {
            // no ones column in V: sum what the tensor core will multiply (bf16-rounded P)
            const float2 pf = __bfloat1622float2(pb);
            lsum0 += pf.x;
            lsum1 += pf.y;
          }
          pk[i] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        if (c == 0 || hc == 64) {
          tmem_st16(pw_tmem + (uint32_t)(ch * (hc >> 1) + c * 16), pk);
        } else if (hc == 48) {                                  // 16 more columns -> 8 packed
          uint32_t p8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) p8[i] = pk[i];
          tmem_st8(pw_tmem + (uint32_t)(ch * (hc >> 1) + 16), p8);
        }
      }
      l_run = l_run * corr + (lsum0 + lsum1);
      m_run = m_new;
      if (tr) p.trace[j * 16 + 8] = clock64();
      tmem_st_wait();
      if (tr) p.trace[j * 16 + 9] = clock64();
      if (j > 0) {
        // O must hold PV_{j-1} before it is rescaled.  PV_j accumulating on top needs no wait
        // here: the tensor pipe is in-order and S_j (already consumed above) was issued after
        // PV_{j-1}, so pv_done's phase j-1 has completed whenever s_full's phase j has — the
        // wait below can only ever spin when the rescale is actually taken.
        if (__any_sync(0xffffffffu, corr != 1.0f)) {
          mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
          tc_fence_after();
#pragma unroll 1
          for (int c = ch * 16; c < p.dv; c += SW * 4) {
            uint32_t oo[16];
            tmem_ld16(o_tmem + (uint32_t)c, oo);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * corr);
            tmem_st16(o_tmem + (uint32_t)c, oo);
          }
          tmem_st_wait();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[sb]);
      if (tr) p.trace[j * 16 + 10] = clock64();
    }
    // epilogue: O / l
    mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
    tc_fence_after();
    if (kOnes) {
      // the ones column of V made the tensor core accumulate l = sum_j P_j (rescaled with O)
      uint32_t oo[16];
      tmem_ld16(o_tmem + (uint32_t)(p.l_col & ~15), oo);
      tmem_ld_wait();
      float lv = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.l_col & 15)) lv = __uint_as_float(oo[i]);
      l_run = lv;
    } else {
      named_bar_sync(2, 256);                        // xch is free again
      xch[0][ch][r] = l_run;
      named_bar_sync(1, 256);
      l_run += xch[0][ch ^ 1][r];
    }
    const float inv_l = 1.0f / l_run;
    const bool valid = (q0 + r) < p.sq;
    __nv_bfloat16* orow = p.out + ((long long)b * p.sq + q0 + r) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = ch * 16; c < p.d; c += SW * 4) {
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
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

static long long* g_attn_trace = nullptr;

static int attn_tc_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                          long long ldv, int dh_stride, void* out, long long ldo, int batch, int sq,
                          int skv, int heads, int d, float scale, int v_ones, cudaStream_t stream) {
  DL_CHECK_ARG(d % 8 == 0 && d >= 8, "attention: head dim %d must be a multiple of 8", d);
  const int d16 = (d + 15) / 16 * 16;
  DL_CHECK_ARG(dh_stride % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "attention(tc): strides must be multiples of 8 elements");
  DL_CHECK_ARG(dh_stride >= d16 || (dh_stride == d && d == d16),
               "attention(tc): d=%d needs a zero-padded per-head stride >= %d (got %d)", d, d16, dh_stride);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.sq = sq; p.skv = skv; p.d = d; p.dh_stride = dh_stride;
  p.ksteps = d16 / 16;
  p.nchunk_qk = (d16 + 63) / 64;
  if (v_ones) {
    DL_CHECK_ARG(dh_stride >= (d + 1 + 15) / 16 * 16,
                 "attention(tc): v_ones needs per-head stride >= ceil16(d+1)");
    p.dv = (d + 1 + 15) / 16 * 16;
    p.l_col = d;
  } else {
    p.dv = d16;
    p.l_col = -1;
  }
  DL_CHECK_ARG(p.dv <= 256, "attention(tc): head dim %d > 240 unsupported (use the GEMM path)", d);
  p.nchunk_v = (p.dv + 63) / 64;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = g_attn_trace;
  const int q_bytes = p.nchunk_qk * AT_CHUNK_BYTES;
  const int overhead = 1024 + 256 + 2048;          // alignment slack, barriers, static xch
  const int half_budget = (227 * 1024) / 2 - 1024;      // two CTAs per SM
  const int full_budget = 227 * 1024 - 3072;
  static int force_mode = -1;
  if (force_mode < 0) { const char* e = getenv("DL_ATTN_MODE"); force_mode = e ? atoi(e) : 0; }
  if (force_mode == 0 && g_attn_trace == nullptr) {
    const int rx = attn_x_launch(q, ldq, k, ldk, v, ldv, dh_stride, out, ldo, batch, sq, skv, heads, d, scale, v_ones,
                                 stream);
    if (rx >= 0) return rx;
    const int rc = attn_pp_launch(q, ldq, k, ldk, v, ldv, dh_stride, out, ldo, batch, sq, skv, heads, d, scale,
                                  v_ones, stream);
    if (rc >= 0) return rc;
  }
  // Configuration search:
  //  1. (DL_ATTN_MODE=3 only; measured 13 % slower on B200) two CTAs per SM with a double-buffered
  //     S: the key tile shrinks to 96 or 64 so that 2*kt + dv <= 256;
  //  2. two CTAs per SM (256 TMEM columns each), 128-key tile, single S buffer: one CTA's softmax
  //     runs under the other's MMAs;
  //  3. one CTA per SM (512 columns), double-buffered S, largest key tile with >= 2 smem stages;
  //  4. one CTA per SM, single S buffer.
  static const int kts[3] = {128, 96, 64};
  int kv_bytes = 0;
  bool done = false;
  for (int ki = 0; ki < 3 && !done && force_mode == 3; ++ki) {   // measured slower: opt-in only
    const int kt = kts[ki];
    kv_bytes = (p.nchunk_qk + p.nchunk_v) * kt * 128;
    if (2 * kt + p.dv <= 256 && q_bytes + 2 * kv_bytes + overhead <= half_budget) {
      p.kt = kt; p.sbuf = 2; p.tmem_cols = 256;
      p.stages = (half_budget - overhead - q_bytes) / kv_bytes;
      done = true;
    }
  }
  if (!done && force_mode != 1 && p.dv <= 128) {
    kv_bytes = (p.nchunk_qk + p.nchunk_v) * 128 * 128;
    if (q_bytes + kv_bytes + overhead <= half_budget) {
      p.kt = 128; p.sbuf = 1; p.tmem_cols = 256;
      p.stages = (half_budget - overhead - q_bytes) / kv_bytes;
      // S (128) + separate P (64) + O (dv) within the CTA's 256 columns, and K_{j+1} needs its
      // own smem stage next to V_j: SD1.5 head dim 40 (dv = 48)
      // (DL_ATTN_MODE=5 only: measured 3 % slower than the fused PV_j + S_{j+1} burst on B200 —
      // with both CTAs free-running their exp passes collide on the MUFU pipe)
      if (128 + 64 + p.dv <= 256 && p.stages >= 2 && force_mode == 5) p.pcol = 128;
      done = true;
    }
  }
  for (int ki = 0; ki < 3 && !done; ++ki) {
    const int kt = kts[ki];
    kv_bytes = (p.nchunk_qk + p.nchunk_v) * kt * 128;
    if (2 * kt + p.dv <= 512 && q_bytes + 2 * kv_bytes + overhead <= full_budget) {
      p.kt = kt; p.sbuf = 2; p.tmem_cols = 512;
      p.stages = (full_budget - overhead - q_bytes) / kv_bytes;
      done = true;
    }
  }
  if (!done) {
    p.kt = 64; p.sbuf = 1; p.tmem_cols = 512;
    kv_bytes = (p.nchunk_qk + p.nchunk_v) * p.kt * 128;
    p.stages = (full_budget - overhead - q_bytes) / kv_bytes;
  }
  if (p.stages > 4) p.stages = 4;
  DL_CHECK_ARG(p.stages >= 1 && p.kt + p.dv <= 512, "attention(tc): head dim %d does not fit shared memory", d);
  const uint32_t box[2] = {64, AT_TILE};
  const uint32_t kbox[2] = {64, (uint32_t)p.kt};
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * sq};
    const uint64_t str[1] = {(uint64_t)ldq * 2};
    if (make_tmap_bf16(&p.tmQ, q, 2, dims, str, box)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * skv};
    const uint64_t str[1] = {(uint64_t)ldk * 2};
    if (make_tmap_bf16(&p.tmK, k, 2, dims, str, kbox)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * skv};
    const uint64_t str[1] = {(uint64_t)ldv * 2};
    if (make_tmap_bf16(&p.tmV, v, 2, dims, str, kbox)) return 1;
  }
  const int smem_bytes = q_bytes + p.stages * kv_bytes + overhead;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    const int lim = 227 * 1024 - 3072;                          // minus the static xch buffer
    cudaError_t e = cudaSuccess;
#define DL_ATTN_ATTR(...)                                                                       \
  if (e == cudaSuccess)                                                                         \
    e = cudaFuncSetAttribute(attn_tc_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim)
    DL_ATTN_ATTR(true, true, 0, 0);
    DL_ATTN_ATTR(true, true, 3, 128);
    DL_ATTN_ATTR(true, true, 4, 128);
    DL_ATTN_ATTR(true, true, 5, 128);
    DL_ATTN_ATTR(false, false, 0, 0);
    DL_ATTN_ATTR(true, true, 3, 128, 2);
    DL_ATTN_ATTR(true, true, 3, 128, 3);
    DL_ATTN_ATTR(true, true, 4, 128, 2);
    DL_ATTN_ATTR(true, true, 5, 128, 2);
#undef DL_ATTN_ATTR
    if (e != cudaSuccess) { set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set[dev & 63] = true;
  }
  dim3 grid((sq + AT_TILE - 1) / AT_TILE, heads, batch);
  static int generic = -1;
  if (generic < 0) { const char* e = getenv("DL_ATTN_GENERIC"); generic = e ? atoi(e) : 0; }
  // exponentials partly on the FMA pipe (see kPoly): 0 = all MUFU, 2 / 3 = 25 % / 37.5 % polynomial.
  // Measured on B200 (B = 16, S = 4096, d = 40): 981 us all-MUFU, 942 us at 25 %, 975 us at 37.5 %;
  // d = 80 (S = 1024) unchanged.  Default: 25 % for head dim 40 only; DL_ATTN_POLY=0 switches it
  // off, =2 / =3 apply it to head dims 64 / 80 as well.
  static int poly = -2;
  if (poly == -2) { const char* e = getenv("DL_ATTN_POLY"); poly = e ? atoi(e) : -1; }
  const bool long_keys = skv >= 256;           // short key lists (cross attention) are not exp-bound
  if (p.l_col < 0) {
    attn_tc_kernel<false, false, 0, 0><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 3 && !generic && poly == 3 && long_keys) {
    attn_tc_kernel<true, true, 3, 128, 3><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 3 && !generic && (poly >= 2 || poly == -1) && long_keys) {
    attn_tc_kernel<true, true, 3, 128, 2><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 4 && !generic && poly >= 2 && long_keys) {
    attn_tc_kernel<true, true, 4, 128, 2><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 5 && !generic && poly >= 2 && long_keys) {
    attn_tc_kernel<true, true, 5, 128, 2><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 3 && !generic) {       // SD1.5 head dim 40
    attn_tc_kernel<true, true, 3, 128><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 4 && !generic) {       // SDXL head dim 64
    attn_tc_kernel<true, true, 4, 128><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else if (p.kt == 128 && p.ksteps == 5 && !generic) {       // SD1.5 head dim 80
    attn_tc_kernel<true, true, 5, 128><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  } else {
    attn_tc_kernel<true, true, 0, 0><<<grid, AT_THREADS, smem_bytes, stream>>>(p);
  }
  return check_launch("attention(tc)");
}

}  // namespace dl

extern "C" int dl_debug_attention_trace(void* device_buf_i64_256) {
  dl::g_attn_trace = reinterpret_cast<long long*>(device_buf_i64_256);
  return 0;
}

extern "C" int dl_attention(const void* q, long long ldq, const void* k, long long ldk,
                            const void* v, long long ldv, int dh_stride, void* out, long long ldo,
                            int batch, int sq, int skv, int heads, int d, float scale, int impl,
                            int v_ones, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DL_CHECK_ARG(q && k && v && out, "attention: null pointer");
  DL_CHECK_ARG(batch > 0 && sq > 0 && skv > 0 && heads > 0 && d > 0, "attention: bad dims");
  if (impl == DL_ATTN_SIMT || impl == DL_ATTN_SIMT_CAUSAL)
    return attn_simt_launch(q, ldq, k, ldk, v, ldv, dh_stride, out, ldo, batch, sq, skv, heads, d,
                            scale, impl == DL_ATTN_SIMT_CAUSAL ? 1 : 0, stream);
  DL_CHECK_ARG(impl == DL_ATTN_TC, "attention: unknown impl %d", impl);
  return attn_tc_launch(q, ldq, k, ldk, v, ldv, dh_stride, out, ldo, batch, sq, skv, heads, d,
                        scale, v_ones, stream);
}
