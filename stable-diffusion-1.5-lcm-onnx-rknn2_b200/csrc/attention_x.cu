// Short-key attention for sm_100a (tcgen05 + TMEM + TMA): the cross-attention of the UNet (attn2: the 77 text
// tokens are the keys, S_kv <= 128), SURVEY.md K6.  Replaces diffusers' AttnProcessor2_0 in
// BasicTransformerBlock.attn2, reached from reference `backends/cuda_worker.py:222`.
//
// With one key tile there is no online softmax and almost no arithmetic per CTA: in `attn_tc_kernel`
// (one CTA per 128 queries) the fixed cost of a CTA — launch, TMEM allocation, barrier setup, the first
// TMA round trip for K / V — was ~10x the work, and the kernel sat at 5-7 % of the tensor roofline (90 us for
// B = 16, S = 4096, 8 heads).  Here a CTA owns one (image, head) and a RANGE of query tiles:
//   warp 0   TMA: K, V once; the Q tiles of its range through a 2-stage ring
//   warp 1   MMA: S_i = Q_i K^T (SS), O_i = P_i V (TS, P from TMEM over the S columns); S_{i+1} is issued
//            right behind PV_i, so it runs under the epilogue of tile i
//   warps 2-5  one thread per query row: S row -> registers, exact row max, exp2, P (packed bf16) back to
//            TMEM, then O_i / l -> global (l = V's ones column, as in attention.cu)
// TMEM: S / P [0, 128), O [128, 128 + dv): 256 columns when dv <= 128 (two CTAs per SM), else 512.
#include <stdlib.h>

#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int AX_THREADS = 192;
constexpr int AX_TILE = 128;

struct AttnXParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  long long ldo;
  int sq, skv, d, dh_stride;
  int ksteps;        // QK^T K-steps of 16
  int nchunk_qk;     // 64-column TMA boxes per Q / K tile
  int dv, nchunk_v;  // PV MMA N (incl. the ones column), 64-column boxes per V tile
  int l_col;
  int kt;            // key tile = ceil16(skv) <= 128
  int tiles_per_cta; // query tiles one CTA walks
  int tmem_cols;
  float scale_log2;
};

__device__ __forceinline__ float ax_fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// KT: key tile (S columns, PV depth, K / V box rows): 80 for the 77 text tokens, 128 otherwise.
template <int KT>
__global__ void __launch_bounds__(AX_THREADS, 2)
attn_x_kernel(const __grid_constant__ AttnXParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_bytes = p.nchunk_qk * AX_TILE * 128;
  constexpr int kchunk = KT * 128;
  const int k_bytes = p.nchunk_qk * kchunk;
  const int v_bytes = p.nchunk_v * kchunk;
  uint8_t* sQ = smem;                                   // 2 stages
  uint8_t* sK = smem + 2 * q_bytes;
  uint8_t* sV = sK + ((k_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ((v_bytes + 1023) & ~1023));
  uint64_t* kv_full = bars;          // 1
  uint64_t* q_full = bars + 1;       // [2]
  uint64_t* q_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;       // 1
  uint64_t* p_ready = bars + 6;      // 1
  uint64_t* pv_done = bars + 7;      // 1
  uint64_t* o_free = bars + 8;       // 1: the epilogue has read O_i
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_qt = (p.sq + AX_TILE - 1) / AX_TILE;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int t1 = min(n_qt, t0 + p.tiles_per_cta);
  const int n_my = t1 - t0;                              // > 0 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(pv_done, 1);
    mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t O_COL = 128u;

  if (warp == 0) {
    // ============================ TMA producer ============================
    const bool issuer = elect_one();
    const int col0 = h * p.dh_stride;
    if (issuer) {
      mbar_expect_tx(kv_full, (uint32_t)(k_bytes + v_bytes));
      for (int c = 0; c < p.nchunk_qk; ++c)
        tma_load_2d(sK + c * kchunk, &p.tmK, kv_full, col0 + c * 64, b * p.skv);
      for (int c = 0; c < p.nchunk_v; ++c)
        tma_load_2d(sV + c * kchunk, &p.tmV, kv_full, col0 + c * 64, b * p.skv);
    }
    __syncwarp();
    for (int i = 0; i < n_my; ++i) {
      const int st = i & 1;
      mbar_wait(&q_empty[st], (uint32_t)(((i >> 1) & 1) ^ 1));
      if (issuer) {
        mbar_expect_tx(&q_full[st], (uint32_t)q_bytes);
        for (int c = 0; c < p.nchunk_qk; ++c)
          tma_load_2d(sQ + st * q_bytes + c * (AX_TILE * 128), &p.tmQ, &q_full[st], col0 + c * 64,
                      b * p.sq + (t0 + i) * AX_TILE);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const bool issuer = elect_one();
    const uint32_t idesc_s = umma_idesc_bf16(128, (uint32_t)KT, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16(128, (uint32_t)p.dv, 0, 1);
    const uint32_t hi_k = umma_desc_hi_sw128(1024);
    const uint32_t k_lo = umma_desc_lo(smem_u32(sK));
    const uint32_t v_lo = umma_desc_lo(smem_u32(sV), (uint32_t)kchunk);
    constexpr int kchunk16 = kchunk >> 4;
    auto issue_s = [&](int i) {
      const int st = i & 1;
      mbar_wait(&q_full[st], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t q_lo = umma_desc_lo(smem_u32(sQ + st * q_bytes));
      if (issuer) {
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint32_t qoff = (uint32_t)((ks >> 2) * ((AX_TILE * 128) >> 4) + (ks & 3) * 2);
          const uint32_t koff = (uint32_t)((ks >> 2) * kchunk16 + (ks & 3) * 2);
          umma_ss_lohi(tmem_base, q_lo + qoff, k_lo + koff, hi_k, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(&q_empty[st]);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    issue_s(0);
    for (int i = 0; i < n_my; ++i) {
      mbar_wait(p_ready, (uint32_t)(i & 1));
      if (i > 0) mbar_wait(o_free, (uint32_t)((i - 1) & 1));     // O_{i-1} has been read out
      tc_fence_after();
      if (issuer) {
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks)
          umma_ts_lohi(tmem_base + O_COL, tmem_base + (uint32_t)(ks * 8), v_lo + (uint32_t)(ks * 128), hi_k, idesc_o,
                       ks > 0 ? 1u : 0u);
        umma_commit(pv_done);
      }
      __syncwarp();
      // S_{i+1} right behind PV_i (the in-order tensor pipe keeps it behind PV_i's reads of P_i): it runs
      // under the epilogue of tile i
      if (i + 1 < n_my) issue_s(i + 1);
    }
  } else {
    // ============================ softmax + epilogue ============================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off;
    const uint32_t o_tmem = tmem_base + lane_off + O_COL;
    const float sc = p.scale_log2;
    for (int i = 0; i < n_my; ++i) {
      mbar_wait(s_full, (uint32_t)(i & 1));
      tc_fence_after();
      uint32_t s[KT];
#pragma unroll
      for (int c = 0; c < KT / 16; ++c)
        tmem_ld16(s_tmem + (uint32_t)(c * 16), reinterpret_cast<uint32_t(&)[16]>(s[c * 16]));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < KT; ++j)
        if (j >= p.skv) s[j] = 0xff800000u;               // padded / foreign key columns: -inf
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < KT; j += 4) {
        m0 = ax_fmax3(m0, __uint_as_float(s[j]), __uint_as_float(s[j + 1]));
        m1 = ax_fmax3(m1, __uint_as_float(s[j + 2]), __uint_as_float(s[j + 3]));
      }
      const float nm = -fmaxf(m0, m1) * sc;
#pragma unroll
      for (int c = 0; c < KT / 16; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack_bf16x2(fast_exp2(fmaf(__uint_as_float(s[c * 16 + 2 * j]), sc, nm)),
                              fast_exp2(fmaf(__uint_as_float(s[c * 16 + 2 * j + 1]), sc, nm)));
        tmem_st8(s_tmem + (uint32_t)(c * 8), pk);          // P over the S columns already in registers
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // epilogue of this tile: O / l
      mbar_wait(pv_done, (uint32_t)(i & 1));
      tc_fence_after();
      float lv = 0.f;
      {
        uint32_t oo[16];
        tmem_ld16(o_tmem + (uint32_t)(p.l_col & ~15), oo);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j == (p.l_col & 15)) lv = __uint_as_float(oo[j]);
      }
      const float inv_l = 1.0f / lv;
      const int qrow = (t0 + i) * AX_TILE + r;
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// -> 0 launched, 1 error, -1 shape not covered (caller falls through to attn_tc_kernel)
int attn_x_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                  int dh_stride, void* out, long long ldo, int batch, int sq, int skv, int heads, int d,
                  float scale, int v_ones, cudaStream_t stream) {
  static int mode = -2;
  if (mode == -2) { const char* e = getenv("DL_ATTN_X"); mode = e ? atoi(e) : 1; }
  const int d16 = (d + 15) / 16 * 16;
  const int dv = (d + 1 + 15) / 16 * 16;
  if (!mode || !v_ones || d % 8 || skv > 128 || dv > 256 || dh_stride < dv || sq < 256) return -1;
  if (dh_stride % 8 || ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return -1;
  AttnXParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.sq = sq; p.skv = skv; p.d = d; p.dh_stride = dh_stride;
  p.ksteps = d16 / 16;
  p.nchunk_qk = (d16 + 63) / 64;
  p.dv = dv;
  p.nchunk_v = (dv + 63) / 64;
  p.l_col = d;
  p.kt = skv <= 80 ? 80 : 128;
  p.tmem_cols = 128 + dv <= 256 ? 256 : 512;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int q_bytes = p.nchunk_qk * AX_TILE * 128;
  const int kchunk = p.kt * 128;
  const int smem_bytes = 2 * q_bytes + ((p.nchunk_qk * kchunk + 1023) & ~1023) + ((p.nchunk_v * kchunk + 1023) & ~1023) +
                         1024 + 256;
  if (smem_bytes > 200 * 1024) return -1;
  // query tiles per CTA: about two resident CTAs per SM in one wave
  const int n_qt = (sq + AX_TILE - 1) / AX_TILE;
  const int sms = num_sms();
  int chunks = (2 * sms + heads * batch - 1) / (heads * batch);
  if (chunks < 1) chunks = 1;
  if (chunks > n_qt) chunks = n_qt;
  p.tiles_per_cta = (n_qt + chunks - 1) / chunks;
  chunks = (n_qt + p.tiles_per_cta - 1) / p.tiles_per_cta;
  const uint32_t qbox[2] = {64, AX_TILE};
  const uint32_t kbox[2] = {64, (uint32_t)p.kt};
  {
    const uint64_t dims[2] = {(uint64_t)heads * dh_stride, (uint64_t)batch * sq};
    const uint64_t str[1] = {(uint64_t)ldq * 2};
    if (make_tmap_bf16(&p.tmQ, q, 2, dims, str, qbox)) return 1;
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
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attn_x_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_x_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("attention(x): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set[dev & 63] = true;
  }
  dim3 grid(chunks, heads, batch);
  if (p.kt == 80) attn_x_kernel<80><<<grid, AX_THREADS, smem_bytes, stream>>>(p);
  else attn_x_kernel<128><<<grid, AX_THREADS, smem_bytes, stream>>>(p);
  return check_launch("attention(x)");
}

}  // namespace dl
