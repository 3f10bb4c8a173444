// Wide-head flash attention for sm_100a (tcgen05 + TMEM + TMA): the AutoencoderKL mid-block attention
// (one head, head dim 512, S = 4096 / 9216 / 16384 tokens), SURVEY.md K5 "split-d for d = 512".  Replaces
// diffusers' Attention(AttnProcessor2_0) in UNetMidBlock2D of the VAE decoder (reference `backends/rknnlcm.py:618`
// -> `vae.decode`).
//
// Until round 2 this attention ran unfused through the GEMM kernel, per image: QK^T -> fp32 scores [S, S] in HBM ->
// softmax -> bf16 probabilities -> PV: 201 MB of traffic per 512^2 image against 16.8 MB algorithmic, growing with S^2.
// A 128 x 512 fp32 output tile is the whole TMEM (512 columns) and leaves no room for S, so the head dim is split:
// a CTA owns 128 queries and ONE HALF of the output columns (256), and computes the full scores itself
// (QK^T is done twice per query tile: +34 GFLOP per image, nothing next to the S^2 traffic it removes).
//   shared memory  Q 128 x 512 (128 KB, once) | K_j 64 x 512 (64 KB, a ring of eight 64-column chunks) |
//                  V_j 64 x 256 (32 KB): 224 KB, nothing left for a second stage
//   TMEM           S / P, two buffers [0, 64) [64, 128)   O [128, 384)
//   warp 0 TMA, warp 1 MMA (S_j: 32 K-steps of M128 N64 K16, SS; O += P_j V_j: 4 K-steps of M128 N256 K16, TS),
//   warps 2-5 softmax, one thread per query row: online softmax with a lazy running max, row sums in registers.
// The tensor pipe runs S_0 S_1 PV_0 S_2 PV_1 ...: S_{j+1} is computed while the softmax warps work on S_j.  Each
// 64-column chunk of K has its own full / empty barrier pair, so chunk c of K_{j+1} is fetched as soon as the four
// K-steps of S_j that read chunk c retire — the K stream runs one tile ahead without a second K buffer.
#include <stdlib.h>

#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int AW_THREADS = 192;
constexpr int AW_TILE = 128;      // queries per CTA
constexpr int AW_KT = 64;         // keys per tile

struct AttnWParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* out;
  long long ldo;
  int sq, skv, d;        // d: head dim (multiple of 64, <= 512)
  int dhalf;             // output columns per CTA = d / 2 (multiple of 64, <= 256)
  float scale_log2;
};

__global__ void __launch_bounds__(AW_THREADS, 1)
attn_wide_kernel(const __grid_constant__ AttnWParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nck = p.d / 64;                 // 64-column chunks of Q / K
  const int ncv = p.dhalf / 64;             // ... of this CTA's V half
  constexpr int QCH = AW_TILE * 128;        // one Q chunk: 128 rows x 128 B
  constexpr int KCH = AW_KT * 128;          // one K / V chunk: 64 rows x 128 B
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + nck * QCH;
  uint8_t* sV = sK + nck * KCH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ncv * KCH);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;       // [8] one per 64-column chunk
  uint64_t* k_empty = bars + 9;      // [8]
  uint64_t* v_full = bars + 17;
  uint64_t* v_empty = bars + 18;
  uint64_t* s_full = bars + 19;      // [2] one per S buffer
  uint64_t* p_ready = bars + 21;     // [2]
  uint64_t* pv_done = bars + 23;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AW_TILE;
  const int half = blockIdx.y, b = blockIdx.z;
  const int n_tiles = p.skv / AW_KT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int c = 0; c < 8; ++c) { mbar_init(k_full + c, 1); mbar_init(k_empty + c, 1); }
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(p_ready + i, 4); }
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512u); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t O_COL = 128u;

  if (warp == 0) {
    // ============================ TMA producer ============================
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(q_full, (uint32_t)(nck * QCH));
      for (int c = 0; c < nck; ++c) tma_load_2d(sQ + c * QCH, &p.tmQ, q_full, c * 64, b * p.sq + q0);
    }
    __syncwarp();
    // in the order the tensor pipe consumes them: K_0 K_1 V_0 K_2 V_1 ...
    auto load_k = [&](int j) {
      for (int c = 0; c < nck; ++c) {
        mbar_wait(k_empty + c, (uint32_t)((j & 1) ^ 1));
        if (issuer) {
          mbar_expect_tx(k_full + c, (uint32_t)KCH);
          tma_load_2d(sK + c * KCH, &p.tmK, k_full + c, c * 64, b * p.skv + j * AW_KT);
        }
        __syncwarp();
      }
    };
    load_k(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) load_k(j + 1);
      mbar_wait(v_empty, (uint32_t)((j & 1) ^ 1));
      if (issuer) {
        mbar_expect_tx(v_full, (uint32_t)(ncv * KCH));
        for (int c = 0; c < ncv; ++c)
          tma_load_2d(sV + c * KCH, &p.tmV, v_full, half * p.dhalf + c * 64, b * p.skv + j * AW_KT);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const bool issuer = elect_one();
    const uint32_t idesc_s = umma_idesc_bf16(128, AW_KT, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16(128, (uint32_t)p.dhalf, 0, 1);
    const uint32_t hi_k = umma_desc_hi_sw128(1024);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ));
    const uint32_t k_lo = umma_desc_lo(smem_u32(sK));
    const uint32_t v_lo = umma_desc_lo(smem_u32(sV), (uint32_t)KCH);     // LBO: stride between 64-wide d chunks
    // S_j goes to buffer j & 1.  That buffer last held S_{j-2} / P_{j-2}: the softmax warps are done with it
    // (p_ready of j-2 was waited on before PV_{j-2}) and PV_{j-2} is ahead of S_j in the in-order pipe.
    auto issue_s = [&](int j) {
      const uint32_t s_col = tmem_base + (uint32_t)((j & 1) * 64);
      for (int c = 0; c < nck; ++c) {
        mbar_wait(k_full + c, (uint32_t)(j & 1));
        tc_fence_after();
        if (issuer) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss_lohi(s_col, q_lo + (uint32_t)(c * (QCH >> 4) + ks * 2), k_lo + (uint32_t)(c * (KCH >> 4) + ks * 2),
                         hi_k, idesc_s, (c > 0 || ks > 0) ? 1u : 0u);
          umma_commit(k_empty + c);
        }
        __syncwarp();
      }
      if (issuer) umma_commit(s_full + (j & 1));
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_s(j + 1);
      mbar_wait(v_full, (uint32_t)(j & 1));
      mbar_wait(p_ready + (j & 1), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      if (issuer) {
        const uint32_t p_col = tmem_base + (uint32_t)((j & 1) * 64);
#pragma unroll
        for (int ks = 0; ks < AW_KT / 16; ++ks)     // 16 keys = two 8-row atoms = 2048 B; P: 8 packed columns
          umma_ts_lohi(tmem_base + O_COL, p_col + (uint32_t)(ks * 8), v_lo + (uint32_t)(ks * 128), hi_k, idesc_o,
                       (ks > 0 || j > 0) ? 1u : 0u);
        umma_commit(v_empty);
        umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ============================ softmax + epilogue ============================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + O_COL;
    const float sc = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full + (j & 1), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t s_tmem = tmem_base + lane_off + (uint32_t)((j & 1) * 64);
      uint32_t s[AW_KT];
      tmem_ld32(s_tmem, reinterpret_cast<uint32_t(&)[32]>(s[0]));
      tmem_ld32(s_tmem + 32u, reinterpret_cast<uint32_t(&)[32]>(s[32]));
      tmem_ld_wait();
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < AW_KT; i += 2) {
        m0 = fmaxf(m0, __uint_as_float(s[i]));
        m1 = fmaxf(m1, __uint_as_float(s[i + 1]));
      }
      // lazy running max (see attention_pp.cu): moved only when the tile max exceeds it by more than 2^8
      const float m_tile = fmaxf(m0, m1) * sc;
      const float m_new = (m_tile > m_run + 8.0f) ? m_tile : m_run;
      const float corr = fast_exp2(m_run - m_new);
      m_run = m_new;
      float ls0 = 0.f, ls1 = 0.f;
      uint32_t pk[AW_KT / 2];
#pragma unroll
      for (int i = 0; i < AW_KT / 2; ++i) {
        const __nv_bfloat162 pb = __floats2bfloat162_rn(fast_exp2(fmaf(__uint_as_float(s[2 * i]), sc, -m_new)),
                                                        fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), sc, -m_new)));
        const float2 pf = __bfloat1622float2(pb);          // sum what the tensor core will multiply
        ls0 += pf.x;
        ls1 += pf.y;
        pk[i] = *reinterpret_cast<const uint32_t*>(&pb);
      }
      l_run = l_run * corr + (ls0 + ls1);
      // P_j over S_j's own columns (PV_{j-1} reads the other buffer)
      tmem_st16(s_tmem, reinterpret_cast<const uint32_t(&)[16]>(pk[0]));
      tmem_st16(s_tmem + 16u, reinterpret_cast<const uint32_t(&)[16]>(pk[16]));
      if (j > 0 && __any_sync(0xffffffffu, corr != 1.0f)) {
        // O is rescaled between PV_{j-1} (retired: pv_done) and PV_j (not issued before p_ready below).  pv_done is
        // at phase j-1 or j here whether or not earlier tiles waited on it, so the parity is unambiguous.
        mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < p.dhalf; c += 16) {
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
      if (lane == 0) mbar_arrive(p_ready + (j & 1));
    }
    mbar_wait(pv_done, (uint32_t)((n_tiles - 1) & 1));
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + r;
    const bool valid = qrow < p.sq;
    __nv_bfloat16* orow = p.out + ((long long)b * p.sq + qrow) * p.ldo + half * p.dhalf;
#pragma unroll 1
    for (int c = 0; c < p.dhalf; c += 16) {
      uint32_t oo[16];
      tmem_ld16(o_tmem + (uint32_t)c, oo);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
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
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

}  // namespace dl

// q, k: bf16 rows [batch * s, ld] (head dim d contiguous), v: bf16 [batch * skv, ldv]; out bf16 [batch * sq, ldo].
// One head.  d a multiple of 128 up to 512, sq a multiple of 128, skv a multiple of 64.
extern "C" int dl_attention_wide(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                 long long ldv, void* out, long long ldo, int batch, int sq, int skv, int d,
                                 float scale, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DL_CHECK_ARG(q && k && v && out && batch > 0 && sq > 0 && skv > 0, "attention_wide: bad args");
  DL_CHECK_ARG(d % 128 == 0 && d >= 128 && d <= 512, "attention_wide: head dim %d must be a multiple of 128 in [128, 512]", d);
  DL_CHECK_ARG(sq % AW_TILE == 0 && skv % AW_KT == 0, "attention_wide: sq %% 128 and skv %% 64 must be 0 (got %d, %d)", sq, skv);
  DL_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention_wide: strides must be multiples of 8");
  AttnWParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.sq = sq; p.skv = skv; p.d = d; p.dhalf = d / 2;
  p.scale_log2 = scale * 1.4426950408889634f;
  const uint32_t qbox[2] = {64, AW_TILE};
  const uint32_t kbox[2] = {64, AW_KT};
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)batch * sq};
    const uint64_t str[1] = {(uint64_t)ldq * 2};
    if (make_tmap_bf16(&p.tmQ, q, 2, dims, str, qbox)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)batch * skv};
    const uint64_t str[1] = {(uint64_t)ldk * 2};
    if (make_tmap_bf16(&p.tmK, k, 2, dims, str, kbox)) return 1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d, (uint64_t)batch * skv};
    const uint64_t str[1] = {(uint64_t)ldv * 2};
    if (make_tmap_bf16(&p.tmV, v, 2, dims, str, kbox)) return 1;
  }
  const int smem_bytes = (d / 64) * (AW_TILE * 128) + (d / 64) * (AW_KT * 128) + (d / 128) * (AW_KT * 128) + 1024 + 256;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attn_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("attention_wide: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
    attr_set[dev & 63] = true;
  }
  DL_CHECK_ARG(smem_bytes <= 227 * 1024, "attention_wide: %d bytes of shared memory needed", smem_bytes);
  dim3 grid(sq / AW_TILE, 2, batch);
  attn_wide_kernel<<<grid, AW_THREADS, smem_bytes, stream>>>(p);
  return check_launch("attention_wide");
}
