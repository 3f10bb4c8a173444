// GroupNorm(+SiLU) and LayerNorm for NHWC bf16 activations (HBM-bound kernels).
//
// GroupNorm is ONE persistent cooperative launch (grid = #SMs): the batch is cut into waves of
// images small enough to stay in the 126 MB L2; per wave every CTA owns a contiguous slab of
// pixels of one image and
//   A. streams it through a 6-deep ring of 16 KB shared-memory buffers filled by 1-D bulk async
//      copies (cp.async.bulk + mbarrier: ~96 KB in flight per SM, no registers tied up) and
//      reduces it to per-group (mean, M2) partials (fp32, fixed order),
//   -- grid barrier --
//   B. Chan-merges the partials of its image (warp shuffles, fixed order => deterministic, no
//      float atomics), builds per-channel scale/shift in registers, streams the slab again (an
//      L2 hit: it was read microseconds ago) -> normalise (+SiLU) -> 16-byte coalesced stores.
// HBM traffic is therefore the algorithmic read-once + write-once (4 B/element).
// Both phases accept two sources and emit the channel concat [x0 | x1] (UNet skip concat folded
// into the norm: the concat tensor is never written in un-normalised form).
#include <stdlib.h>

#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int GN_THREADS = 512;
constexpr int GN_MAX_C = 2560;
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_STAGES = 6;                        // 6 x 32 KB in flight per SM
constexpr int GN_CHUNK_BYTES = 32 * 1024;
constexpr long long GN_WAVE_BYTES = 48ll << 20;     // input bytes per wave of 148 CTAs (L2-resident)

__device__ __forceinline__ uint4 ld_vec8(const __nv_bfloat16* x0, int c0,
                                         const __nv_bfloat16* x1, int c1, long long pix, int v) {
  // vector v covers channels [8v, 8v+8) of the virtual concat
  const int ch = v * 8;
  const __nv_bfloat16* p = (ch < c0) ? (x0 + pix * c0 + ch) : (x1 + pix * c1 + (ch - c0));
  return *reinterpret_cast<const uint4*>(p);
}

struct GnParams {
  const __nv_bfloat16* x0; int c0;
  const __nv_bfloat16* x1; int c1;
  int nimg, hw, groups, V, L;
  int P;                     // pixels per streamed chunk (multiple of L)
  float eps;
  const float* gamma; const float* beta;
  int apply_silu;
  __nv_bfloat16* out;
  float* partial;            // [2][gridDim][groups][2]
  unsigned int* counter;     // grid barrier (zeroed before launch)
  int wave_imgs, slabs;      // images per wave, slabs (CTAs) per image
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// Consumer side of the ring: waits for each chunk of the slab [p_begin, p_end) of image `img`
// and calls f(vec, pixel) for this thread's (channel-vector v, pixel-lane l) elements.  The
// chunks were requested by the producer warp (gn_produce) in exactly the same order.
template <class F>
__device__ __forceinline__ void gn_consume(const GnParams& p, const uint8_t* ring, uint64_t* full,
                                           uint64_t* empty, uint32_t& chunk_base, int img,
                                           int p_begin, int p_end, int v, int l, bool lane_ok, F&& f) {
  const int npix = p_end - p_begin;
  const int n_chunks = (npix + p.P - 1) / p.P;
  const long long pix0 = (long long)img * p.hw + p_begin;
  const size_t part0 = (size_t)p.P * p.c0 * 2;          // bytes of source 0 in a full chunk buffer
  const bool from0 = v * 8 < p.c0;
  const size_t my_off = from0 ? ((size_t)l * p.c0 + v * 8) * 2
                              : part0 + ((size_t)l * p.c1 + (v * 8 - p.c0)) * 2;
  const size_t my_step = (size_t)p.L * (from0 ? p.c0 : p.c1) * 2;
  for (int i = 0; i < n_chunks; ++i) {
    const uint32_t g = chunk_base + (uint32_t)i;
    const int stage = (int)(g % GN_STAGES);
    mbar_wait(&full[stage], (g / GN_STAGES) & 1);
    const uint8_t* src = ring + (size_t)stage * GN_CHUNK_BYTES + my_off;
    const int cp = min(p.P, npix - i * p.P);
    if (lane_ok) {
      const long long pixc = pix0 + (long long)i * p.P;
      for (int px = l; px < cp; px += p.L, src += my_step)
        f(*reinterpret_cast<const uint4*>(src), pixc + px);
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage]);     // this warp is done with the buffer
  }
  chunk_base += (uint32_t)n_chunks;
}

// Producer (one elected thread): requests the chunks of a slab; runs ahead of the consumers by
// up to GN_STAGES buffers, across phase and wave boundaries (inputs are read-only).
__device__ __forceinline__ void gn_produce(const GnParams& p, uint8_t* ring, uint64_t* full,
                                           uint64_t* empty, uint32_t& chunk_base, int img,
                                           int p_begin, int p_end) {
  const int npix = p_end - p_begin;
  const int n_chunks = (npix + p.P - 1) / p.P;
  const long long pix0 = (long long)img * p.hw + p_begin;
  const size_t part0 = (size_t)p.P * p.c0 * 2;
  for (int i = 0; i < n_chunks; ++i) {
    const uint32_t g = chunk_base + (uint32_t)i;
    const int stage = (int)(g % GN_STAGES);
    mbar_wait(&empty[stage], ((g / GN_STAGES) & 1) ^ 1);
    const int cp = min(p.P, npix - i * p.P);
    uint8_t* buf = ring + (size_t)stage * GN_CHUNK_BYTES;
    const uint32_t b0 = (uint32_t)cp * p.c0 * 2, b1 = (uint32_t)cp * p.c1 * 2;
    mbar_expect_tx(&full[stage], b0 + b1);
    bulk_load_1d(buf, p.x0 + (pix0 + (long long)i * p.P) * p.c0, b0, &full[stage]);
    if (p.c1 > 0)
      bulk_load_1d(buf + part0, p.x1 + (pix0 + (long long)i * p.P) * p.c1, b1, &full[stage]);
  }
  chunk_base += (uint32_t)n_chunks;
}

__global__ void __launch_bounds__(GN_THREADS + 32, 1) gn_fused_kernel(const GnParams p) {
  extern __shared__ __align__(128) uint8_t gn_smem[];
  uint8_t* ring = gn_smem;                                                   // GN_STAGES x chunk
  float* s_sum = reinterpret_cast<float*>(gn_smem + GN_STAGES * GN_CHUNK_BYTES);   // [512*8]
  float* s_sq = s_sum + GN_THREADS * 8;
  __shared__ float s_mean[GN_MAX_GROUPS];
  __shared__ float s_rstd[GN_MAX_GROUPS];
  __shared__ __align__(8) uint64_t full[GN_STAGES];
  __shared__ __align__(8) uint64_t empty[GN_STAGES];
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GN_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GN_THREADS / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t chunk_base = 0;
  const int num_waves = (p.nimg + p.wave_imgs - 1) / p.wave_imgs;
  const int pps = (p.hw + p.slabs - 1) / p.slabs;     // pixels per slab

  if (warp == GN_THREADS / 32) {
    // ============================ producer warp ============================
    if (lane == 0) {
      for (int w = 0; w < num_waves; ++w) {
        const int item = blockIdx.x;
        const int img = w * p.wave_imgs + item / p.slabs;
        const int p_begin = (item % p.slabs) * pps;
        const int p_end = min(p.hw, p_begin + pps);
        const bool active = (item < p.wave_imgs * p.slabs) && (img < p.nimg) && (p_end > p_begin);
        if (!active) continue;
        gn_produce(p, ring, full, empty, chunk_base, img, p_begin, p_end);   // phase A stream
        gn_produce(p, ring, full, empty, chunk_base, img, p_begin, p_end);   // phase B stream
      }
    }
    return;
  }

  // ============================ 16 consumer warps ============================
  const int v = threadIdx.x % p.V;
  const int l = threadIdx.x / p.V;
  const bool lane_ok = l < p.L;                       // threads beyond V*L idle in the loops
  for (int w = 0; w < num_waves; ++w) {
    const int item = blockIdx.x;
    const int img = w * p.wave_imgs + item / p.slabs;
    const int slab = item % p.slabs;
    const int p_begin = slab * pps;
    const int p_end = min(p.hw, p_begin + pps);
    const bool active = (item < p.wave_imgs * p.slabs) && (img < p.nimg) && (p_end > p_begin);
    float* part = p.partial + ((size_t)(w & 1) * gridDim.x) * p.groups * 2;
    // ---------------- phase A: partial statistics of my slab ----------------
    if (active) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
      gn_consume(p, ring, full, empty, chunk_base, img, p_begin, p_end, v, l, lane_ok,
                 [&](const uint4& u, long long) {
                   const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                   for (int j = 0; j < 4; ++j) {
                     const float2 f = unpack_bf16x2(ww[j]);
                     s[2 * j] += f.x; q[2 * j] += f.x * f.x;
                     s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
                   }
                 });
      if (lane_ok) {
        // layout [l][C]: channel-contiguous so the per-group gather below is a linear walk
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s_sum[l * C + v * 8 + j] = s[j];
          s_sq[l * C + v * 8 + j] = q[j];
        }
      }
    }
    named_bar_sync(1, GN_THREADS);
    if (active) {
      // one warp per group (round-robin): lanes stride over the L x cpg partial sums, then a
      // fixed-order shuffle tree => deterministic
      const int nvals = p.L * cpg;
      for (int g = warp; g < p.groups; g += GN_THREADS / 32) {
        float ts = 0.f, tq = 0.f;
        for (int i = lane; i < nvals; i += 32) {
          const int ll = i / cpg, c = g * cpg + (i - ll * cpg);
          ts += s_sum[ll * C + c];
          tq += s_sq[ll * C + c];
        }
        ts = warp_sum(ts);
        tq = warp_sum(tq);
        if (lane == 0) {
          const float cnt = (float)(p_end - p_begin) * (float)cpg;
          const float mean = ts / cnt;
          part[((size_t)item * p.groups + g) * 2 + 0] = mean;
          part[((size_t)item * p.groups + g) * 2 + 1] = fmaxf(tq - ts * mean, 0.f);
        }
      }
    }
    // grid barrier among the consumer threads of all CTAs (the producers keep prefetching)
    named_bar_sync(1, GN_THREADS);
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(p.counter, 1u);
      const unsigned int target = (unsigned int)(w + 1) * gridDim.x;
      while (*reinterpret_cast<volatile unsigned int*>(p.counter) < target) __nanosleep(32);
      __threadfence();
    }
    named_bar_sync(1, GN_THREADS);
    // ---------------- phase B: merge my image's partials, normalise my slab ----------------
    if (active) {
      {
        // Chan merge of the image's slab partials: warp per group, lane i folds slabs
        // i, i+32, ... (loads issued together), then a fixed shuffle tree => deterministic
        const int first = (item / p.slabs) * p.slabs;
        for (int g = warp; g < p.groups; g += GN_THREADS / 32) {
          float2 mm[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int sidx = lane + 32 * k;
            mm[k] = make_float2(0.f, 0.f);
            if (sidx < p.slabs)
              mm[k] = __ldcg(reinterpret_cast<const float2*>(
                  part + ((size_t)(first + sidx) * p.groups + g) * 2));
          }
          float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int sidx = lane + 32 * k;
            const int pb = sidx * pps;
            const int pe = min(p.hw, pb + pps);
            if (sidx >= p.slabs || pe <= pb) continue;
            const float n_b = (float)(pe - pb) * (float)cpg;
            const float n_ab = n_a + n_b;
            const float delta = mm[k].x - mean_a;
            mean_a += delta * (n_b / n_ab);
            m2_a += mm[k].y + delta * delta * (n_a * n_b / n_ab);
            n_a = n_ab;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float n_b = __shfl_down_sync(0xffffffffu, n_a, o);
            const float mean_b = __shfl_down_sync(0xffffffffu, mean_a, o);
            const float m2_b = __shfl_down_sync(0xffffffffu, m2_a, o);
            const float n_ab = n_a + n_b;
            if (n_ab > 0.f && n_b > 0.f) {
              const float delta = mean_b - mean_a;
              mean_a += delta * (n_b / n_ab);
              m2_a += m2_b + delta * delta * (n_a * n_b / n_ab);
              n_a = n_ab;
            }
          }
          if (lane == 0) {
            s_mean[g] = mean_a;
            s_rstd[g] = rsqrtf(m2_a / n_a + p.eps);
          }
        }
      }
      named_bar_sync(2, GN_THREADS);
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = min(v * 8 + j, C - 1);
        const int g = c / cpg;
        sc[j] = s_rstd[g] * __ldg(p.gamma + c);
        sh[j] = __ldg(p.beta + c) - s_mean[g] * sc[j];
      }
      __nv_bfloat16* outv = p.out + v * 8;
      gn_consume(p, ring, full, empty, chunk_base, img, p_begin, p_end, v, l, lane_ok,
                 [&](const uint4& u, long long pix) {
                   const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
                   uint32_t o[4];
#pragma unroll
                   for (int j = 0; j < 4; ++j) {
                     float2 f = unpack_bf16x2(ww[j]);
                     f.x = f.x * sc[2 * j] + sh[2 * j];
                     f.y = f.y * sc[2 * j + 1] + sh[2 * j + 1];
                     if (p.apply_silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
                     o[j] = pack_bf16x2(f.x, f.y);
                   }
                   *reinterpret_cast<uint4*>(outv + pix * C) = make_uint4(o[0], o[1], o[2], o[3]);
                 });
    }
    named_bar_sync(1, GN_THREADS);      // s_sum / s_mean are reused by the next wave
  }
}

// ---- small images (hw <= 1024): group-sliced GroupNorm, no grid barrier -----------------------
// One CTA owns `gpc` whole groups (a contiguous channel slice) of one image, so the statistics
// never leave the CTA: pass 1 reduces the slice, pass 2 re-reads it (L1/L2 hit) and writes.
// grid = (groups/gpc, nimg); the partition depends only on (C, groups) => batch-invariant.
constexpr int GNS_THREADS = 256;

struct GnSlicedParams {
  const __nv_bfloat16* x0; int c0;
  const __nv_bfloat16* x1; int c1;
  int hw, groups, gpc, Vs, L;
  float eps;
  const float* gamma; const float* beta;
  int apply_silu;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(GNS_THREADS) gn_sliced_kernel(const GnSlicedParams p) {
  __shared__ float s_sum[GNS_THREADS * 8];
  __shared__ float s_sq[GNS_THREADS * 8];
  __shared__ float s_mean[32];
  __shared__ float s_rstd[32];
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const int Wc = p.gpc * cpg;                          // channels of this slice (multiple of 8)
  const int ch0 = blockIdx.x * Wc;
  const int img = blockIdx.y;
  const int v = threadIdx.x % p.Vs;
  const int l = threadIdx.x / p.Vs;
  const bool lane_ok = l < p.L;
  const int vg = (ch0 >> 3) + v;                        // vector index in the virtual concat
  const long long base = (long long)img * p.hw;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (lane_ok) {
    int px = l;
    for (; px + 3 * p.L < p.hw; px += 4 * p.L) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_vec8(p.x0, p.c0, p.x1, p.c1, base + px + k * p.L, vg);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ww[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(ww[j]);
          s[2 * j] += f.x; q[2 * j] += f.x * f.x;
          s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
        }
      }
    }
    for (; px < p.hw; px += p.L) {
      const uint4 u = ld_vec8(p.x0, p.c0, p.x1, p.c1, base + px, vg);
      const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(ww[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_sum[l * Wc + v * 8 + j] = s[j];
      s_sq[l * Wc + v * 8 + j] = q[j];
    }
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvals = p.L * cpg;
    for (int g = warp; g < p.gpc; g += GNS_THREADS / 32) {
      float ts = 0.f, tq = 0.f;
      for (int i = lane; i < nvals; i += 32) {
        const int ll = i / cpg, c = g * cpg + (i - ll * cpg);
        ts += s_sum[ll * Wc + c];
        tq += s_sq[ll * Wc + c];
      }
      ts = warp_sum(ts);
      tq = warp_sum(tq);
      if (lane == 0) {
        const float cnt = (float)p.hw * (float)cpg;
        const float mean = ts / cnt;
        s_mean[g] = mean;
        s_rstd[g] = rsqrtf(fmaxf(tq / cnt - mean * mean, 0.f) + p.eps);
      }
    }
  }
  __syncthreads();
  if (lane_ok) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cl = v * 8 + j;                         // channel within the slice
      const int g = cl / cpg;
      sc[j] = s_rstd[g] * __ldg(p.gamma + ch0 + cl);
      sh[j] = __ldg(p.beta + ch0 + cl) - s_mean[g] * sc[j];
    }
    auto emit = [&](const uint4& u, long long pix) {
      const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack_bf16x2(ww[j]);
        f.x = f.x * sc[2 * j] + sh[2 * j];
        f.y = f.y * sc[2 * j + 1] + sh[2 * j + 1];
        if (p.apply_silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
        o[j] = pack_bf16x2(f.x, f.y);
      }
      *reinterpret_cast<uint4*>(p.out + pix * C + ch0 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    };
    int px = l;
    for (; px + 3 * p.L < p.hw; px += 4 * p.L) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = ld_vec8(p.x0, p.c0, p.x1, p.c1, base + px + k * p.L, vg);
#pragma unroll
      for (int k = 0; k < 4; ++k) emit(u[k], base + px + k * p.L);
    }
    for (; px < p.hw; px += p.L) emit(ld_vec8(p.x0, p.c0, p.x1, p.c1, base + px, vg), base + px);
  }
}

// LayerNorm: a warp walks rows (grid-stride), each lane holds up to MAXV vectors of 8 channels
// (C <= 256*MAXV... in practice 320 / 640 / 1280 -> MAXV 2 / 3 / 5).  The next row's vectors are
// requested before the current row is reduced, so every warp keeps two rows of loads in flight
// (one row per warp left the SM at ~20 KB in flight, about half of what HBM latency needs);
// gamma / beta live in registers across rows.  Two-pass statistics in fp32 as before, so results
// are unchanged.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long rows,
                                                        int C, float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int V = C / 8;
  float gg[MAXV][8], bb[MAXV][8];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < V) {
      const int c = v * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      gg[i][0] = g0.x; gg[i][1] = g0.y; gg[i][2] = g0.z; gg[i][3] = g0.w;
      gg[i][4] = g1.x; gg[i][5] = g1.y; gg[i][6] = g1.z; gg[i][7] = g1.w;
      bb[i][0] = b0.x; bb[i][1] = b0.y; bb[i][2] = b0.z; bb[i][3] = b0.w;
      bb[i][4] = b1.x; bb[i][5] = b1.y; bb[i][6] = b1.z; bb[i][7] = b1.w;
    }
  }
  uint4 nxt[MAXV];
  auto fetch = [&](long long row) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < V) nxt[i] = __ldg(reinterpret_cast<const uint4*>(x + row * C + v * 8));
    }
  };
  if (warp0 < rows) fetch(warp0);
  for (long long row = warp0; row < rows; row += nwarps) {
    float f[MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        const uint32_t w[4] = {nxt[i].x, nxt[i].y, nxt[i].z, nxt[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_bf16x2(w[j]);
          f[i][2 * j] = t.x; f[i][2 * j + 1] = t.y;
          sum += t.x + t.y;
        }
      }
    }
    if (row + nwarps < rows) fetch(row + nwarps);         // next row in flight during the reductions
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; sq += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      if (v < V) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = pack_bf16x2((f[i][2 * j] - mean) * rstd * gg[i][2 * j] + bb[i][2 * j],
                             (f[i][2 * j + 1] - mean) * rstd * gg[i][2 * j + 1] + bb[i][2 * j + 1]);
        *reinterpret_cast<uint4*>(out + row * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---- split GroupNorm (statistics | apply) for row-strip patch-parallel execution -------------
// A rank holds a strip of rows of every image: gn_stats_kernel reduces the strip to per-(image,
// group) (mean, M2); the host all-gathers those [nimg, groups, 2] fp32 records across the ranks
// (SURVEY.md §8e exchange X3); gn_apply_kernel Chan-merges the R records in rank order (every
// rank computes bit-identical statistics) and normalises its strip, optionally writing into a
// halo-padded buffer (out_img_stride > hw*C).  Deterministic: fixed slab partition, fixed merge
// order, the last-arriving CTA of an image does the slab merge (no float atomics).
constexpr int GNX_THREADS = 512;
constexpr int GNX_MAX_SLABS = 256;

struct GnSplitParams {
  const __nv_bfloat16* x0; int c0;
  const __nv_bfloat16* x1; int c1;
  int nimg, hw, groups, V, L, slabs, pps;
  float* partial;            // [nimg][slabs][groups][2]
  float* stats;              // [nimg][groups][2]  (mean, M2) of the local strip
  unsigned int* counters;    // [nimg], zero on entry, zero on exit
  // apply side
  const float* stats_all;    // [R][nimg][groups][2]
  int R;
  float eps;
  const float* gamma; const float* beta;
  int apply_silu;
  __nv_bfloat16* out;
  long long out_img_stride;  // elements between images of `out`
};

__device__ __forceinline__ void chan_merge(float& n_a, float& mean_a, float& m2_a, float n_b,
                                           float mean_b, float m2_b) {
  const float n_ab = n_a + n_b;
  const float delta = mean_b - mean_a;
  mean_a += delta * (n_b / n_ab);
  m2_a += m2_b + delta * delta * (n_a * n_b / n_ab);
  n_a = n_ab;
}

__global__ void __launch_bounds__(GNX_THREADS) gn_stats_kernel(const GnSplitParams p) {
  __shared__ float s_sum[GNX_THREADS * 8];
  __shared__ float s_sq[GNX_THREADS * 8];
  __shared__ int s_last;
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const int img = blockIdx.y, slab = blockIdx.x;
  const int p_begin = slab * p.pps;
  const int p_end = min(p.hw, p_begin + p.pps);
  const int v = threadIdx.x % p.V, l = threadIdx.x / p.V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (l < p.L) {
    const long long base = (long long)img * p.hw;
#pragma unroll 4
    for (int px = p_begin + l; px < p_end; px += p.L) {
      const uint4 u = ld_vec8(p.x0, p.c0, p.x1, p.c1, base + px, v);
      const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(ww[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_sum[l * C + v * 8 + j] = s[j]; s_sq[l * C + v * 8 + j] = q[j]; }
  }
  __syncthreads();
  float* part = p.partial + ((size_t)img * p.slabs + slab) * p.groups * 2;
  const int nvals = p.L * cpg;
  for (int g = warp; g < p.groups; g += GNX_THREADS / 32) {
    float ts = 0.f, tq = 0.f;
    for (int i = lane; i < nvals; i += 32) {
      const int ll = i / cpg, c = g * cpg + (i - ll * cpg);
      ts += s_sum[ll * C + c];
      tq += s_sq[ll * C + c];
    }
    ts = warp_sum(ts);
    tq = warp_sum(tq);
    if (lane == 0) {
      const float cnt = (float)(p_end - p_begin) * (float)cpg;
      const float mean = cnt > 0.f ? ts / cnt : 0.f;
      part[g * 2 + 0] = mean;
      part[g * 2 + 1] = fmaxf(tq - ts * mean, 0.f);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.counters[img], 1u) == (unsigned)(p.slabs - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < p.groups) {
    const int g = threadIdx.x;
    float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
    for (int sidx = 0; sidx < p.slabs; ++sidx) {
      const int pb = sidx * p.pps, pe = min(p.hw, pb + p.pps);
      if (pe <= pb) continue;
      const float2 mm = __ldcg(reinterpret_cast<const float2*>(
          p.partial + (((size_t)img * p.slabs + sidx) * p.groups + g) * 2));
      chan_merge(n_a, mean_a, m2_a, (float)(pe - pb) * (float)cpg, mm.x, mm.y);
    }
    p.stats[((size_t)img * p.groups + g) * 2 + 0] = mean_a;
    p.stats[((size_t)img * p.groups + g) * 2 + 1] = m2_a;
  }
  if (threadIdx.x == 0) p.counters[img] = 0u;
}

__global__ void __launch_bounds__(GNX_THREADS) gn_apply_kernel(const GnSplitParams p) {
  __shared__ float s_mean[GN_MAX_GROUPS];
  __shared__ float s_rstd[GN_MAX_GROUPS];
  const int C = p.c0 + p.c1;
  const int cpg = C / p.groups;
  const int img = blockIdx.y, slab = blockIdx.x;
  const int p_begin = slab * p.pps;
  const int p_end = min(p.hw, p_begin + p.pps);
  const int v = threadIdx.x % p.V, l = threadIdx.x / p.V;
  if (threadIdx.x < p.groups) {
    const int g = threadIdx.x;
    const float n_r = (float)p.hw * (float)cpg;       // equal strips: same count on every rank
    float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
    for (int r = 0; r < p.R; ++r) {
      const float2 mm = __ldg(reinterpret_cast<const float2*>(
          p.stats_all + (((size_t)r * p.nimg + img) * p.groups + g) * 2));
      chan_merge(n_a, mean_a, m2_a, n_r, mm.x, mm.y);
    }
    s_mean[g] = mean_a;
    s_rstd[g] = rsqrtf(m2_a / n_a + p.eps);
  }
  __syncthreads();
  if (l >= p.L) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = min(v * 8 + j, C - 1);
    const int g = c / cpg;
    sc[j] = s_rstd[g] * __ldg(p.gamma + c);
    sh[j] = __ldg(p.beta + c) - s_mean[g] * sc[j];
  }
  const long long base = (long long)img * p.hw;
  __nv_bfloat16* outv = p.out + (long long)img * p.out_img_stride + v * 8;
#pragma unroll 4
  for (int px = p_begin + l; px < p_end; px += p.L) {
    const uint4 u = ld_vec8(p.x0, p.c0, p.x1, p.c1, base + px, v);
    const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(ww[j]);
      f.x = f.x * sc[2 * j] + sh[2 * j];
      f.y = f.y * sc[2 * j + 1] + sh[2 * j + 1];
      if (p.apply_silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
      o[j] = pack_bf16x2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(outv + (long long)px * C) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// (sum, sumsq) per (image, M tile, group) from the producing conv's epilogue -> (mean, M2)
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float* __restrict__ partial, int slots,
                                                          int groups, double count, float* __restrict__ stats) {
  __shared__ double sh[2][128];
  const int img = blockIdx.x / groups, g = blockIdx.x % groups;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < slots; i += 128) {
    const float2 r = __ldg(reinterpret_cast<const float2*>(partial + (((long long)img * slots + i) * groups + g) * 2));
    s += (double)r.x;
    q += (double)r.y;
  }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {                      // fixed tree: deterministic
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = sh[0][0] / count;
    const double m2 = sh[1][0] - sh[0][0] * mean;
    stats[((long long)img * groups + g) * 2 + 0] = (float)mean;
    stats[((long long)img * groups + g) * 2 + 1] = (float)(m2 > 0.0 ? m2 : 0.0);
  }
}

// per-CHANNEL (sum, sumsq) records of one or two producers (channel concat [x0 | x1]) -> (mean, M2) per
// (image, group).  A group may straddle the two sources (C0 = 1280, C1 = 640, 60 channels per group).
// partial_k: fp32 [nimg][slots_k][C_k][2].  One CTA per (image, group), fp64 fixed-tree sum: deterministic.
__global__ void __launch_bounds__(128) gn_finalize_chan_kernel(const float* __restrict__ part0, int slots0, int c0,
                                                               const float* __restrict__ part1, int slots1, int c1,
                                                               int groups, double count, float* __restrict__ stats) {
  __shared__ double sh[2][128];
  const int img = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = (c0 + c1) / groups;
  const int smax = slots0 > slots1 ? slots0 : slots1;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < smax * cpg; i += 128) {
    const int slot = i / cpg, c = g * cpg + (i - slot * cpg);
    float2 r = make_float2(0.f, 0.f);
    if (c < c0) {
      if (slot < slots0) r = __ldg(reinterpret_cast<const float2*>(part0 + (((long long)img * slots0 + slot) * c0 + c) * 2));
    } else if (slot < slots1) {
      r = __ldg(reinterpret_cast<const float2*>(part1 + (((long long)img * slots1 + slot) * c1 + (c - c0)) * 2));
    }
    s += (double)r.x;
    q += (double)r.y;
  }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = sh[0][0] / count;
    const double m2 = sh[1][0] - sh[0][0] * mean;
    stats[((long long)img * groups + g) * 2 + 0] = (float)mean;
    stats[((long long)img * groups + g) * 2 + 1] = (float)(m2 > 0.0 ? m2 : 0.0);
  }
}

static int gn_split_setup(GnSplitParams& p, const void* x0, int c0, const void* x1, int c1, int nimg,
                          int hw, int groups) {
  const int C = c0 + c1;
  DL_CHECK_ARG(x0 && nimg > 0 && hw > 0, "groupnorm(split): bad args");
  DL_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && C > 0, "groupnorm(split): channels must be multiples of 8");
  DL_CHECK_ARG(c1 == 0 || x1, "groupnorm(split): c1>0 needs x1");
  DL_CHECK_ARG(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0,
               "groupnorm(split): bad groups=%d for C=%d", groups, C);
  DL_CHECK_ARG(C / 8 <= GNX_THREADS, "groupnorm(split): C=%d too wide", C);
  memset(&p, 0, sizeof(p));
  p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); p.c0 = c0;
  p.x1 = reinterpret_cast<const __nv_bfloat16*>(x1); p.c1 = c1;
  p.nimg = nimg; p.hw = hw; p.groups = groups;
  p.V = C / 8;
  p.L = GNX_THREADS / p.V;
  // slab partition: a function of (hw, C) only => batch-invariant statistics
  int slabs = (int)(((long long)hw * C * 2 + (1 << 17) - 1) >> 17);      // ~128 KB of input per CTA
  const int max_slabs = (hw + p.L - 1) / p.L;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs > GNX_MAX_SLABS) slabs = GNX_MAX_SLABS;
  if (slabs < 1) slabs = 1;
  p.pps = (hw + slabs - 1) / slabs;
  p.slabs = (hw + p.pps - 1) / p.pps;
  return 0;
}

}  // namespace dl

extern "C" int dl_groupnorm_finalize(const float* partial, int nimg, int slots, int groups, long long count,
                                     float* stats, void* stream_) {
  using namespace dl;
  DL_CHECK_ARG(partial && stats && nimg > 0 && slots > 0 && groups > 0 && count > 0, "groupnorm_finalize: bad args");
  gn_finalize_kernel<<<nimg * groups, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(partial, slots, groups,
                                                                                       (double)count, stats);
  return check_launch("groupnorm_finalize");
}

extern "C" int dl_groupnorm_finalize_channels(const float* part0, int slots0, int c0, const float* part1, int slots1,
                                              int c1, int nimg, int groups, long long count, float* stats,
                                              void* stream_) {
  using namespace dl;
  DL_CHECK_ARG(part0 && stats && nimg > 0 && slots0 > 0 && c0 > 0 && groups > 0 && count > 0,
               "groupnorm_finalize_channels: bad args");
  DL_CHECK_ARG(c1 == 0 || (part1 && slots1 > 0), "groupnorm_finalize_channels: c1 > 0 needs part1 / slots1");
  DL_CHECK_ARG((c0 + c1) % groups == 0, "groupnorm_finalize_channels: C=%d not divisible by groups=%d", c0 + c1, groups);
  gn_finalize_chan_kernel<<<nimg * groups, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      part0, slots0, c0, part1, c1 > 0 ? slots1 : 0, c1, groups, (double)count, stats);
  return check_launch("groupnorm_finalize_channels");
}

extern "C" size_t dl_groupnorm_split_workspace_bytes(int nimg, int groups) {
  return 4096 + (size_t)(nimg > 0 ? nimg : 1) * dl::GNX_MAX_SLABS * (groups > 0 ? groups : 32) * 2 * sizeof(float);
}

extern "C" int dl_groupnorm_stats(const void* x0, int c0, const void* x1, int c1, int nimg, int hw,
                                  int groups, float* stats, void* workspace, void* stream_) {
  using namespace dl;
  GnSplitParams p;
  if (int rc = gn_split_setup(p, x0, c0, x1, c1, nimg, hw, groups)) return rc;
  DL_CHECK_ARG(stats && workspace, "groupnorm_stats: null pointer");
  DL_CHECK_ARG(nimg <= 1024, "groupnorm_stats: nimg=%d exceeds 1024", nimg);
  p.stats = stats;
  p.counters = reinterpret_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 4096);
  gn_stats_kernel<<<dim3(p.slabs, nimg), GNX_THREADS, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(p);
  return check_launch("groupnorm_stats");
}

extern "C" int dl_groupnorm_apply(const void* x0, int c0, const void* x1, int c1, int nimg, int hw,
                                  int groups, float eps, const float* gamma, const float* beta,
                                  int apply_silu, const float* stats_all, int nranks, void* out,
                                  long long out_img_stride, void* stream_) {
  using namespace dl;
  GnSplitParams p;
  if (int rc = gn_split_setup(p, x0, c0, x1, c1, nimg, hw, groups)) return rc;
  DL_CHECK_ARG(stats_all && out && gamma && beta && nranks >= 1, "groupnorm_apply: bad args");
  const long long dense = (long long)hw * (c0 + c1);
  DL_CHECK_ARG(out_img_stride == 0 || (out_img_stride >= dense && out_img_stride % 8 == 0),
               "groupnorm_apply: bad out_img_stride");
  p.stats_all = stats_all; p.R = nranks; p.eps = eps; p.gamma = gamma; p.beta = beta;
  p.apply_silu = apply_silu;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.out_img_stride = out_img_stride > 0 ? out_img_stride : dense;
  gn_apply_kernel<<<dim3(p.slabs, nimg), GNX_THREADS, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(p);
  return check_launch("groupnorm_apply");
}

extern "C" size_t dl_groupnorm_workspace_bytes(int nimg, int groups) {
  (void)nimg;
  // barrier counter (256 B slot) + double-buffered partials for up to 1024 CTAs
  return 256 + (size_t)2 * 1024 * (groups > 0 ? groups : 32) * 2 * sizeof(float);
}

extern "C" int dl_groupnorm(const void* x0, int c0, const void* x1, int c1, int nimg, int hw,
                            int groups, float eps, const float* gamma, const float* beta,
                            int apply_silu, void* out, void* workspace, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int C = c0 + c1;
  DL_CHECK_ARG(x0 && out && workspace && gamma && beta, "groupnorm: null pointer");
  DL_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && C > 0, "groupnorm: channels must be multiples of 8");
  DL_CHECK_ARG(c1 == 0 || x1, "groupnorm: c1>0 needs x1");
  DL_CHECK_ARG(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0,
               "groupnorm: bad groups=%d for C=%d", groups, C);
  DL_CHECK_ARG(C <= GN_MAX_C, "groupnorm: C=%d exceeds %d", C, GN_MAX_C);
  if (hw <= 1024) {
    // small images: group-sliced kernel (no grid barrier, ordinary launch)
    const int cpg = C / groups;
    int gpc = 0;
    for (int g = 2; g <= groups; g <<= 1)
      if (groups % g == 0 && (g * cpg) % 8 == 0 && g * cpg <= GNS_THREADS * 8 / 1) { gpc = g; break; }
    if (gpc == 0 && cpg % 8 == 0) gpc = 1;
    if (gpc > 0 && (gpc * cpg) / 8 <= GNS_THREADS) {
      GnSlicedParams sp;
      sp.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); sp.c0 = c0;
      sp.x1 = reinterpret_cast<const __nv_bfloat16*>(x1); sp.c1 = c1;
      sp.hw = hw; sp.groups = groups; sp.gpc = gpc;
      sp.Vs = gpc * cpg / 8;
      sp.L = GNS_THREADS / sp.Vs;
      sp.eps = eps; sp.gamma = gamma; sp.beta = beta; sp.apply_silu = apply_silu;
      sp.out = reinterpret_cast<__nv_bfloat16*>(out);
      gn_sliced_kernel<<<dim3(groups / gpc, nimg), GNS_THREADS, 0, stream>>>(sp);
      return check_launch("groupnorm(sliced)");
    }
  }
  GnParams p;
  p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); p.c0 = c0;
  p.x1 = reinterpret_cast<const __nv_bfloat16*>(x1); p.c1 = c1;
  p.nimg = nimg; p.hw = hw; p.groups = groups;
  p.V = C / 8;
  DL_CHECK_ARG(p.V <= GN_THREADS, "groupnorm: C=%d too wide", C);
  p.L = GN_THREADS / p.V;
  {
    int P = GN_CHUNK_BYTES / (C * 2);                 // pixels per 16 KB chunk
    P = (P / p.L) * p.L;
    if (P < p.L) P = p.L;
    DL_CHECK_ARG((long long)P * C * 2 <= GN_CHUNK_BYTES, "groupnorm: chunk does not fit (C=%d)", C);
    p.P = P;
  }
  p.eps = eps; p.gamma = gamma; p.beta = beta; p.apply_silu = apply_silu;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.counter = reinterpret_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  int grid = num_sms();                       // 1 CTA/SM (launch bounds + 33 KB smem): co-resident
  if (grid > 1024) grid = 1024;
  // Slabs per image depend ONLY on the image size (never on the batch): a request normalises
  // bit-identically alone or inside any batch (reduction tree shape is a function of hw, C).
  const long long img_bytes = (long long)hw * C * 2;
  long long slabs = ((long long)grid * img_bytes + GN_WAVE_BYTES / 2) / GN_WAVE_BYTES;
  const int max_slabs = (hw + p.L - 1) / p.L;          // at least one pixel per lane row
  if (slabs > grid) slabs = grid;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  if (slabs > 160) slabs = 160;                       // merge step folds <= 5 x 32 partials per group
  p.slabs = (int)slabs;
  long long wave = grid / p.slabs;                     // images that fit one wave of CTAs
  if (wave > nimg) wave = nimg;
  const long long n_waves = (nimg + wave - 1) / wave;
  wave = (nimg + n_waves - 1) / n_waves;               // balance the waves (16 -> 8+8, not 15+1)
  p.wave_imgs = (int)wave;
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), stream);
  if (e != cudaSuccess) { set_error("groupnorm: memset: %s", cudaGetErrorString(e)); return 2; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GN_THREADS + 32);            // 16 consumer warps + 1 producer warp
  const size_t dyn_smem = (size_t)GN_STAGES * GN_CHUNK_BYTES + (size_t)2 * GN_THREADS * 8 * sizeof(float);
  {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
      cudaError_t ea = cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)dyn_smem);
      if (ea != cudaSuccess) { set_error("groupnorm: cudaFuncSetAttribute: %s", cudaGetErrorString(ea)); return 1; }
      attr_set[dev & 63] = true;
    }
  }
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // guarantees co-residency for the grid barrier
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, gn_fused_kernel, p);
  if (e != cudaSuccess) { set_error("groupnorm: launch failed: %s", cudaGetErrorString(e)); return 2; }
  return check_launch("groupnorm");
}

extern "C" int dl_layernorm(const void* x, long long rows, int c, float eps, const float* gamma,
                            const float* beta, void* out, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DL_CHECK_ARG(x && out && gamma && beta, "layernorm: null pointer");
  DL_CHECK_ARG(c % 8 == 0 && c > 0 && c <= 1280, "layernorm: C=%d must be a multiple of 8, <= 1280", c);
  const int wpb = 8;
  long long blocks = (rows + wpb - 1) / wpb;
  const long long cap = (long long)num_sms() * 8;        // resident blocks: the warps stride over rows
  if (blocks > cap) blocks = cap;
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  if (c <= 512) layernorm_kernel<2><<<(unsigned)blocks, wpb * 32, 0, stream>>>(xi, rows, c, eps, gamma, beta, xo);
  else if (c <= 768) layernorm_kernel<3><<<(unsigned)blocks, wpb * 32, 0, stream>>>(xi, rows, c, eps, gamma, beta, xo);
  else layernorm_kernel<5><<<(unsigned)blocks, wpb * 32, 0, stream>>>(xi, rows, c, eps, gamma, beta, xo);
  return check_launch("layernorm");
}
