// GroupNorm(+SiLU) and LayerNorm for NHWC bf16 activations (HBM-bound kernels).
//
// GroupNorm is two deterministic launches (no float atomics: same seed => same bytes):
//   1. gn_stats: each CTA reduces a slab of pixels of one image to per-group (mean, M2)
//      partials (fp32 sums inside the slab, Chan-merged across slabs later);
//   2. gn_apply: merges the partials (fixed order), builds per-channel scale/shift in smem,
//      then streams x -> y with 16-byte vector loads/stores, SiLU fused.
// Both accept two sources and emit the channel concat [x0 | x1] (UNet skip concat folded
// into the norm: the concat tensor is never written in un-normalised form).
#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int GN_MAX_CHUNKS = 128;
constexpr int GN_MAX_C = 2560;

__device__ __forceinline__ uint4 ld_vec8(const __nv_bfloat16* x0, int c0,
                                         const __nv_bfloat16* x1, int c1, long long pix, int v) {
  // vector v covers channels [8v, 8v+8) of the virtual concat
  const int ch = v * 8;
  const __nv_bfloat16* p = (ch < c0) ? (x0 + pix * c0 + ch) : (x1 + pix * c1 + (ch - c0));
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// grid: (chunks, nimg); block: V*L threads (V = C/8 vectors per pixel, L pixel lanes)
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x0, int c0,
                                const __nv_bfloat16* __restrict__ x1, int c1, int hw, int groups,
                                int V, int L, float* __restrict__ partial /* [nimg][chunks][G][2] */) {
  extern __shared__ float sm[];            // [L][C] sums, [L][C] sumsq
  const int C = c0 + c1;
  const int cpg = C / groups;
  const int img = blockIdx.y;
  const int chunks = gridDim.x;
  const int ppc = (hw + chunks - 1) / chunks;
  const int p_begin = blockIdx.x * ppc;
  const int p_end = min(hw, p_begin + ppc);
  const int v = threadIdx.x % V;
  const int l = threadIdx.x / V;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  for (int p = p_begin + l; p < p_end; p += L) {
    const uint4 u = ld_vec8(x0, c0, x1, c1, (long long)img * hw + p, v);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s[2 * j] += f.x; q[2 * j] += f.x * f.x;
      s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
    }
  }
  float* ssum = sm;
  float* ssq = sm + (size_t)L * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ssum[(size_t)l * C + v * 8 + j] = s[j];
    ssq[(size_t)l * C + v * 8 + j] = q[j];
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    float ts = 0.f, tq = 0.f;
    for (int ll = 0; ll < L; ++ll)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        ts += ssum[(size_t)ll * C + c];
        tq += ssq[(size_t)ll * C + c];
      }
    const float cnt = (float)(p_end - p_begin) * (float)cpg;
    float mean = 0.f, m2 = 0.f;
    if (cnt > 0.f) {
      mean = ts / cnt;
      m2 = fmaxf(tq - ts * mean, 0.f);
    }
    float* o = partial + (((size_t)img * chunks + blockIdx.x) * groups + g) * 2;
    o[0] = mean;
    o[1] = m2;
  }
}

// grid: (pixel blocks, nimg); block 256
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int c0,
                                const __nv_bfloat16* __restrict__ x1, int c1, int hw, int groups,
                                int chunks, float eps, const float* __restrict__ partial,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                int apply_silu, __nv_bfloat16* __restrict__ out) {
  __shared__ float s_scale[GN_MAX_C];
  __shared__ float s_shift[GN_MAX_C];
  __shared__ float s_mean[64];
  __shared__ float s_rstd[64];
  const int C = c0 + c1;
  const int cpg = C / groups;
  const int img = blockIdx.y;
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    const int ppc = (hw + chunks - 1) / chunks;
    // Chan et al. pairwise merge, fixed chunk order (deterministic)
    float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      const int pb = ch * ppc;
      const int pe = min(hw, pb + ppc);
      if (pe <= pb) break;
      const float n_b = (float)(pe - pb) * (float)cpg;
      const float* pp = partial + (((size_t)img * chunks + ch) * groups + g) * 2;
      const float mean_b = pp[0], m2_b = pp[1];
      const float n_ab = n_a + n_b;
      const float delta = mean_b - mean_a;
      mean_a += delta * (n_b / n_ab);
      m2_a += m2_b + delta * delta * (n_a * n_b / n_ab);
      n_a = n_ab;
    }
    s_mean[g] = mean_a;
    s_rstd[g] = rsqrtf(m2_a / n_a + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * gamma[c];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - s_mean[g] * sc;
  }
  __syncthreads();
  const int V = C / 8;
  const long long total = (long long)hw * V;
  const long long per_block = (total + gridDim.x - 1) / gridDim.x;
  const long long begin = (long long)blockIdx.x * per_block;
  const long long end = min(total, begin + per_block);
  for (long long i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const long long p = i / V;
    const int v = (int)(i - p * V);
    const long long pix = (long long)img * hw + p;
    const uint4 u = ld_vec8(x0, c0, x1, c1, pix, v);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = unpack_bf16x2(w[j]);
      const int c = v * 8 + 2 * j;
      f.x = f.x * s_scale[c] + s_shift[c];
      f.y = f.y * s_scale[c + 1] + s_shift[c + 1];
      if (apply_silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
      o[j] = pack_bf16x2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(out + pix * C + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one warp per row; C <= 1280 (multiple of 8): each lane holds up to 5 vectors of 8
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C,
                                 float eps, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int V = C / 8;
  constexpr int MAXV = 5;
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < V) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * C + v * 8));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2(w[j]);
        f[i][2 * j] = t.x; f[i][2 * j + 1] = t.y;
        sum += t.x + t.y;
      }
    }
  }
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
      const int c = v * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = pack_bf16x2((f[i][2 * j] - mean) * rstd * gg[2 * j] + bb[2 * j],
                           (f[i][2 * j + 1] - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1]);
      *reinterpret_cast<uint4*>(out + row * C + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

static int gn_chunks(int hw) {
  int chunks = (hw + 255) / 256;          // >= 256 pixels per slab
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  if (chunks < 1) chunks = 1;
  return chunks;
}

}  // namespace dl

extern "C" size_t dl_groupnorm_workspace_bytes(int nimg, int groups) {
  return (size_t)nimg * dl::GN_MAX_CHUNKS * groups * 2 * sizeof(float);
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
  DL_CHECK_ARG(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: bad groups=%d for C=%d", groups, C);
  DL_CHECK_ARG(C <= GN_MAX_C, "groupnorm: C=%d exceeds %d", C, GN_MAX_C);
  const int V = C / 8;
  int L = 256 / V;
  if (L < 1) L = 1;
  const int threads = V * L;
  DL_CHECK_ARG(threads <= 1024 && threads >= groups, "groupnorm: unsupported C=%d", C);
  const int chunks = gn_chunks(hw);
  const size_t smem = (size_t)2 * L * C * sizeof(float);
  DL_CHECK_ARG(smem <= 48 * 1024, "groupnorm: smem %zu too large", smem);
  gn_stats_kernel<<<dim3(chunks, nimg), threads, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x0), c0, reinterpret_cast<const __nv_bfloat16*>(x1), c1,
      hw, groups, V, L, reinterpret_cast<float*>(workspace));
  if (int e = check_launch("gn_stats")) return e;
  // enough CTAs to fill 148 SMs a few times over, >= 16 KB of work each
  const long long total_vec = (long long)hw * V;
  long long blocks = (total_vec + 2047) / 2048;
  const long long max_blocks = (long long)(num_sms() * 8 + nimg - 1) / nimg;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  gn_apply_kernel<<<dim3((unsigned)blocks, nimg), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x0), c0, reinterpret_cast<const __nv_bfloat16*>(x1), c1,
      hw, groups, chunks, eps, reinterpret_cast<const float*>(workspace), gamma, beta, apply_silu,
      reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("gn_apply");
}

extern "C" int dl_layernorm(const void* x, long long rows, int c, float eps, const float* gamma,
                            const float* beta, void* out, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DL_CHECK_ARG(x && out && gamma && beta, "layernorm: null pointer");
  DL_CHECK_ARG(c % 8 == 0 && c > 0 && c <= 1280, "layernorm: C=%d must be a multiple of 8, <= 1280", c);
  const int wpb = 8;
  const long long blocks = (rows + wpb - 1) / wpb;
  layernorm_kernel<<<(unsigned)blocks, wpb * 32, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), rows, c, eps, gamma, beta,
      reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("layernorm");
}
