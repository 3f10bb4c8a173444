// CUDA-core attention (DL_ATTN_SIMT): a simple, slow, obviously-correct flash-style kernel used
// as the on-device checker for the tcgen05 kernel at full sizes (tests only; the product path
// uses DL_ATTN_TC).  One warp per query row, 32 keys per step, online softmax in fp32.
#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int SA_WARPS = 8;
constexpr int SA_KEYS = 32;
constexpr int SA_MAXD = 512;

__global__ void __launch_bounds__(SA_WARPS * 32)
attn_simt_kernel(const __nv_bfloat16* __restrict__ q, long long ldq,
                 const __nv_bfloat16* __restrict__ k, long long ldk,
                 const __nv_bfloat16* __restrict__ v, long long ldv, int dh_stride,
                 __nv_bfloat16* __restrict__ out, long long ldo, int sq, int skv, int d,
                 float scale, int causal) {
  extern __shared__ float sm[];
  const int dp = d + 1;                          // padded row: conflict-free column reads
  float* sK = sm;                                // [32][dp]
  float* sV = sK + SA_KEYS * dp;                 // [32][dp]
  float* sQ = sV + SA_KEYS * dp;                 // [SA_WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * SA_WARPS + warp;
  const bool active = qi < sq;
  const __nv_bfloat16* qrow = q + ((long long)b * sq + (active ? qi : 0)) * ldq + h * dh_stride;
  for (int i = lane; i < d; i += 32) sQ[warp * d + i] = __bfloat162float(qrow[i]) * scale;
  constexpr int NACC = SA_MAXD / 32;
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < skv; k0 += SA_KEYS) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < SA_KEYS * d; idx += blockDim.x) {
      const int j = idx / d, i = idx % d;
      const bool ok = (k0 + j) < skv;
      const long long r = (long long)b * skv + k0 + j;
      sK[j * dp + i] = ok ? __bfloat162float(k[r * ldk + h * dh_stride + i]) : 0.f;
      sV[j * dp + i] = ok ? __bfloat162float(v[r * ldv + h * dh_stride + i]) : 0.f;
    }
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += sQ[warp * d + i] * sK[lane * dp + i];
    if (k0 + lane >= skv || (causal && k0 + lane > qi)) s = -INFINITY;   // causal: keys <= query
    const float m_new = fmaxf(m, warp_max(s));
    if (m_new == -INFINITY) continue;                    // a whole key block in the masked future
    const float p = __expf(s - m_new);
    const float corr = __expf(m - m_new);
    l = l * corr + warp_sum(p);
    m = m_new;
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] *= corr;
    for (int j = 0; j < SA_KEYS; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        const int dim = lane + 32 * i;
        if (dim < d) acc[i] += pj * sV[j * dp + dim];
      }
    }
  }
  if (active) {
    __nv_bfloat16* orow = out + ((long long)b * sq + qi) * ldo + h * d;
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      const int dim = lane + 32 * i;
      if (dim < d) orow[dim] = __float2bfloat16(acc[i] * inv);
    }
  }
}

int attn_simt_launch(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                     long long ldv, int dh_stride, void* out, long long ldo, int batch, int sq,
                     int skv, int heads, int d, float scale, int causal, cudaStream_t stream) {
  DL_CHECK_ARG(d <= SA_MAXD, "attention(simt): d=%d exceeds %d", d, SA_MAXD);
  const size_t smem = (size_t)(2 * SA_KEYS * (d + 1) + SA_WARPS * d) * sizeof(float);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev & 63] = true;
  }
  dim3 grid((sq + SA_WARPS - 1) / SA_WARPS, heads, batch);
  attn_simt_kernel<<<grid, SA_WARPS * 32, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(k), ldk,
      reinterpret_cast<const __nv_bfloat16*>(v), ldv, dh_stride,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, sq, skv, d, scale, causal);
  return check_launch("attention(simt)");
}

}  // namespace dl
