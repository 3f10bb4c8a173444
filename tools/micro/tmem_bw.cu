// Micro-benchmark: TMEM read (tcgen05.ld 32x32b.x32) and write (tcgen05.st) bandwidth per SM as a function of how
// many warps issue at once.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I <csrc> tmem_bw.cu -o tmem_bw
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace dl;

__global__ void __launch_bounds__(512, 1) k_ld(int nwarps, int iters, long long* out, uint32_t* sink, int mode) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512u); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {          // 4 x ld32 (128 columns) then one wait: the attention softmax pattern
        uint32_t a[32], b[32], c[32], d[32];
        const uint32_t col = (uint32_t)(((warp >> 2) & 1) * 128);
        tmem_ld32(base + lane_off + col, a); tmem_ld32(base + lane_off + col + 32, b);
        tmem_ld32(base + lane_off + col + 64, c); tmem_ld32(base + lane_off + col + 96, d);
        tmem_ld_wait();
        acc ^= a[0] ^ b[5] ^ c[9] ^ d[31];
      } else if (mode == 1) {   // ld32 + wait (latency chain)
        uint32_t a[32];
        tmem_ld32(base + lane_off + (uint32_t)((it & 3) * 32), a);
        tmem_ld_wait();
        acc ^= a[3];
      } else {                  // 4 x st16 (64 packed columns) + wait: the P store pattern
        uint32_t a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = acc + i;
        const uint32_t col = 256u + (uint32_t)(((warp >> 2) & 1) * 64);
        tmem_st16(base + lane_off + col, a); tmem_st16(base + lane_off + col + 16, a);
        tmem_st16(base + lane_off + col + 32, a); tmem_st16(base + lane_off + col + 48, a);
        tmem_st_wait();
        acc += 1;
      }
    }
  }
  long long t1 = clock64();
  if (warp < nwarps && (threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512u);
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4);
  long long h[148 * 16];
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 2, 4, 8, 16}) {
      k_ld<<<148, 512>>>(nw, iters, out, sink, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      const double bytes = mode == 0 ? 128.0 * 128 : (mode == 1 ? 32.0 * 128 : 64.0 * 128);   // per warp per iter (32 lanes x cols x 4 B)
      printf("mode %d (%s) warps %2d: %8.1f clk/iter/warp  -> %7.1f B/clk/SM\n", mode,
             mode == 0 ? "4x ld32 + wait" : mode == 1 ? "ld32 + wait   " : "4x st16 + wait", nw, (double)mx / iters,
             bytes * nw * iters / (double)mx);
    }
  return 0;
}
