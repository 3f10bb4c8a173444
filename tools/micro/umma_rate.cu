// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N and of
// where the operands come from — SS (A and B in shared memory, 128-byte swizzle, K-major, the layout the conv / linear
// and attention kernels use) and TS (A in TMEM).  One CTA per SM, one issuing thread, batches of MMAs closed by a commit.
// The floor is 128 * N / 256 clk per MMA; what SS adds on top is shared-memory operand fetch (A: 4 KB, B: N * 32 B per MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I <csrc> umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace dl;

// mode 0: SS, A walks 8 chunks x 4 K-steps of a 128-row tile (128 KB), B walks the same of an N-row tile
// mode 1: TS, A = 8 packed TMEM columns, B as above
// mode 2: SS with B MN-major (the PV layout): B = 16 keys x N columns, chunks of 64 columns
__global__ void __launch_bounds__(128, 1) k_rate(int n, int mode, int batches, int per_batch, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512u); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  uint8_t* sA = smem;                       // 8 chunks x 128 rows x 128 B = 128 KB
  uint8_t* sB = smem + 8 * 16384;           // up to 8 chunks x n rows x 128 B (n <= 64 for the full 8, see host)
  const int bchunk = n * 128;
  int nbch = 1;                             // power of two, so the issuing thread's index math stays trivial
  while (mode != 2 && nbch < 8 && nbch * 2 * bchunk <= 96 * 1024) nbch *= 2;
  if (warp == 1) {
    const bool issuer = elect_one();
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA));
    const uint32_t b_lo = (mode == 2) ? umma_desc_lo(smem_u32(sB), 64 * 128) : umma_desc_lo(smem_u32(sB));
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)n, 0, mode == 2 ? 1 : 0);
    long long t0 = clock64();
    for (int bt = 0; bt < batches; ++bt) {
      if (issuer) {
        for (int i = 0; i < per_batch; ++i) {
          const int c = (i >> 2) & 7, ks = i & 3;
          const uint32_t aoff = (uint32_t)(c * (16384 >> 4) + ks * 2);
          const uint32_t boff = (mode == 2) ? (uint32_t)((i & 3) * 128) : (uint32_t)((c & (nbch - 1)) * (bchunk >> 4) + ks * 2);
          if (mode == 1) umma_ts_lohi(base + 256u, base + (uint32_t)((i & 7) * 8), b_lo + boff, hi, idesc, 1u);
          else umma_ss_lohi(base + 256u, a_lo + aoff, b_lo + boff, hi, idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)(bt & 1));
    }
    long long t1 = clock64();
    if (issuer) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512u);
}

// MMA (SS, K-major, 4 A chunks + 2..4 B chunks) with the TMA engine streaming bulk copies into a ring of the same CTA's
// shared memory at the same time: do the operand reads of the tensor core and the fills of the copy engine share
// one shared-memory port?
__global__ void __launch_bounds__(128, 1) k_mix(int n, int do_mma, int do_tma, int n_mma, int slot_bytes,
                                                const uint8_t* __restrict__ src, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, ring_bar[2];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&ring_bar[0], 1); mbar_init(&ring_bar[1], 1); done = 0; fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512u); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  uint8_t* sA = smem;                       // 4 chunks x 16 KB
  uint8_t* sB = smem + 64 * 1024;           // 64 KB
  uint8_t* ring = smem + 128 * 1024;        // 2 slots x slot_bytes (<= 48 KB each)
  const int bchunk = n * 128;
  int nbch = 1;
  while (nbch < 4 && nbch * 2 * bchunk <= 64 * 1024) nbch *= 2;
  if (warp == 1 && do_mma) {
    const bool issuer = elect_one();
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA));
    const uint32_t b_lo = umma_desc_lo(smem_u32(sB));
    const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)n, 0, 0);
    long long t0 = clock64();
    for (int bt = 0; bt < n_mma / 256; ++bt) {
      if (issuer) {
        for (int i = 0; i < 256; ++i) {
          const int c = (i >> 2) & 3, ks = i & 3;
          umma_ss_lohi(base + 256u, a_lo + (uint32_t)(c * 1024 + ks * 2),
                       b_lo + (uint32_t)((c & (nbch - 1)) * (bchunk >> 4) + ks * 2), hi, idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)(bt & 1));
    }
    long long t1 = clock64();
    if (issuer) { out[blockIdx.x * 4 + 0] = t1 - t0; done = 1; }
  }
  if (warp == 2 && do_tma) {
    const bool issuer = elect_one();
    long long t0 = clock64();
    long long copies = 0;
    const uint8_t* my = src + (size_t)(blockIdx.x % 64) * (2u << 20);       // 2 MB window per CTA, L2-resident
    if (issuer) {
      mbar_expect_tx(&ring_bar[0], (uint32_t)slot_bytes);
      bulk_load_1d(ring, my, (uint32_t)slot_bytes, &ring_bar[0]);
    }
    for (int it = 0;; ++it) {
      const int nxt = (it + 1) & 1;
      if (issuer) {
        mbar_expect_tx(&ring_bar[nxt], (uint32_t)slot_bytes);
        bulk_load_1d(ring + nxt * slot_bytes, my + (size_t)((it + 1) & 31) * slot_bytes, (uint32_t)slot_bytes, &ring_bar[nxt]);
      }
      __syncwarp();
      mbar_wait(&ring_bar[it & 1], (uint32_t)((it >> 1) & 1));
      ++copies;
      if (do_mma ? done : (copies >= n_mma / 4)) break;
    }
    long long t1 = clock64();
    mbar_wait(&ring_bar[copies & 1], (uint32_t)((copies >> 1) & 1));        // drain the copy still in flight
    if (issuer) { out[blockIdx.x * 4 + 1] = t1 - t0; out[blockIdx.x * 4 + 2] = copies * slot_bytes; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512u);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  long long h[148];
  const int smem_bytes = 226 * 1024;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int batches = 50, per_batch = 256;
  for (int mode = 0; mode < 3; ++mode)
    for (int n : {32, 64, 96, 128, 192, 256}) {
      if (mode == 2 && n % 64) continue;
      for (int grid : {1, 148}) {
        k_rate<<<grid, 128, smem_bytes>>>(n, mode, batches, per_batch, out);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double clk = (double)mx / (batches * per_batch);
        const double abytes = mode == 1 ? 0.0 : 4096.0, bbytes = n * 32.0;
        printf("%s N=%3d grid %3d: %6.1f clk/MMA (floor %5.1f)  smem operand bytes %5.0f -> %6.1f B/clk\n",
               mode == 0 ? "SS K-major " : mode == 1 ? "TS K-major " : "SS B MN-maj", n, grid, clk, 128.0 * n / 256.0,
               abytes + bbytes, (abytes + bbytes) / clk);
      }
    }
  {
    uint8_t* src;
    cudaMalloc(&src, 128u << 20);
    cudaMemset(src, 1, 128u << 20);
    long long* o4;
    cudaMalloc(&o4, 148 * 4 * 8);
    long long h4[148 * 4];
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const int n_mma = 256 * 40;
    for (int n : {64, 128, 256})
      for (int slot_kb : {16, 48})
        for (int cfg = 0; cfg < 3; ++cfg) {       // 0: MMA only, 1: TMA only, 2: both
          const int do_mma = cfg != 1, do_tma = cfg != 0;
          cudaMemset(o4, 0, sizeof(h4));
          k_mix<<<148, 128, smem_bytes>>>(n, do_mma, do_tma, n_mma, slot_kb * 1024, src, o4);
          cudaError_t e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h4, o4, sizeof(h4), cudaMemcpyDeviceToHost);
          double mma_clk = 0, tma_rate = 0;
          for (int i = 0; i < 148; ++i) {
            mma_clk += (double)h4[i * 4] / n_mma / 148;
            if (h4[i * 4 + 1] > 0) tma_rate += (double)h4[i * 4 + 2] / (double)h4[i * 4 + 1] / 148;
          }
          const double rd = do_mma ? (4096.0 + n * 32.0) / mma_clk : 0.0;
          printf("mix N=%3d ring slot %2d KB %-8s: %6.1f clk/MMA (operand reads %6.1f B/clk)  TMA fill %6.1f B/clk/SM  sum %6.1f\n", n, slot_kb,
                 cfg == 0 ? "MMA only" : cfg == 1 ? "TMA only" : "both", mma_clk, rd, tma_rate, rd + tma_rate);
        }
  }
  return 0;
}
