// All-gather over NVLink peer memory (SURVEY.md §8e exchanges X1-X3 of the SDXL patch-parallel
// path): ONE kernel per exchange instead of pack + ncclAllGather + unpack.
//
// Every rank owns a symmetric staging buffer (torch symmetric memory: the same allocation mapped
// into every peer) of two halves x R slots, and a signal pad with one 32-bit flag per peer.
//   1. each CTA stores its share of the local message into slot `rank` of EVERY peer's staging
//      half (plain 16-byte stores through the peer mapping: NVLink writes);
//   2. the last CTA to finish publishes the epoch into flag[rank] of every peer (release, system
//      scope) — the message is complete in all peers before anybody can see the flag;
//   3. every CTA waits until its own pad shows the epoch for all R ranks (acquire), then copies
//      the gathered half into the caller's ordinary destination tensor.
// Epochs live on the device and advance once per call, so a captured CUDA graph replays
// correctly; halves alternate with the epoch: when a rank has seen everybody's flag for call
// k+1, every peer has finished (stream order) the kernels that read call k's half.
#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int PG_MAX_RANKS = 8;
constexpr int PG_THREADS = 256;

struct PeerGatherParams {
  uint8_t* stage[PG_MAX_RANKS];        // staging buffer of rank r (peer mapping; own for r == rank)
  uint32_t* flags[PG_MAX_RANKS];       // signal pad of rank r
  int R, rank;
  long long slot_bytes;                // capacity of one slot
  long long half_bytes;                // R * slot_bytes
  unsigned int* state;                 // local: [0] epoch, [1] CTAs done storing, [2] CTAs exited
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(PG_THREADS) peer_allgather_kernel(const uint4* __restrict__ src,
                                                                    uint4* __restrict__ dst, long long n16,
                                                                    const PeerGatherParams p) {
  __shared__ unsigned int s_epoch, s_last;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned int*>(p.state) + 1u;
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const long long half_off = (long long)(epoch & 1u) * p.half_bytes;
  const long long gstride = (long long)gridDim.x * PG_THREADS;
  const long long gtid = (long long)blockIdx.x * PG_THREADS + threadIdx.x;
  // 1. my message -> slot `rank` of every rank's staging half
  for (int r = 0; r < p.R; ++r) {
    const int peer = (p.rank + r) % p.R;             // start with myself, spread the link load
    uint4* d = reinterpret_cast<uint4*>(p.stage[peer] + half_off + (long long)p.rank * p.slot_bytes);
    for (long long i = gtid; i < n16; i += gstride) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. last CTA done storing publishes the epoch to everybody
  if (threadIdx.x == 0) s_last = (atomicAdd(p.state + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < p.R) st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
    if (threadIdx.x == 0) p.state[1] = 0u;
  }
  // 3. wait for all ranks, then hand the gathered half to the caller
  if (threadIdx.x < p.R) {
    const uint32_t* f = p.flags[p.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
  for (int r = 0; r < p.R; ++r) {
    const uint4* s = reinterpret_cast<const uint4*>(p.stage[p.rank] + half_off + (long long)r * p.slot_bytes);
    uint4* d = dst + (long long)r * n16;
    for (long long i = gtid; i < n16; i += gstride) d[i] = s[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(p.state + 2, 1u) == gridDim.x - 1) {   // last CTA out: this call is over
      p.state[2] = 0u;
      __threadfence();
      p.state[0] = epoch;
    }
  }
}

}  // namespace dl

extern "C" int dl_peer_allgather(const void* src, void* dst, long long nbytes, void* const* stage_ptrs,
                                 void* const* flag_ptrs, int nranks, int rank, long long slot_bytes,
                                 void* state, void* stream_) {
  using namespace dl;
  DL_CHECK_ARG(stage_ptrs && flag_ptrs && state && (nbytes == 0 || (src && dst)), "peer_allgather: null pointer");
  DL_CHECK_ARG(nranks >= 1 && nranks <= PG_MAX_RANKS && rank >= 0 && rank < nranks, "peer_allgather: bad ranks");
  DL_CHECK_ARG(nbytes >= 0 && nbytes % 16 == 0 && nbytes <= slot_bytes && slot_bytes % 16 == 0,
               "peer_allgather: message of %lld bytes (slot %lld) must be a multiple of 16 and fit a slot", nbytes,
               slot_bytes);
  PeerGatherParams p;
  for (int r = 0; r < nranks; ++r) {
    p.stage[r] = reinterpret_cast<uint8_t*>(stage_ptrs[r]);
    p.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r]);
  }
  p.R = nranks; p.rank = rank; p.slot_bytes = slot_bytes; p.half_bytes = slot_bytes * nranks;
  p.state = reinterpret_cast<unsigned int*>(state);
  const long long n16 = nbytes / 16;
  long long blocks = (n16 + 4 * PG_THREADS - 1) / (4 * PG_THREADS);
  if (blocks > 96) blocks = 96;                          // all co-resident (148 SMs): CTAs spin on the flags
  if (blocks < 1) blocks = 1;
  peer_allgather_kernel<<<(unsigned)blocks, PG_THREADS, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), n16, p);
  return check_launch("peer_allgather");
}
