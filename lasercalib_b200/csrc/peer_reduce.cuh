// All-reduce of the small per-iteration payloads (camera sums, Gram sums, the reduced camera system: 8 B .. 2 MB)
// over NVLink peer memory instead of ncclAllReduce: one kernel, one-shot ("every rank reads every rank").
//
// Every rank owns one cudaMalloc'ed block [2 slots x PEER_SLOT doubles | 2 slots x ranks x PEER_CTAS flags],
// exported with cudaIpcGetMemHandle and mapped by all peers (handles exchanged with ncclAllGather, lcba.cu).
// All-reduce number `epoch` (counted identically on all ranks, like the call order NCCL requires):
//   CTA c owns elements [count c / G, count (c+1) / G):
//   1. copies its part of the input into the own slot (epoch & 1),
//   2. releases flag[slot][own rank][c] = epoch in EVERY peer's flag array (st.release.sys over NVLink),
//   3. waits until its own flags [slot][r][c] of all ranks r carry the epoch (ld.acquire.sys, local polls),
//   4. sums the part over r = 0 .. n-1 IN RANK ORDER, reading the peers' slots through NVLink, and writes the
//      result in place: every rank adds the same numbers in the same order -> bit-identical results on all
//      ranks (the ranks take the same accept / reject decisions from them).
// Two slots are enough: a rank can enter all-reduce e + 2 (same slot as e) only after every peer has released
// e + 1, which a peer does after it has finished reading e.
// A wait is bounded (~2 s): a lost peer ends in an error code (fail word in mapped host memory), not in a hung GPU.
#pragma once
#include "common.cuh"

namespace lcba {

constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_CTAS = 16;
constexpr size_t PEER_SLOT = (size_t)1 << 19;          // doubles per slot: 4 MB (64 cameras: 2.0 MB of Sred)
constexpr int PEER_THREADS = 256;
constexpr int PEER_OP_SUM = 0, PEER_OP_MAX = 1;

struct PeerPtrs {
  double* sym[PEER_MAX_RANKS];
  unsigned* flags[PEER_MAX_RANKS];
};
inline size_t peer_block_bytes() {
  return 2 * PEER_SLOT * sizeof(double) + (size_t)2 * PEER_MAX_RANKS * PEER_CTAS * sizeof(unsigned);
}
inline int peer_grid(size_t count) {
  return (int)std::max<size_t>(1, std::min<size_t>(PEER_CTAS, (count + 2047) / 2048));
}

__device__ __forceinline__ void peer_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned peer_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double peer_ld(const double* p) {     // never from a stale L1 line
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(PEER_THREADS)
k_peer_allreduce(PeerPtrs pp, int rank, int n, unsigned epoch, double* __restrict__ data, size_t count, int op,
                 volatile int* fail /* mapped host memory */) {
  const int c = blockIdx.x, G = gridDim.x, t = threadIdx.x;
  const size_t i0 = count * c / G, i1 = count * (c + 1) / G;
  const unsigned slot = epoch & 1u;
  double* mine = pp.sym[rank] + slot * PEER_SLOT;
  for (size_t i = i0 + t; i < i1; i += PEER_THREADS) mine[i] = data[i];
  __threadfence_system();
  __syncthreads();
  const size_t fidx = ((size_t)slot * PEER_MAX_RANKS) * PEER_CTAS + c;
  if (t < n) peer_st_release_sys(pp.flags[t] + fidx + (size_t)rank * PEER_CTAS, epoch);
  if (t < n) {
    const unsigned* f = pp.flags[rank] + fidx + (size_t)t * PEER_CTAS;
    const long long t0 = clock64();
    unsigned spins = 0;
    while ((int)(peer_ld_acquire_sys(f) - epoch) < 0) {
      if ((++spins & 0xfffu) == 0 && *fail) break;           // another wait already gave up: do not add 2 s each
      if (clock64() - t0 > 4000000000LL) { *fail = 1; break; }
    }
  }
  __syncthreads();
  for (size_t i = i0 + t; i < i1; i += PEER_THREADS) {
    double s = peer_ld(pp.sym[0] + slot * PEER_SLOT + i);
    for (int r = 1; r < n; ++r) {
      const double v = peer_ld(pp.sym[r] + slot * PEER_SLOT + i);
      s = (op == PEER_OP_SUM) ? s + v : fmax(s, v);
    }
    data[i] = s;
  }
}

}  // namespace lcba
