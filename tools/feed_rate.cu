// Microbenchmark: how fast can one SM pull operand bytes into shared memory with cp.async.bulk while all
// 148 SMs do the same?  (Is the 31..33 B/clk/SM that k_i8_syrk reaches a limit of the copy path or of the
// copy path competing with the tensor core's operand reads?)
// One CTA per SM; a 3-stage mbarrier ring; the producer warp issues `ncopy` bulk copies of `cbytes` per stage
// (one per lane), a consumer thread releases the stage as soon as it is full.  Optional: `mma` > 0 issues that
// many tcgen05.mma kind::i8 (M = 128, N = 48, K = 32) per stage on the stage's bytes from a third warp, so
// that the tensor core's shared-memory reads run against the copies like in the product kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/feed_rate tools/feed_rate.cu
//   tools/feed_rate <total MB> <cta stride KB (0 = all CTAs read the same bytes)> <ncopy> <cbytes> <iters> <mma per stage> [N] [grid] [arun = consecutive MMAs on one A] [coll = 1: collector::a fill/use/lastuse] [streams: CTA b reads stream b % streams]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 72 * 1024;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// COLL: 0 no qualifier (discard), 1 collector::a::fill, 2 ::use, 3 ::lastuse
template <int COLL>
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
#define MMA_ASM(Q) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8" Q " [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" \
                                ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u), "r"(0u) : "memory")
  if (COLL == 0) MMA_ASM("");
  else if (COLL == 1) MMA_ASM(".collector::a::fill");
  else if (COLL == 2) MMA_ASM(".collector::a::use");
  else MMA_ASM(".collector::a::lastuse");
#undef MMA_ASM
}
// one stage's MMAs, fully unrolled: every descriptor offset, TMEM column and qualifier is a compile-time constant
template <int NMMA, int ARUN, int COLL>
__device__ __forceinline__ void issue_stage(uint64_t ad0, uint64_t bd0, uint32_t tmem, uint32_t idesc) {
#pragma unroll
  for (int m = 0; m < NMMA; ++m) {
    const int run = m / ARUN, pos = m % ARUN;
    const uint64_t ad = ad0 + (uint64_t)(((run % 6) * 8192) >> 4);
    const uint64_t bd = bd0 + (uint64_t)(((m % 6) * 3072) >> 4);
    const uint32_t d = tmem + (uint32_t)(NMMA == 10 ? (m % 2) * 256 : (m % 7) * 64);
    const int mode = (!COLL || ARUN == 1) ? 0 : (pos == 0 ? 1 : ((pos == ARUN - 1 || m == NMMA - 1) ? 3 : 2));
    if (mode == 0) mma_i8<0>(d, ad, bd, idesc);
    else if (mode == 1) mma_i8<1>(d, ad, bd, idesc);
    else if (mode == 2) mma_i8<2>(d, ad, bd, idesc);
    else mma_i8<3>(d, ad, bd, idesc);
  }
}
__global__ void __launch_bounds__(96, 1)
k_feed(const uint8_t* __restrict__ src, long long total, long long cta_stride, int ncopy, int cbytes, int iters,
       int nmma, int N, int arun, int coll, int streams, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], mdone[STAGES];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1u) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mdone[s])), "r"(1u) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long t0 = clock64();
  if (wid == 0) {
    // streams > 0: CTA b reads stream b % streams (CTAs b, b + streams, ... read the SAME bytes, like the tiles of one
    // K range of k_i8_syrk)
    long long off = (long long)(streams > 0 ? blockIdx.x % streams : blockIdx.x) * cta_stride;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(smem_u32(&empty[s]), ((it / STAGES) - 1) & 1);
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)(ncopy * cbytes)) : "memory");
      __syncwarp();
      for (int c = lane; c < ncopy; c += 32) {
        long long a = (off + (long long)c * cbytes) % total;
        if (a + cbytes > total) a = 0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + (size_t)s * STAGE_BYTES + (size_t)c * cbytes)), "l"(src + a), "r"((uint32_t)cbytes),
                       "r"(smem_u32(&full[s])) : "memory");
      }
      off += (long long)ncopy * cbytes;
    }
  } else if (wid == 1) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        mbar_wait(smem_u32(&full[s]), (it / STAGES) & 1);
        if (nmma == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      }
    }
  } else if (nmma > 0) {
    // MMA issuer: nmma instructions per stage on the stage's bytes, then commit -> empty[s]
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    if (elected) {
      const uint32_t tmem = s_tmem;
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        mbar_wait(smem_u32(&full[s]), (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t ad0 = make_desc(base, 128, 256), bd0 = make_desc(base + 49152, 128, 256);
        if (nmma == 52 && arun == 1) issue_stage<52, 1, 0>(ad0, bd0, tmem, idesc);
        else if (nmma == 52 && arun == 4 && !coll) issue_stage<52, 4, 0>(ad0, bd0, tmem, idesc);
        else if (nmma == 52 && arun == 4 && coll) issue_stage<52, 4, 1>(ad0, bd0, tmem, idesc);
        else if (nmma == 26 && arun == 1) issue_stage<26, 1, 0>(ad0, bd0, tmem, idesc);
        else if (nmma == 26 && arun == 4 && coll) issue_stage<26, 4, 1>(ad0, bd0, tmem, idesc);
        else if (nmma == 48 && arun == 6 && coll) issue_stage<48, 6, 1>(ad0, bd0, tmem, idesc);
        else if (nmma == 48 && arun == 6 && !coll) issue_stage<48, 6, 0>(ad0, bd0, tmem, idesc);
        else if (nmma == 10 && arun == 1) issue_stage<10, 1, 0>(ad0, bd0, tmem, idesc);
        else if (nmma == 10 && arun == 2 && coll) issue_stage<10, 2, 1>(ad0, bd0, tmem, idesc);
        else if (nmma == 10 && arun == 2 && !coll) issue_stage<10, 2, 0>(ad0, bd0, tmem, idesc);
        else if (nmma == 10 && arun == 5 && coll) issue_stage<10, 5, 1>(ad0, bd0, tmem, idesc);
        else issue_stage<8, 1, 0>(ad0, bd0, tmem, idesc);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mdone[0])) : "memory");
      mbar_wait(smem_u32(&mdone[0]), 0);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
  const long long total = (argc > 1 ? atoll(argv[1]) : 4096) << 20;
  const long long stride = (argc > 2 ? atoll(argv[2]) : 0) << 10;
  const int ncopy = argc > 3 ? atoi(argv[3]) : 12;
  const int cbytes = argc > 4 ? atoi(argv[4]) : 6144;
  const int iters = argc > 5 ? atoi(argv[5]) : 2000;
  const int nmma = argc > 6 ? atoi(argv[6]) : 0;
  const int N = argc > 7 ? atoi(argv[7]) : 48;
  const int grid = argc > 8 ? atoi(argv[8]) : 148;
  const int arun = argc > 9 ? atoi(argv[9]) : 1;
  const int coll = argc > 10 ? atoi(argv[10]) : 0;
  const int streams = argc > 11 ? atoi(argv[11]) : 0;
  if (ncopy * cbytes > STAGE_BYTES || cbytes % 16) { printf("stage too large\n"); return 1; }
  uint8_t* src;
  long long* out;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  CK(cudaMalloc(&out, 1024 * sizeof(long long)));
  CK(cudaFuncSetAttribute(k_feed, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    k_feed<<<grid, 96, STAGES * STAGE_BYTES>>>(src, total, stride, ncopy, cbytes, iters, nmma, N, arun, coll, streams, out);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    const long long mx = *std::max_element(h.begin(), h.end()), mn = *std::min_element(h.begin(), h.end());
    const double bytes = (double)ncopy * cbytes * iters;
    if (rep == 2)
      printf("total %lld MB stride %lld KB copies %d x %d B iters %d mma %d N %d grid %d arun %d coll %d streams %d: %.3f ms  %.1f GB/s chip  cycles/stage max %.0f min %.0f  "
             "B/clk/SM %.1f\n", total >> 20, stride >> 10, ncopy, cbytes, iters, nmma, N, grid, arun, coll, streams, ms, bytes * grid / ms * 1e-6,
             (double)mx / iters, (double)mn / iters, bytes / (double)mx);
  }
  return 0;
}
