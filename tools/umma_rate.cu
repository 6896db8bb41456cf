// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, M = 128) for the operand layouts and
// shapes the Ozaki SYRK prototype uses.  One CTA per SM, one thread issues `iters` MMAs back to
// back (operands = whatever is in shared memory), commits, waits; prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// variant: 0 SW_NONE i8, 1 SW_NONE bf16, 2 SW128 i8, 3 SW128 bf16 ; dstep: column step of D between MMAs
__global__ void __launch_bounds__(128, 1) k_rate(int variant, int N, int iters, int dstep, int bstep, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x01010101u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const bool sw = variant >= 2;
    const bool i8 = (variant & 1) == 0;
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
    const uint32_t idesc = (i8 ? ((2u << 4) | (1u << 7) | (1u << 10)) : ((1u << 4) | (1u << 7) | (1u << 10))) |
                           ((128u >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t boff = (uint32_t)((it % 4) * bstep);
      const uint64_t ad = sw ? make_desc(a0, 16, 1024, 2) : make_desc(a0, 128, 256, 0);
      const uint64_t bd = sw ? make_desc(b0 + boff, 16, 1024, 2) : make_desc(b0 + boff, 128, 256, 0);
      const uint32_t dcol = dstep ? (uint32_t)((it % ((512 - N) / 64 + 1)) * 64) : 0u;   // 64-column aligned
      if (i8) mma<0>(tmem + dcol, ad, bd, idesc, it > 0); else mma<1>(tmem + dcol, ad, bd, idesc, it > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
// The Ozaki prototype's k-step: for slice i one or two MMAs against slices j = 0..DMAX-i (N = 64 per j),
// D block d = i + j; optional commit per k-step onto an mbarrier ring that nobody waits on.
__global__ void __launch_bounds__(128, 1) k_kstep(int NS, int DMAX, int ksteps, int commit_each, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x01010101u;
  if (tid == 0) {
    for (int b = 0; b < 8; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t base = smem_u32(smem);
    const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < ksteps; ++it) {
      const uint32_t sA = base + (uint32_t)(it % 4) * 36864u, sB = sA + 24576u;
      for (int i = 0; i < NS && i <= DMAX; ++i) {
        const uint64_t ad = make_desc(sA + i * 4096, 128, 256, 0);
        const int jmax = min(NS - 1, DMAX - i);
        for (int j = 0; j <= jmax; j += 4) {
          const int nj = min(4, jmax - j + 1);
          mma<0>(tmem + (uint32_t)((i + j) * 64), ad, make_desc(sB + j * 2048, 128, 256, 0),
                 idesc0 | ((uint32_t)(nj * 64 >> 3) << 17), (it > 0 || i > 0) ? 1u : 0u);
        }
      }
      if (commit_each)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1 + it % 4])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024);
  cudaFuncSetAttribute(k_kstep, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024);
  const int iters = 2000;
  for (int ce = 0; ce < 2; ++ce)
    for (int grid : {1, 148}) {
      k_kstep<<<grid, 128, 161 * 1024>>>(6, 6, 500, ce, d);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("k-step sequence NS=6 DMAX=6 commit_each=%d grid=%3d: issue %.0f cyc/k-step, complete %.0f cyc/k-step (ideal 832)\n", ce, grid, (double)h[0] / 500, (double)h[1] / 500);
    }
  const char* names[] = {"SW_NONE i8", "SW_NONE bf16", "SW128 i8", "SW128 bf16"};
  for (int variant = 0; variant < 4; ++variant)
    for (int N : {256})
      for (int dstep : {0})
        for (int grid : {148}) {
          k_rate<<<grid, 128, 161 * 1024>>>(variant, N, iters, dstep, 2048, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%-12s N=%3d dstep=%2d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)%s\n", names[variant], N, dstep, grid,
                 (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
