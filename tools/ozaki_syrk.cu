// Prototype: the Schur SYRK  S = A A^T  (A = [Y; z], 11 C + 1 rows, K = 3 P columns, FP64) through an
// error-free integer split on the 5th-generation tensor cores (tcgen05.mma kind::i8, int32
// accumulators in TMEM) -- the "Ozaki scheme" route below the FP64 floor named in DESIGN.md.
//
//   a_rk = 2^e_r * sum_i d_i(r,k) 2^-8(i+1),  d_i int8 balanced digits, e_r from the row maximum
//   S_rc = 2^(e_r+e_c) * sum_{i+j <= DMAX} 2^-8(i+j+2) * (D_i D_j^T)_rc,   D_i D_j^T exact in int32
//
// Layout: the digit planes are stored in HBM exactly as the UMMA shared-memory operand layout
// (K-major, no swizzle: 8 rows x 16 bytes core matrices; [k-step of 32][slice][row group of 8]
// [k half][8][16]), so a CTA fetches its operands with plain 1-D bulk copies (cp.async.bulk +
// mbarrier complete_tx, no tensor map).  A CTA owns one output tile (128 rows x 8*n_nrg columns,
// all DMAX+1 anti-diagonals d = i+j side by side in TMEM: (DMAX+1) * 64 <= 512 columns) and a range of
// k-steps; for slice i of the rows ONE instruction multiplies against the slices j = 0..DMAX-i
// of the columns at once (they are contiguous in shared memory: N = 64 (DMAX-i+1) <= 256 per
// instruction), so the 128 x 32 byte A tile is read once per i instead of once per (i, j).
// Warp roles: 0..3 = epilogue (TMEM -> FP64 registers every FLUSH k-steps, before an int32
// accumulator can overflow), 4 = bulk-copy producer, 5 = MMA issuer (one thread; it must be the
// highest warp id of its scheduler and the waiting warps must sleep, or their polling starves it:
// the first version ran 10x slower for exactly that reason).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ozaki_syrk tools/ozaki_syrk.cu
// Run:    tools/ozaki_syrk [cameras=24] [points=1000000] [NS=6] [DMAX=6]
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int MAX_NS = 7;
constexpr int STAGES = 5;
constexpr int FLUSH = 448;               // k-steps between TMEM flushes: 7 pairs * 32 * 128^2 * 448 < 2^31
constexpr int OZ_THREADS = 192;

struct Work {                            // one CTA
  int m_rg0, m_nrg, n_rg0, n_nrg;        // row groups (8 rows) of the output tile
  int ks0, ks1;                          // k-steps [ks0, ks1)
  int transposed, pad;                   // tile holds S[c][r] (rows = columns of S)
};

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ int* g_err_flag;
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (BACKOFF) __nanosleep(256);    // waiting warps must not out-prioritise the MMA issuer (scheduler: highest warp id first)
    if (clock64() - t0 > 4000000000LL) {           // ~2 s: report instead of hanging the GPU
      printf("ozaki: mbarrier timeout tag %d block %d thread %d\n", tag, blockIdx.x, threadIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// timing experiments only (numerically meaningless on int8 digit planes)
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ bool elect_one() {      // one lane of the (converged) warp; lets ptxas issue UTC*MMA without a per-instruction election loop
  uint32_t p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_unused(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}

// ------------------------------------------------------------------------------ test data
__device__ __forceinline__ double hash_unit(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void k_fill(double* __restrict__ A, long long K, int R, int ldr) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * ldr) return;
  const long long k = idx / ldr;
  const int r = (int)(idx - k * ldr);
  double v = 0.0;
  if (r < R) {
    const double rs = pow(10.0, -1.0 + 4.0 * hash_unit(0x9e3779b97f4a7c15ull * (r + 1)));
    const double ps = 0.2 + 1.8 * hash_unit(0xbf58476d1ce4e5b9ull * (k / 3 + 1));
    v = rs * ps * (2.0 * hash_unit((unsigned long long)idx * 0x94d049bb133111ebull + 12345) - 1.0 + 0.3);
  }
  A[idx] = v;
}

// ------------------------------------------------------------------------------ slicing
// A is stored like the engine's Yg: [k][ldr] doubles (k = 3 p + kappa, rows contiguous).
__global__ void k_rowmax(const double* __restrict__ A, long long K, int R, int ldr, double* __restrict__ rmax_part) {
  // block b handles a K range; thread = row (R <= blockDim)
  const int r = threadIdx.x;
  const long long per = (K + gridDim.x - 1) / gridDim.x;
  const long long k0 = blockIdx.x * per, k1 = min(K, k0 + per);
  double m = 0.0;
  if (r < R) for (long long k = k0; k < k1; ++k) m = fmax(m, fabs(A[k * ldr + r]));
  if (r < R) rmax_part[(size_t)blockIdx.x * R + r] = m;
}
__global__ void k_rowexp(const double* __restrict__ rmax_part, int nb, int R, int* __restrict__ e_row) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double m = 0.0;
  for (int b = 0; b < nb; ++b) m = fmax(m, rmax_part[(size_t)b * R + r]);
  int e = 0;
  if (m > 0.0) { frexp(m, &e); e += 2; }       // m = f 2^e, f in [0.5, 1)  =>  |a| 2^-(e+2) < 0.25
  e_row[r] = e;
}
// one thread per (k, row): digits -> planes[((ks*NS + i)*NRG + rg)*256 + kh*128 + r8*16 + kb]
__global__ void k_slice(const double* __restrict__ A, long long K, int R, int ldr, int NRG, int NS,
                        const int* __restrict__ e_row, int8_t* __restrict__ planes) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int RP = NRG * 8;
  const long long Kp = (K + 31) / 32 * 32;
  if (idx >= Kp * RP) return;
  const long long k = idx / RP;
  const int r = (int)(idx - k * RP);
  long long T = 0;
  if (k < K && r < R) T = llrint(ldexp(A[k * ldr + r], 8 * NS - e_row[r]));
  const long long ks = k >> 5;
  const int kh = (int)((k >> 4) & 1), kb = (int)(k & 15), rg = r >> 3, r8 = r & 7;
  for (int i = NS - 1; i >= 0; --i) {          // least significant digit first
    const long long d = ((T + 128) & 255) - 128;
    T = (T - d) >> 8;
    planes[(((size_t)ks * NS + i) * NRG + rg) * 256 + kh * 128 + r8 * 16 + kb] = (int8_t)d;
  }
}

// ------------------------------------------------------------------------------ MMA issue plan
// The issuing thread is the bottleneck unless its instruction stream per MMA is a handful of
// instructions (first version: descriptor arithmetic and chunking loops at run time = ~480 cycles per
// MMA against 64-128 cycles of tensor work).  The (i, j-range, accumulate) list of one k-step is
// therefore a compile-time table and the issue loop is fully unrolled.
struct MmaOp { int i, j, nj, fresh; };
template <int NS, int DMAX, int NCOL, bool FIRST>
struct KstepPlan {
  MmaOp ops[24];
  int n;
  constexpr KstepPlan() : ops{}, n(0) {
    for (int i = 0; i < NS && i <= DMAX; ++i) {
      const int jmax = (NS - 1 < DMAX - i) ? NS - 1 : DMAX - i;
      int j = 0;
      while (j <= jmax) {
        // block d = i + j is initialised (accumulate = 0) by slice 0, except d = NS-1+i which
        // slice i is the first to touch (j = NS-1)
        const bool fresh = FIRST && (i == 0 || j == NS - 1);
        int j1 = j;
        while (j1 + 1 <= jmax && (j1 + 2 - j) * NCOL <= 256 && (FIRST && (i == 0 || j1 + 1 == NS - 1)) == fresh) ++j1;
        ops[n].i = i; ops[n].j = j; ops[n].nj = j1 - j + 1; ops[n].fresh = fresh ? 1 : 0;
        ++n;
        j = j1 + 1;
      }
    }
  }
};

template <int NS, int DMAX, int NCOL, bool FIRST>
__device__ __forceinline__ void issue_kstep(uint32_t tmem, uint32_t sA, uint32_t sB, uint32_t a_bytes, uint32_t b_bytes) {
  constexpr KstepPlan<NS, DMAX, NCOL, FIRST> plan{};
  constexpr uint32_t DESC_HI = (256u >> 4) | (1u << 14);              // SBO = 256 B, descriptor version 1
  constexpr uint32_t DESC_LO = (128u >> 4) << 16;                     // LBO = 128 B
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);   // s32 += s8 x s8, M = 128
#pragma unroll
  for (int q = 0; q < plan.n; ++q) {
    const uint64_t adesc = ((uint64_t)DESC_HI << 32) | (DESC_LO | (((sA + plan.ops[q].i * a_bytes) >> 4) & 0x3FFF));
    const uint64_t bdesc = ((uint64_t)DESC_HI << 32) | (DESC_LO | (((sB + plan.ops[q].j * b_bytes) >> 4) & 0x3FFF));
    mma_i8(tmem + (uint32_t)((plan.ops[q].i + plan.ops[q].j) * NCOL), adesc, bdesc,
           IDESC | ((uint32_t)(plan.ops[q].nj * NCOL >> 3) << 17), plan.ops[q].fresh ? 0u : 1u);
  }
}

template <int NS, int DMAX>
__device__ __forceinline__ void issue_kstep_any(int ncol, bool first, uint32_t tmem, uint32_t sA, uint32_t sB,
                                                uint32_t a_bytes, uint32_t b_bytes) {
  if (ncol == 64) { if (first) issue_kstep<NS, DMAX, 64, true>(tmem, sA, sB, a_bytes, b_bytes); else issue_kstep<NS, DMAX, 64, false>(tmem, sA, sB, a_bytes, b_bytes); }
  else if (ncol == 32) { if (first) issue_kstep<NS, DMAX, 32, true>(tmem, sA, sB, a_bytes, b_bytes); else issue_kstep<NS, DMAX, 32, false>(tmem, sA, sB, a_bytes, b_bytes); }
  else { if (first) issue_kstep<NS, DMAX, 16, true>(tmem, sA, sB, a_bytes, b_bytes); else issue_kstep<NS, DMAX, 16, false>(tmem, sA, sB, a_bytes, b_bytes); }
}

// ------------------------------------------------------------------------------ the SYRK kernel
__global__ void __launch_bounds__(OZ_THREADS, 1)
k_ozaki_syrk(const int8_t* __restrict__ planes, int NRG, int NS, int DMAX, const Work* __restrict__ work,
             double* __restrict__ partial /* [cta][128][64] */, int mode, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar[2 * STAGES + 2];
  __shared__ uint32_t s_tmem;
  const Work W = work[blockIdx.x];
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = (uint32_t)W.m_nrg * 256, b_bytes = (uint32_t)W.n_nrg * 256;   // per slice
  const uint32_t stage_bytes = (uint32_t)NS * (16 * 256 + 8 * 256);                     // fixed stride
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[STAGES]);
  const uint32_t bar_tfull = smem_u32(&s_bar[2 * STAGES]), bar_tempty = smem_u32(&s_bar[2 * STAGES + 1]);
  const int ND = DMAX + 1;
  const int ncol = W.n_nrg * 8;                 // columns per anti-diagonal block
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const int nks = W.ks1 - W.ks0;

  if (wid == 4) {
    // ------------------------------------------------ producer: bulk copies into the ring
    // one bulk copy per lane (2 NS copies per k-step): a single thread issuing them one after the
    // other was the bottleneck of the second version (~110 cycles per UBLKCP: 1340 cycles per k-step)
    if (!(mode & 64)) {
      const int ci = lane >> 1, cb = lane & 1;              // slice, operand (0 = rows/A, 1 = columns/B)
      for (int it = 0; it < nks; ++it) {
        const int st = it % STAGES;
        if (it >= STAGES) mbar_wait<true>(bar_empty + 8 * st, ((it / STAGES) - 1) & 1, 1);
        if (lane == 0) mbar_expect_tx(bar_full + 8 * st, (uint32_t)NS * (a_bytes + b_bytes));
        __syncwarp();
        if (lane < 2 * NS) {
          const int8_t* src = planes + (size_t)(W.ks0 + it) * NS * NRG * 256;
          const uint32_t dstA = smem0 + st * stage_bytes, dstB = dstA + (uint32_t)NS * 16 * 256;
          if (cb == 0) bulk_g2s(dstA + ci * a_bytes, src + ((size_t)ci * NRG + W.m_rg0) * 256, a_bytes, bar_full + 8 * st);
          else         bulk_g2s(dstB + ci * b_bytes, src + ((size_t)ci * NRG + W.n_rg0) * 256, b_bytes, bar_full + 8 * st);
        }
      }
    }
  } else if (wid == 5) {
    // ------------------------------------------------ MMA issuer (highest warp id of its scheduler)
    if (elect_one()) {
      const uint32_t idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
      int since_flush = 0, nflush = 0;
      const long long tm0 = clock64();
      for (int it = 0; it < nks; ++it) {
        const int st = it % STAGES;
        if (since_flush == 0 && nflush > 0) {                 // accumulators were handed to the epilogue
          mbar_wait(bar_tempty, (nflush - 1) & 1, 2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (!(mode & 64)) {
        mbar_wait(bar_full + 8 * st, (it / STAGES) & 1, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t sA = smem0 + st * stage_bytes, sB = sA + (uint32_t)NS * 16 * 256;
        const bool first = since_flush == 0;
        if (NS == 6 && DMAX == 6) issue_kstep_any<6, 6>(ncol, first, tmem, sA, sB, a_bytes, b_bytes);
        else if (NS == 6 && DMAX == 5) issue_kstep_any<6, 5>(ncol, first, tmem, sA, sB, a_bytes, b_bytes);
        else if (NS == 7 && DMAX == 6) issue_kstep_any<7, 6>(ncol, first, tmem, sA, sB, a_bytes, b_bytes);
        else if (NS == 5 && DMAX == 5) issue_kstep_any<5, 5>(ncol, first, tmem, sA, sB, a_bytes, b_bytes);
        else { if (it == 0) printf("ozaki: unsupported NS/DMAX\n"); }
        if (!(mode & 64)) mma_commit(bar_empty + 8 * st);     // frees the ring slot when these MMAs retire
        ++since_flush;
        if (since_flush == FLUSH || it == nks - 1) {
          mma_commit(bar_tfull);
          since_flush = 0;
          ++nflush;
        }
      }
      dbg[2 * blockIdx.x] = clock64() - tm0; dbg[2 * blockIdx.x + 1] = nks;
    }
  } else {
    // ------------------------------------------------ epilogue: TMEM -> FP64 registers
    const int q = wid & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;               // tile row
    double acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = 0.0;
    const int nfl = (nks + FLUSH - 1) / FLUSH;
    for (int f = 0; f < nfl; ++f) {
      mbar_wait<true>(bar_tfull, f & 1, 4);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int d = 0; d < ND; ++d) {
        const double sc = ldexp(1.0, -8 * (d + 2));
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * ncol);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (16 * h < ncol) {                   // warp-uniform
            uint32_t v[16];
            tmem_ld16(taddr + 16 * h, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[16 * h + c] = fma((double)(int)v[c], sc, acc[16 * h + c]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_tempty);
    }
    double* out = partial + ((size_t)blockIdx.x * 128 + row) * 64;
#pragma unroll
    for (int c = 0; c < 64; ++c) out[c] = acc[c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// S[r][c] (lower triangle, mirrored) = 2^(e_r + e_c) * sum of the CTA partials of the tile that owns (r, c)
__global__ void k_gather(const double* __restrict__ partial, const Work* __restrict__ work, int nwork, int R,
                         const int* __restrict__ e_row, double* __restrict__ S) {
  const int w = blockIdx.x;
  const Work W = work[w];
  for (int idx = threadIdx.x; idx < W.m_nrg * 8 * W.n_nrg * 8; idx += blockDim.x) {
    const int tr = idx / (W.n_nrg * 8), tc = idx % (W.n_nrg * 8);
    int r = W.m_rg0 * 8 + tr, c = W.n_rg0 * 8 + tc;
    if (r >= R || c >= R) continue;
    if (W.transposed) { const int t = r; r = c; c = t; }
    if (c > r) continue;
    const double v = ldexp(partial[((size_t)w * 128 + tr) * 64 + tc], e_row[r] + e_row[c]);
    atomicAdd(&S[(size_t)r * R + c], v);
    if (r != c) atomicAdd(&S[(size_t)c * R + r], v);
  }
}

// FP64 reference on the GPU: one thread per lower-triangle entry (slow but simple)
__global__ void k_syrk_ref(const double* __restrict__ A, long long K, int R, int ldr, double* __restrict__ S) {
  const int r = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > r || r >= R) return;
  double s = 0.0;
  for (long long k = 0; k < K; ++k) s = fma(A[k * ldr + r], A[k * ldr + c], s);
  S[(size_t)r * R + c] = s;
  S[(size_t)c * R + r] = s;
}

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 24;
  const long long P = argc > 2 ? atoll(argv[2]) : 1000000;
  const int NS = argc > 3 ? atoi(argv[3]) : 6;
  const int DMAX = argc > 4 ? atoi(argv[4]) : 6;
  const int mode = argc > 5 ? atoi(argv[5]) : 0;
  const int R = 11 * C + 1;
  const long long K = 3 * P;
  int NRG = (R + 7) / 8;
  NRG += NRG & 1;                                   // N must be a multiple of 16
  const int ldr = NRG * 8;
  if (NS > MAX_NS || (DMAX + 1) * 64 > 512 || DMAX > 2 * (NS - 1)) { fprintf(stderr, "bad NS/DMAX\n"); return 1; }
  int sm_count = 148;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  sm_count = prop.multiProcessorCount;
  printf("ozaki_syrk: %s, %d SMs; C=%d R=%d (NRG=%d) P=%lld K=%lld NS=%d DMAX=%d\n", prop.name, sm_count, C, R, NRG, P, K, NS, DMAX);

  // ---- synthetic A with the statistics of Y: per-row scales over 4 decades, per-point spread, a bias
  double* dA;
  CK(cudaMalloc(&dA, (size_t)K * ldr * 8));
  {
    const long long tot = K * ldr;
    k_fill<<<(unsigned)((tot + 255) / 256), 256>>>(dA, K, R, ldr);
    CK(cudaDeviceSynchronize());
  }

  // ---- plan: tiles of the lower triangle and their k-ranges
  const long long nks_total = (K + 31) / 32;
  struct Tile { int m0, mn, n0, nn, tr; };
  std::vector<Tile> tiles;
  const int nmt = (NRG + 15) / 16;
  const int last_m = NRG - 16 * (nmt - 1);
  const bool fold_last = nmt > 1 && last_m <= 4;      // few rows left: handle them as COLUMNS (transposed)
  const int nfull = fold_last ? nmt - 1 : nmt;
  for (int mt = 0; mt < nfull; ++mt) {
    const int m0 = 16 * mt, mn = std::min(16, NRG - m0);
    for (int n0 = 0; n0 < m0 + mn; n0 += 8) tiles.push_back({m0, mn, n0, std::min(8, m0 + mn - n0), 0});
  }
  if (fold_last) {
    const int c0 = 16 * (nmt - 1);
    for (int mt = 0; mt < nfull; ++mt) tiles.push_back({16 * mt, 16, c0, last_m, 1});
    tiles.push_back({c0, last_m, c0, last_m, 0});
  }
  const double narrow_cost = getenv("OZ_NARROW_COST") ? atof(getenv("OZ_NARROW_COST")) : 0.45;
  auto cost_of = [&](const Tile& t) { return t.nn >= 8 ? 1.0 : std::max(narrow_cost, t.nn / 8.0); };
  double units = 0;
  for (auto& t : tiles) units += cost_of(t);
  std::vector<Work> work;
  for (auto& t : tiles) {
    int n = std::max(1, (int)floor(sm_count * cost_of(t) / units));
    n = (int)std::min<long long>(n, nks_total);
    for (int q = 0; q < n; ++q)
      work.push_back({t.m0, t.mn, t.n0, t.nn, (int)(nks_total * q / n), (int)(nks_total * (q + 1) / n), t.tr, 0});
  }
  printf("plan: %zu tiles, %zu CTAs, %lld k-steps\n", tiles.size(), work.size(), nks_total);
  for (auto& t : tiles) printf("  tile rows rg %d+%d x cols rg %d+%d%s\n", t.m0, t.mn, t.n0, t.nn, t.tr ? " (transposed)" : "");

  // ---- slice
  const int NB = 512;
  double* d_rpart; int* d_erow; int8_t* d_planes; Work* d_work; double *d_partial, *d_S, *d_Sref;
  CK(cudaMalloc(&d_rpart, (size_t)NB * R * 8));
  CK(cudaMalloc(&d_erow, R * sizeof(int)));
  const size_t plane_bytes = (size_t)nks_total * NS * NRG * 256;
  CK(cudaMalloc(&d_planes, plane_bytes));
  CK(cudaMalloc(&d_work, work.size() * sizeof(Work)));
  CK(cudaMemcpy(d_work, work.data(), work.size() * sizeof(Work), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_partial, work.size() * 128 * 64 * 8));
  CK(cudaMalloc(&d_S, (size_t)R * R * 8));
  CK(cudaMalloc(&d_Sref, (size_t)R * R * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms_slice = 0, ms_syrk = 0;
  long long* d_dbg;
  CK(cudaMalloc(&d_dbg, 16 * 1024));
  CK(cudaMemset(d_dbg, 0, 16 * 1024));
  const size_t smem_bytes = (size_t)STAGES * NS * (16 * 256 + 8 * 256) + 1024;
  CK(cudaFuncSetAttribute(k_ozaki_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int nrep = getenv("OZ_REPS") ? atoi(getenv("OZ_REPS")) : 3;
  for (int rep = 0; rep < nrep; ++rep) {
    CK(cudaEventRecord(e0));
    k_rowmax<<<NB, ((R + 31) / 32) * 32>>>(dA, K, R, ldr, d_rpart);
    k_rowexp<<<(R + 127) / 128, 128>>>(d_rpart, NB, R, d_erow);
    const long long tot = nks_total * 32 * NRG * 8;
    k_slice<<<(unsigned)((tot + 255) / 256), 256>>>(dA, K, R, ldr, NRG, NS, d_erow, d_planes);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_slice, e0, e1));
    CK(cudaMemset(d_partial, 0, work.size() * 128 * 64 * 8));
    CK(cudaEventRecord(e0));
    k_ozaki_syrk<<<(unsigned)work.size(), OZ_THREADS, smem_bytes>>>(d_planes, NRG, NS, DMAX, d_work, d_partial, mode, d_dbg);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms_syrk, e0, e1));
    std::vector<long long> hd(2 * work.size());
    CK(cudaMemcpy(hd.data(), d_dbg, hd.size() * 8, cudaMemcpyDeviceToHost));
    if (rep < 3 || rep == nrep - 1) {
      printf("rep %d: slice %.3f ms (naive kernel), int8 SYRK %.3f ms\n", rep, ms_slice, ms_syrk);
      if (rep == nrep - 1) {
        size_t w = 0;
        for (auto& t : tiles) {      // first CTA of every tile
          printf("   tile m%d+%d n%d+%d: %lld cycles / %lld k-steps = %.0f per k-step\n", t.m0, t.mn, t.n0, t.nn, hd[2 * w], hd[2 * w + 1],
                 (double)hd[2 * w] / (double)std::max<long long>(1, hd[2 * w + 1]));
          while (w < work.size() && work[w].m_rg0 == t.m0 && work[w].n_rg0 == t.n0) ++w;
        }
      }
    }
  }
  CK(cudaMemset(d_S, 0, (size_t)R * R * 8));
  k_gather<<<(unsigned)work.size(), 256>>>(d_partial, d_work, (int)work.size(), R, d_erow, d_S);
  CK(cudaEventRecord(e0));
  k_syrk_ref<<<dim3((R + 63) / 64, R), 64>>>(dA, K, R, ldr, d_Sref);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms_ref = 0;
  CK(cudaEventElapsedTime(&ms_ref, e0, e1));
  std::vector<double> S((size_t)R * R), Sr((size_t)R * R);
  CK(cudaMemcpy(S.data(), d_S, S.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Sr.data(), d_Sref, Sr.size() * 8, cudaMemcpyDeviceToHost));
  double smax = 0, emax = 0, escaled = 0;
  for (size_t i = 0; i < S.size(); ++i) smax = std::max(smax, fabs(Sr[i]));
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < R; ++c) {
      const double e = fabs(S[(size_t)r * R + c] - Sr[(size_t)r * R + c]);
      emax = std::max(emax, e);
      escaled = std::max(escaled, e / sqrt(Sr[(size_t)r * R + r] * Sr[(size_t)c * R + c]));
    }
  const double flop = (double)K * R * (R + 1);        // 2 * K * R(R+1)/2
  printf("accuracy vs FP64 FMA reference (%.1f ms): max|dS|/max|S| = %.3e, Jacobi-scaled max = %.3e\n", ms_ref, emax / smax, escaled);
  printf("int8 SYRK: %.3f ms = %.1f FP64-equivalent TFLOP/s (lower triangle %.3e flop); %d digit products -> %.1f int8 TOP/s issued\n",
         ms_syrk, flop / ms_syrk * 1e-9, flop, (DMAX + 1) * (DMAX + 2) / 2 - std::max(0, DMAX - NS + 1) * (std::max(0, DMAX - NS + 1) + 1),
         0.0);
  return 0;
}
