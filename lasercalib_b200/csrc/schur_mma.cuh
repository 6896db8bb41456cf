// Reduced camera system on the FP64 tensor path (DMMA, mma.sync m8n8k4 f64) for dense rigs.
//
// S - U = - sum_p Y_p Y_p^T is a symmetric rank-3P update: a SYRK with K = 3 P.  The DFMA
// kernel (schur.cuh) is bound by the shared-memory crossbar (12 operand doubles per 36 DFMA
// and lane) and by FP64 issue slots; with DMMA a lane loads ONE double per 8 x 8 x 4 fragment
// and the operand broadcast happens inside the tensor datapath: 12 doubles per 36 DMMA =
// 288 FMA per lane (8x less shared-memory traffic, 8x fewer issue slots).  Measured on B200
// (tools/dmma_syrk.cu, fragment loads included): 33.6 TFLOP/s with one warp per SM
// sub-partition, 36.9 with two.
//
// Decomposition.  Cameras are grouped in quads (4 x 12 padded rows = 48 = 6 tiles of 8); a
// consumer WARP owns one 48 x 48 region (quad_i x quad_j, j <= i): 36 tiles, 72 FP64
// accumulators per lane.  A CTA ("kind") holds up to 8 regions = 8 consumer warps + 4 producer
// warps (register budget moved from the producers to the consumers with setmaxnreg: two
// consumers and one producer per sub-partition), and streams a slice of the points:
//   producer: (point, camera slot) -> Y = (Jc^T Jp) L^-T (12 rows: 9 + cx, cy + the z row that
//             yields the reduced right-hand side, see schur.cuh) into a 2-stage ring laid out
//             [kappa = 3 q + k][row = 12 slot + r], row stride = 4 (mod 16) doubles, so that the
//             fragment loads (lane -> row lane/4, kappa lane%4) are conflict free;
//   consumer: per K-step (4 kappa = 4/3 point) 12 LDS.64 + 36 DMMA.
// Invisible (point, camera) pairs and the tail of the last chunk are exact zeros.  The
// camera blocks U_j = sum Jc^T Jc are accumulated by k_cam_normal (one thread per (point,
// camera), fixed camera per thread, 66 register accumulators) and added to the diagonal pair
// blocks before the all-reduce, so everything downstream is unchanged.
#pragma once
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "schur.cuh"

namespace lcba {

constexpr int MMA_REGIONS = 8;        // consumer warps per CTA (two warpgroups)
constexpr int MMA_PROD_WARPS = 4;     // producer warps (one warpgroup)
constexpr int MMA_THREADS = 32 * (MMA_REGIONS + MMA_PROD_WARPS);
// 12 warps are launched with 168 registers each; the producer warpgroup gives registers back
// (setmaxnreg.dec) and the two consumer warpgroups take them (setmaxnreg.inc): every SM
// sub-partition then holds 2 consumer warps x 208 + 1 producer warp x 88 registers.
#define MMA_CONS_REGS "208"
#define MMA_PROD_REGS "88"

struct MmaRegion {
  int8_t srow[4], scol[4];   // shared-memory slots of the row / column cameras (0 when absent)
  int8_t crow[4], ccol[4];   // camera ids or -1
};

struct MmaKind {
  int nslots, nreg, reg_base, rp;   // rp: row stride of the ring in doubles (= 4 mod 16)
  int sp, pad0, pad1, pad2;         // points per stage (multiple of 4)
  uint8_t slot_cam[LCBA_MAX_CAMERAS];
};

struct MmaPlan {
  int C = 0, nkinds = 0, nslices = 0;
  size_t smem_bytes = 0;
  std::vector<MmaKind> kinds;
  std::vector<MmaRegion> regions;
};

inline MmaPlan make_mma_plan(int C, int sm_count, size_t smem_limit) {
  MmaPlan pl;
  pl.C = C;
  const int nq = (C + 3) / 4;
  std::vector<std::vector<char>> freeb(nq, std::vector<char>(nq, 0));
  size_t nfree = 0;
  for (int j = 0; j < nq; ++j)
    for (int k = 0; k <= j; ++k) { freeb[j][k] = 1; ++nfree; }
  // same greedy clustering as the duo plan: a kind keeps a quad set D, takes the free regions
  // inside D x D and grows D by the quad that unlocks the most free regions; regions are spread
  // evenly over the minimum number of kinds (24 cameras: 21 regions = 3 kinds x 7)
  const int nk_min = (int)((nfree + MMA_REGIONS - 1) / MMA_REGIONS);
  const int cap = (int)((nfree + nk_min - 1) / nk_min);
  while (nfree > 0) {
    MmaKind K{};
    K.reg_base = (int)pl.regions.size();
    std::vector<char> inD(nq, 0);
    std::vector<std::pair<int, int>> blks;
    while (nfree > 0 && (int)blks.size() < cap) {
      int bj = -1, bk = -1;
      for (int j = 0; j < nq && bj < 0; ++j) {
        if (!inD[j]) continue;
        for (int k = 0; k <= j; ++k)
          if (inD[k] && freeb[j][k]) { bj = j; bk = k; break; }
      }
      if (bj < 0) {
        int best = -1, gain_best = 0;
        for (int d = 0; d < nq; ++d) {
          if (inD[d]) continue;
          int gain = freeb[d][d] ? 1 : 0;
          for (int x = 0; x < nq; ++x)
            if (inD[x]) gain += (x < d) ? freeb[d][x] : freeb[x][d];
          if (gain > gain_best) { gain_best = gain; best = d; }
        }
        if (best < 0) {
          for (int j = 0; j < nq && best < 0; ++j)
            for (int k = 0; k <= j; ++k)
              if (freeb[j][k]) { inD[j] = 1; inD[k] = 1; best = j; break; }
        } else {
          inD[best] = 1;
        }
        continue;
      }
      freeb[bj][bk] = 0;
      --nfree;
      blks.push_back({bj, bk});
    }
    bool used[LCBA_MAX_CAMERAS] = {false};
    for (auto& bl : blks)
      for (int d = 0; d < 4; ++d) {
        if (4 * bl.first + d < C) used[4 * bl.first + d] = true;
        if (4 * bl.second + d < C) used[4 * bl.second + d] = true;
      }
    int slot_of[LCBA_MAX_CAMERAS];
    for (int c = 0; c < C; ++c)
      if (used[c]) { slot_of[c] = K.nslots; K.slot_cam[K.nslots++] = (uint8_t)c; }
    for (auto& bl : blks) {
      MmaRegion r{};
      for (int d = 0; d < 4; ++d) {
        const int cr = 4 * bl.first + d, cc = 4 * bl.second + d;
        r.crow[d] = (int8_t)(cr < C ? cr : -1);
        r.srow[d] = (int8_t)(cr < C ? slot_of[cr] : 0);
        r.ccol[d] = (int8_t)(cc < C ? cc : -1);
        r.scol[d] = (int8_t)(cc < C ? slot_of[cc] : 0);
      }
      pl.regions.push_back(r);
    }
    K.nreg = (int)blks.size();
    K.rp = K.nslots * 12;
    while (K.rp % 16 != 4) ++K.rp;
    const size_t fixed = (size_t)C * CAMTAB * 8 + 64;
    int sp = (int)((smem_limit - fixed) / ((size_t)3 * K.rp * 8 * SCHUR_STAGES));
    sp = std::min(16, sp / 4 * 4);
    K.sp = std::max(4, sp);
    pl.smem_bytes = std::max(pl.smem_bytes, (size_t)3 * K.sp * K.rp * 8 * SCHUR_STAGES + fixed);
    pl.kinds.push_back(K);
  }
  pl.nkinds = (int)pl.kinds.size();
  pl.nslices = std::max(1, sm_count / pl.nkinds);
  return pl;
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// One (point, camera slot) item of the producer: 12 rows x 3 kappa into the ring.
__device__ __forceinline__ void mma_produce(const double* __restrict__ T, const double (&X)[3],
                                            const double (&li)[9], double w, bool live,
                                            double* __restrict__ Yk0, int rp) {
  ObsLin L;
  obs_linearize<false>(T, X[0], X[1], X[2], 0.0, 0.0, w, L, live);
  double Q[2][3];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    Q[i][0] = L.Jp[i][0] * li[0];
    Q[i][1] = fma(L.Jp[i][0], li[1], L.Jp[i][1] * li[2]);
    Q[i][2] = fma(L.Jp[i][0], li[3], fma(L.Jp[i][1], li[4], L.Jp[i][2] * li[5]));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double y[9];
#pragma unroll
    for (int a = 0; a < 9; ++a) y[a] = fma(L.Jc[0][a], Q[0][k], L.Jc[1][a] * Q[1][k]);
    double2* o = reinterpret_cast<double2*>(Yk0 + (size_t)k * rp);
    o[0] = make_double2(y[0], y[1]);
    o[1] = make_double2(y[2], y[3]);
    o[2] = make_double2(y[4], y[5]);
    o[3] = make_double2(y[6], y[7]);
    o[4] = make_double2(y[8], w * Q[0][k]);
    o[5] = make_double2(w * Q[1][k], li[6 + k]);      // row 11 = z: rhs for free (schur.cuh)
  }
}

// part[slice][npairs*121 + 11 C]: - sum Y Y^T per pair block (11 x 11) and the reduced rhs.
__global__ void __launch_bounds__(MMA_THREADS, 1)
k_schur_mma(const double* __restrict__ tab, const double* __restrict__ pts,
            const double* __restrict__ wgt, const uint32_t* __restrict__ obs_start,
            const unsigned long long* __restrict__ mask, const double* __restrict__ Lz,
            long long P, long long N, int C, const MmaKind* __restrict__ kinds,
            const MmaRegion* __restrict__ regions, int nslices, size_t part_stride, int npairs,
            double* __restrict__ part, long long* __restrict__ stats) {
  extern __shared__ __align__(16) double s_dyn[];
  const MmaKind& K = kinds[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nreg = K.nreg, nthreads = 32 * (nreg + MMA_PROD_WARPS);   // barrier participants
  const int nslots = K.nslots, SP = K.sp, rp = K.rp;
  const size_t stage_doubles = (size_t)3 * SP * rp;
  double* s_tab = s_dyn + SCHUR_STAGES * stage_doubles;
  enum { BAR_FULL = 2, BAR_EMPTY = BAR_FULL + SCHUR_STAGES, BAR_PROD = BAR_EMPTY + SCHUR_STAGES };
  const long long t_start = clock64();
  long long t_wait = 0;

  const int slice = blockIdx.x;
  long long pa, pb;
  {
    const unsigned long long ta = (unsigned long long)N * slice / nslices;
    const unsigned long long tb = (unsigned long long)N * (slice + 1) / nslices;
    pa = (slice == 0) ? 0 : lower_bound_u32(obs_start, P, ta);
    pb = (slice == nslices - 1) ? P : lower_bound_u32(obs_start, P, tb);
  }
  const long long nchunks = (pb - pa + SP - 1) / SP;

  if (wid >= MMA_REGIONS) {
    // ================================ producer warpgroup ================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " MMA_PROD_REGS ";");
    const int ptid = tid - 32 * MMA_REGIONS, nprod = 32 * MMA_PROD_WARPS;
    for (int i = ptid; i < C * CAMTAB; i += nprod) s_tab[i] = tab[i];
    nbar_sync(BAR_PROD, nprod);
    const int total = SP * nslots;
    for (long long c = 0; c < nchunks + SCHUR_STAGES; ++c) {
      const int st = (int)(c % SCHUR_STAGES);
      if (c >= SCHUR_STAGES) { const long long t0 = clock64(); nbar_sync(BAR_EMPTY + st, nthreads); t_wait += clock64() - t0; }
      if (c >= nchunks) continue;
      double* s_Y = s_dyn + st * stage_doubles;
      const long long q0 = pa + c * SP;
      for (int idx = ptid; idx < total; idx += nprod) {
        const int q = idx / nslots, sl = idx - q * nslots;
        const long long p = q0 + q;
        double* Yk0 = s_Y + (size_t)3 * q * rp + sl * 12;
        if (p >= pb) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double2* o = reinterpret_cast<double2*>(Yk0 + (size_t)k * rp);
#pragma unroll
            for (int e = 0; e < 6; ++e) o[e] = make_double2(0.0, 0.0);
          }
          continue;
        }
        const unsigned long long m = mask[p];
        const int cam = K.slot_cam[sl];
        const bool live = (m >> cam) & 1ull;
        double w = 0.0;
        if (live) w = wgt ? wgt[(long long)obs_start[p] + __popcll(m & ((1ull << cam) - 1ull))] : 1.0;
        double X[3], li[9];
#pragma unroll
        for (int a = 0; a < 3; ++a) X[a] = pts[3 * p + a];
#pragma unroll
        for (int a = 0; a < 9; ++a) li[a] = Lz[p * 9 + a];
        mma_produce(s_tab + cam * CAMTAB, X, li, w, live, Yk0, rp);
      }
      __threadfence_block();
      nbar_arrive(BAR_FULL + st, nthreads);
    }
    if (stats && ptid == 0) {
      long long* o = stats + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
      o[2] = t_wait;
      o[3] = clock64() - t_start;
    }
    return;
  }

  // ================================ consumer warpgroups ================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 " MMA_CONS_REGS ";");
  if (wid >= nreg) return;                           // no region for this warp in this kind
  const MmaRegion R = regions[K.reg_base + wid];
  int offA[6], offB[6];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int rho = 8 * t + (lane >> 2);
    offA[t] = R.srow[rho / 12] * 12 + rho % 12 + (lane & 3) * rp;
    offB[t] = R.scol[rho / 12] * 12 + rho % 12 + (lane & 3) * rp;
  }
  double acc[6][6][2];
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int u = 0; u < 6; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;

  const int ksteps = 3 * SP / 4;
  for (long long c = 0; c < nchunks; ++c) {
    const int st = (int)(c % SCHUR_STAGES);
    const double* s_Y = s_dyn + st * stage_doubles;
    { const long long t0 = clock64(); nbar_sync(BAR_FULL + st, nthreads); t_wait += clock64() - t0; }
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      const double* base = s_Y + (size_t)ks * 4 * rp;
      double a[6], b[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) { a[t] = base[offA[t]]; b[t] = base[offB[t]]; }
#pragma unroll
      for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int u = 0; u < 6; ++u) dmma884(acc[t][u], a[t], b[u]);
    }
    nbar_arrive(BAR_EMPTY + st, nthreads);
  }
  if (stats && tid == 0) {
    long long* o = stats + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
    o[0] = t_wait;
    o[1] = clock64() - t_start;
  }
  // ---------------- slice partial: every lower-triangle pair entry exactly once ----------------
  double* out = part + (size_t)slice * part_stride;
  double* rout = out + (size_t)npairs * 121;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int rho = 8 * t + (lane >> 2);
    const int cj = R.crow[rho / 12], ra = rho % 12;
    if (cj < 0 || ra >= NCP) continue;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int sig = 8 * u + 2 * (lane & 3) + e;
        const int ck = R.ccol[sig / 12], cb = sig % 12;
        if (ck < 0 || ck > cj) continue;
        if (cb < NCP) out[(size_t)(cj * (cj + 1) / 2 + ck) * 121 + ra * NCP + cb] = -acc[t][u][e];
        else if (ck == cj) rout[cj * NCP + ra] = -acc[t][u][e];
      }
    }
  }
}

// ---- camera blocks U_c = sum_obs Jc^T Jc ------------------------------------------------
// thread = (camera c = t % C fixed, point lane t / C); 66 register accumulators (upper
// triangle); block partial -> part[block][C * 66].
constexpr int CAMN_VALS = 66;
__global__ void __launch_bounds__(256, 1)
k_cam_normal(const double* __restrict__ tab, const double* __restrict__ pts,
             const double* __restrict__ wgt, const uint32_t* __restrict__ obs_start,
             const unsigned long long* __restrict__ mask, long long P, int C, int PB,
             double* __restrict__ part) {
  extern __shared__ double s_dyn[];
  double* s_tab = s_dyn;                                   // C * CAMTAB
  double* s_red = s_dyn + ((C * CAMTAB + 1) & ~1);         // PB * C  (one value at a time)
  const int t = threadIdx.x;
  load_tables_smem(tab, s_tab, C);
  __syncthreads();
  const int c = t % C, pl = t / C;
  const bool worker = pl < PB;
  double acc[CAMN_VALS];
#pragma unroll
  for (int i = 0; i < CAMN_VALS; ++i) acc[i] = 0.0;
  if (worker) {
    for (long long p = (long long)blockIdx.x * PB + pl; p < P; p += (long long)gridDim.x * PB) {
      const unsigned long long m = mask[p];
      if (!((m >> c) & 1ull)) continue;
      const double w = wgt ? wgt[(long long)obs_start[p] + __popcll(m & ((1ull << c) - 1ull))] : 1.0;
      ObsLin L;
      obs_linearize<false>(s_tab + c * CAMTAB, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], 0.0, 0.0, w, L);
      double J0[11], J1[11];
#pragma unroll
      for (int a = 0; a < 9; ++a) { J0[a] = L.Jc[0][a]; J1[a] = L.Jc[1][a]; }
      J0[9] = w; J0[10] = 0.0; J1[9] = 0.0; J1[10] = w;
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 11; ++a)
#pragma unroll
        for (int b = a; b < 11; ++b, ++idx) acc[idx] = fma(J0[a], J0[b], fma(J1[a], J1[b], acc[idx]));
    }
  }
  // reduce over the PB point lanes of each camera (fixed order)
  double* out = part + (size_t)blockIdx.x * C * CAMN_VALS;
#pragma unroll
  for (int i = 0; i < CAMN_VALS; ++i) {
    __syncthreads();
    if (worker) s_red[pl * C + c] = acc[i];
    __syncthreads();
    if (t < C) {
      double s = 0.0;
      for (int q = 0; q < PB; ++q) s += s_red[q * C + t];
      out[t * CAMN_VALS + i] = s;
    }
  }
}

// Sred diagonal pair blocks += U (symmetric expansion of the 66 upper-triangle values)
__global__ void k_add_cam_blocks(const double* __restrict__ U, int C, double* __restrict__ Sred) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * 121) return;
  const int c = i / 121, e = i % 121, a = e / NCP, b = e % NCP;
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  const int idx = lo * NCP - lo * (lo - 1) / 2 + (hi - lo);
  Sred[(size_t)(c * (c + 1) / 2 + c) * 121 + e] += U[c * CAMN_VALS + idx];
}

}  // namespace lcba
