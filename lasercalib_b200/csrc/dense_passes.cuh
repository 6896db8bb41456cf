// Streaming passes for DENSE rigs (most cameras see most points; no repeated (camera, point)
// rows): the same arithmetic as linearize.cuh, organised so that a thread keeps ONE camera for
// its whole life:
//     thread t of a 256-thread block  ->  camera c = t % C, point lane pl = t / C
//     tile = PB = 256 / C consecutive points; observation of (point p, camera c) =
//            obs_start[p] + popc(mask[p] & ((1 << c) - 1))        (point-major, camera-ascending)
// Consequences (what the round-1 ncu capture asked for, profiles/r01_ncu_full_summary.txt):
//   * per-camera sums (g_c, diag(J^T J), and the full camera blocks U_c) accumulate in
//     REGISTERS: no shared-memory read-modify-write per observation (k_linearize: 22 RMWs per
//     observation, shared-memory pipe 69 %, 50.9 M bank conflicts); k_cam_normal's separate model
//     evaluation disappears (U_c comes out of the same pass);
//   * the inputs of the NEXT tile are loaded before the current tile is evaluated (the round-1
//     kernels stalled 50 % on the index -> point -> table load chain);
//   * per-point sums go through one double-buffered shared-memory tile, one barrier per tile,
//     fixed summation order (bit-reproducible, no atomics).
// Invisible (point, camera) pairs idle their thread for that tile, so the sparse rigs keep the
// observation-major kernels of linearize.cuh.
#pragma once
#include "common.cuh"
#include "linearize.cuh"

namespace lcba {

constexpr int DP_THREADS = 256;
// per camera: 66 upper-triangle entries of U_c = sum Jc^T Jc (row-major a <= b), then g_c (11)
constexpr int DP_CAM_VALS = 66 + 11;

// Two-stage prefetch.  The observation index of a (point, camera) item depends on the point's
// mask and CSR offset, so a one-step prefetch still stalls on that dependent load (ncu round 2,
// first version: long-scoreboard stalls 3-11 per issued instruction with 2 warps per scheduler).
// Stage A loads (mask, obs_start) two tiles ahead, stage B turns them into the item's loads one
// tile ahead, stage C computes: no load is consumed in the iteration that issued it.
struct DenseIdx {
  unsigned long long m;
  uint32_t os;
  int ok;
};

struct DenseItem {          // one (point, camera) work item
  double X[3];
  double E[3];              // optional per-point 3-vector (g~_p)
  double2 ob;
  double w;
  int live;
};

__device__ __forceinline__ DenseIdx dense_idx(long long p, long long P, bool worker,
                                              const uint32_t* __restrict__ obs_start,
                                              const unsigned long long* __restrict__ mask) {
  DenseIdx d;
  d.ok = worker && p < P;
  d.m = 0ull;
  d.os = 0u;
  if (d.ok) { d.m = mask[p]; d.os = obs_start[p]; }
  return d;
}

__device__ __forceinline__ DenseItem dense_item(const DenseIdx& d, long long p, int c,
                                                const double* __restrict__ pts,
                                                const double2* __restrict__ uv,
                                                const double* __restrict__ wgt,
                                                const double* __restrict__ extra3 = nullptr) {
  DenseItem it;
  it.E[0] = it.E[1] = it.E[2] = 0.0;
  it.w = 0.0;
  it.ob = make_double2(0.0, 0.0);
  it.X[0] = it.X[1] = it.X[2] = 0.0;
  it.live = d.ok && ((d.m >> c) & 1ull);
  if (it.live) {
    const long long o = (long long)d.os + __popcll(d.m & ((1ull << c) - 1ull));
    if (uv) it.ob = uv[o];
    it.w = wgt ? wgt[o] : 1.0;
    it.X[0] = pts[3 * p];
    it.X[1] = pts[3 * p + 1];
    it.X[2] = pts[3 * p + 2];
    if (extra3) { it.E[0] = extra3[3 * p]; it.E[1] = extra3[3 * p + 1]; it.E[2] = extra3[3 * p + 2]; }
  }
  return it;
}

// U (66 upper-triangle entries, row-major a <= b) += Jc^T Jc with the structural zeros of the
// cx / cy columns (Jc[0][10] = Jc[1][9] = 0, Jc[0][9] = Jc[1][10] = w) written out by hand.
__device__ __forceinline__ void accumulate_U(double (&acc)[66 + 11], const ObsLin& L, double w) {
  int idx = 0;
#pragma unroll
  for (int a = 0; a < 9; ++a) {
#pragma unroll
    for (int b = a; b < 9; ++b, ++idx)
      acc[idx] = fma(L.Jc[0][a], L.Jc[0][b], fma(L.Jc[1][a], L.Jc[1][b], acc[idx]));
    acc[idx] = fma(L.Jc[0][a], w, acc[idx]); ++idx;      // (a, cx)
    acc[idx] = fma(L.Jc[1][a], w, acc[idx]); ++idx;      // (a, cy)
  }
  acc[idx] = fma(w, w, acc[idx]); idx += 2;              // (cx, cx); (cx, cy) stays 0
  acc[idx] = fma(w, w, acc[idx]);                        // (cy, cy)
}

// ---- linearise: V, g_p per point; U_c, g_c per camera; cost --------------------------------
// Vg[p][0..5] = V (00,01,02,11,12,22), Vg[p][6..8] = g_p   (same layout as k_linearize)
// cam_part[block][C][DP_CAM_VALS], cost_part[block]
// dynamic smem (doubles): tab[C*CAMTAB | even] pv[2][DP_THREADS*9] red[DP_THREADS]
__host__ __device__ inline size_t dense_lin_smem_doubles(int C) {
  return (size_t)((C * CAMTAB + 1) & ~1) + 2 * DP_THREADS * 9 + DP_THREADS;
}

__global__ void __launch_bounds__(DP_THREADS, 1)
k_linearize_dense(const double* __restrict__ tab, const double* __restrict__ pts,
                  const double2* __restrict__ uv, const double* __restrict__ wgt,
                  const uint32_t* __restrict__ obs_start,
                  const unsigned long long* __restrict__ mask, long long P, int C, int PB,
                  double* __restrict__ Vg, double* __restrict__ cam_part,
                  double* __restrict__ cost_part) {
  extern __shared__ double s_dyn[];
  double* s_tab = s_dyn;
  double* s_pv = s_dyn + ((C * CAMTAB + 1) & ~1);
  double* s_red = s_pv + 2 * DP_THREADS * 9;
  __shared__ double s_blk[32];
  const int t = threadIdx.x;
  load_tables_smem(tab, s_tab, C);
  const int c = t % C, pl = t / C;
  const bool worker = pl < PB;
  // the thread's camera never changes: volatile keeps the table reads in shared memory (hoisting
  // 29 doubles into registers next to 77 accumulators spills)
  const volatile double* vT = s_tab + c * CAMTAB;
  double acc[DP_CAM_VALS];
#pragma unroll
  for (int i = 0; i < DP_CAM_VALS; ++i) acc[i] = 0.0;
  double cost = 0.0;
  const long long ntiles = (P + PB - 1) / PB;
  long long tile = blockIdx.x;
  const long long G = gridDim.x;
  DenseItem nxt = dense_item(dense_idx(tile * PB + pl, P, worker && tile < ntiles, obs_start, mask),
                             tile * PB + pl, c, pts, uv, wgt);
  DenseIdx idx2 = dense_idx((tile + G) * PB + pl, P, worker && tile + G < ntiles, obs_start, mask);
  __syncthreads();
  int buf = 0;
  for (; tile < ntiles; tile += G, buf ^= 1) {
    const DenseItem cur = nxt;
    nxt = dense_item(idx2, (tile + G) * PB + pl, c, pts, uv, wgt);
    idx2 = dense_idx((tile + 2 * G) * PB + pl, P, worker && tile + 2 * G < ntiles, obs_start, mask);
    double* pv = s_pv + (size_t)buf * DP_THREADS * 9 + t * 9;
    if (cur.live) {
      double T[CAMTAB];
#pragma unroll
      for (int i = 0; i < CAMTAB; ++i) T[i] = vT[i];
      ObsLin L;
      obs_linearize<true>(T, cur.X[0], cur.X[1], cur.X[2], cur.ob.x, cur.ob.y, cur.w, L);
      cost = fma(L.ru, L.ru, fma(L.rv, L.rv, cost));
      pv[0] = fma(L.Jp[0][0], L.Jp[0][0], L.Jp[1][0] * L.Jp[1][0]);
      pv[1] = fma(L.Jp[0][0], L.Jp[0][1], L.Jp[1][0] * L.Jp[1][1]);
      pv[2] = fma(L.Jp[0][0], L.Jp[0][2], L.Jp[1][0] * L.Jp[1][2]);
      pv[3] = fma(L.Jp[0][1], L.Jp[0][1], L.Jp[1][1] * L.Jp[1][1]);
      pv[4] = fma(L.Jp[0][1], L.Jp[0][2], L.Jp[1][1] * L.Jp[1][2]);
      pv[5] = fma(L.Jp[0][2], L.Jp[0][2], L.Jp[1][2] * L.Jp[1][2]);
      pv[6] = fma(L.Jp[0][0], L.ru, L.Jp[1][0] * L.rv);
      pv[7] = fma(L.Jp[0][1], L.ru, L.Jp[1][1] * L.rv);
      pv[8] = fma(L.Jp[0][2], L.ru, L.Jp[1][2] * L.rv);
      accumulate_U(acc, L, cur.w);
#pragma unroll
      for (int a = 0; a < 9; ++a) acc[66 + a] = fma(L.Jc[0][a], L.ru, fma(L.Jc[1][a], L.rv, acc[66 + a]));
      acc[66 + 9] = fma(cur.w, L.ru, acc[66 + 9]);
      acc[66 + 10] = fma(cur.w, L.rv, acc[66 + 10]);
    } else if (worker) {
#pragma unroll
      for (int e = 0; e < 9; ++e) pv[e] = 0.0;
    }
    __syncthreads();
    // per-point sums over the cameras, fixed order; the other threads go on with the next tile
    const long long p0 = tile * PB;
    const int npts = (int)min((long long)PB, P - p0);
    for (int j = t; j < npts * 9; j += DP_THREADS) {     // (more than one round only below 9 cameras)
      const int q = j / 9, e = j - 9 * q;
      const double* src = s_pv + (size_t)buf * DP_THREADS * 9 + (size_t)q * C * 9 + e;
      double s = 0.0;
      for (int k = 0; k < C; ++k) s += src[k * 9];
      Vg[(p0 + q) * 9 + e] = s;
    }
  }
  // block partial of the camera sums: over the PB point lanes of each camera, fixed order
  double* out = cam_part + (size_t)blockIdx.x * C * DP_CAM_VALS;
#pragma unroll
  for (int i = 0; i < DP_CAM_VALS; ++i) {
    __syncthreads();
    if (worker) s_red[pl * C + c] = acc[i];
    __syncthreads();
    if (t < C) {
      double s = 0.0;
      for (int q = 0; q < PB; ++q) s += s_red[q * C + t];
      out[t * DP_CAM_VALS + i] = s;
    }
  }
  const double cs = block_sum(cost, s_blk);
  if (t == 0) cost_part[blockIdx.x] = cs;
}

// camsum[c][0..10] = g_c, [11..21] = diag(J^T J) (the layout k_ctl_lin reads) and U[c][66] from
// the reduced block partials red[c][DP_CAM_VALS]
__global__ void k_dense_cam_unpack(const double* __restrict__ red, int C, double* __restrict__ camsum,
                                   double* __restrict__ U) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * DP_CAM_VALS) return;
  const int c = i / DP_CAM_VALS, v = i % DP_CAM_VALS;
  const double x = red[i];
  if (v >= 66) { camsum[c * CAMSUM + (v - 66)] = x; return; }
  U[c * 66 + v] = x;
  // diagonal entries of the upper triangle: index of (a, a) = a * 11 - a (a - 1) / 2
  int a = 0, idx = 0;
  while (idx < v) { idx += NCP - a; ++a; }
  if (idx == v) camsum[c * CAMSUM + NCP + a] = x;
}

// ---- residual (fun): cost only ---------------------------------------------------------------
__global__ void __launch_bounds__(DP_THREADS, 2)
k_residual_dense(const double* __restrict__ tab, const double* __restrict__ pts,
                 const double2* __restrict__ uv, const double* __restrict__ wgt,
                 const uint32_t* __restrict__ obs_start,
                 const unsigned long long* __restrict__ mask, long long P, int C, int PB,
                 double* __restrict__ part) {
  extern __shared__ double s_dyn[];
  double* s_tab = s_dyn;
  __shared__ double s_red[32];
  const int t = threadIdx.x;
  load_tables_smem(tab, s_tab, C);
  __syncthreads();
  const int c = t % C, pl = t / C;
  const bool worker = pl < PB;
  double T[CAMTAB];               // only the 17 entries project_tab reads stay live (registers)
#pragma unroll
  for (int i = 0; i < CAMTAB; ++i) T[i] = s_tab[c * CAMTAB + i];
  double acc = 0.0;
  const long long ntiles = (P + PB - 1) / PB;
  long long tile = blockIdx.x;
  const long long G = gridDim.x;
  DenseItem nxt = dense_item(dense_idx(tile * PB + pl, P, worker && tile < ntiles, obs_start, mask),
                             tile * PB + pl, c, pts, uv, wgt);
  DenseIdx idx2 = dense_idx((tile + G) * PB + pl, P, worker && tile + G < ntiles, obs_start, mask);
  for (; tile < ntiles; tile += G) {
    const DenseItem cur = nxt;
    nxt = dense_item(idx2, (tile + G) * PB + pl, c, pts, uv, wgt);
    idx2 = dense_idx((tile + 2 * G) * PB + pl, P, worker && tile + 2 * G < ntiles, obs_start, mask);
    if (cur.live) {
      double pu, pv;
      project_tab(T, cur.X[0], cur.X[1], cur.X[2], pu, pv);
      const double ru = cur.w * (pu - cur.ob.x), rv = cur.w * (pv - cur.ob.y);
      acc = fma(ru, ru, fma(rv, rv, acc));
    }
  }
  const double s = block_sum(acc, s_red);
  if (t == 0) part[blockIdx.x] = s;
}

// ---- |J g~|^2 ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(DP_THREADS, 2)
k_jdot_dense(const double* __restrict__ tab, const double* __restrict__ pts,
             const double* __restrict__ wgt, const uint32_t* __restrict__ obs_start,
             const unsigned long long* __restrict__ mask, const double* __restrict__ gt_c,
             const double* __restrict__ gt_p, long long P, int C, int PB,
             double* __restrict__ part) {
  extern __shared__ double s_dyn[];
  double* s_tab = s_dyn;
  __shared__ double s_red[32];
  const int t = threadIdx.x;
  load_tables_smem(tab, s_tab, C);
  __syncthreads();
  const int c = t % C, pl = t / C;
  const bool worker = pl < PB;
  const volatile double* vT = s_tab + c * CAMTAB;
  double gc[NCP];
#pragma unroll
  for (int a = 0; a < NCP; ++a) gc[a] = gt_c[c * NCP + a];
  double acc = 0.0;
  const long long ntiles = (P + PB - 1) / PB;
  long long tile = blockIdx.x;
  const long long G = gridDim.x;
  DenseItem nxt = dense_item(dense_idx(tile * PB + pl, P, worker && tile < ntiles, obs_start, mask),
                             tile * PB + pl, c, pts, nullptr, wgt, gt_p);
  DenseIdx idx2 = dense_idx((tile + G) * PB + pl, P, worker && tile + G < ntiles, obs_start, mask);
  for (; tile < ntiles; tile += G) {
    const DenseItem cur = nxt;
    nxt = dense_item(idx2, (tile + G) * PB + pl, c, pts, nullptr, wgt, gt_p);
    idx2 = dense_idx((tile + 2 * G) * PB + pl, P, worker && tile + 2 * G < ntiles, obs_start, mask);
    if (cur.live) {
      double T[CAMTAB];
#pragma unroll
      for (int i = 0; i < CAMTAB; ++i) T[i] = vT[i];
      ObsLin L;
      obs_linearize<false>(T, cur.X[0], cur.X[1], cur.X[2], 0.0, 0.0, cur.w, L);
      double a0 = cur.w * gc[9], a1 = cur.w * gc[10];
#pragma unroll
      for (int a = 0; a < 9; ++a) { a0 = fma(L.Jc[0][a], gc[a], a0); a1 = fma(L.Jc[1][a], gc[a], a1); }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        a0 = fma(L.Jp[0][a], cur.E[a], a0);
        a1 = fma(L.Jp[1][a], cur.E[a], a1);
      }
      acc = fma(a0, a0, fma(a1, a1, acc));
    }
  }
  const double s = block_sum(acc, s_red);
  if (t == 0) part[blockIdx.x] = s;
}

}  // namespace lcba
