// Normal equations of the two squared-residual variants of the reference
//   fun_camonly              (pySBA.py:151-156): cameras free, points fixed
//   fun_transform_points_3d  (pySBA.py:176-188): one affine map [A|t] of all points, cameras fixed
// both with f = w * (proj - obs)^2 per pixel coordinate, handed by the reference to a dense
// scipy least_squares (tr_solver='exact').  With e = proj - obs the Jacobian row of f_d is
// 2 w e_d * d proj_d / d theta; one pass over the resident observation stream produces
// cost = 0.5 sum f^2, g = J^T f and J^T J, which is all the dense trust-region loop needs
// (lasercalib_b200/_trf_dense.py).  HBM-bound: 29 B per observation, one read.
#pragma once
#include "common.cuh"

namespace lcba {

constexpr int SQ_THREADS = 256;
constexpr int SQC_VALS = 77;    // per camera: upper triangle of the 11x11 block (66) then g (11)
constexpr int SQT_M = 6, SQT_Q = 10;
constexpr int SQT_VALS = SQT_M * SQT_Q + 12;   // Kronecker factors (60) then g (12)

__host__ __device__ inline int sq_camonly_warps(int C, size_t smem_limit_bytes) {
  const long long avail = (long long)(smem_limit_bytes / 8) - ((C * CAMTAB + 1) & ~1) - 64;
  long long w = avail / ((long long)C * SQC_VALS);
  return (int)(w > 8 ? 8 : (w < 1 ? 1 : w));
}

// part[block][C*77 + 1]: per-camera sums then the block's sum of f^2.
// Per-camera accumulation: warp-private tables in shared memory; lanes of a warp that hold
// the same camera take turns (rank among their peers), so no atomics and a fixed order.
template <bool DERIVS>
__global__ void __launch_bounds__(SQ_THREADS)
k_sq_camonly(const double* __restrict__ tab, const double* __restrict__ pts,
             const double2* __restrict__ uv, const uint8_t* __restrict__ cam,
             const int32_t* __restrict__ pt, const double* __restrict__ wgt, long long N, int C,
             double* __restrict__ part) {
  extern __shared__ double s_dyn[];
  __shared__ double s_red[32];
  double* s_tab = s_dyn;
  double* s_acc = s_dyn + ((C * CAMTAB + 1) & ~1);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nwarps = blockDim.x >> 5;
  const int K = C * SQC_VALS;
  load_tables_smem(tab, s_tab, C);
  if (DERIVS)
    for (int i = t; i < nwarps * K; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  double* my = s_acc + (size_t)wid * K;
  double ss = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < N; base += stride) {
    const long long o = base + t;
    int c8 = 255;
    double J0[11], J1[11], f0 = 0.0, f1 = 0.0;
    if (o < N) {
      c8 = cam[o];
      const long long p = pt[o];
      const double2 ob = uv[o];
      const double w = wgt ? wgt[o] : 1.0;
      ObsLin L;
      if (DERIVS) {
        obs_linearize<true>(s_tab + c8 * CAMTAB, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], ob.x,
                            ob.y, 1.0, L);
      } else {
        double pu, pv;
        project_tab(s_tab + c8 * CAMTAB, pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], pu, pv);
        L.ru = pu - ob.x;
        L.rv = pv - ob.y;
      }
      const double su = w * L.ru, sv = w * L.rv;
      f0 = su * L.ru;
      f1 = sv * L.rv;
      ss = fma(f0, f0, fma(f1, f1, ss));
      if (DERIVS) {
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          J0[a] = 2.0 * su * L.Jc[0][a];
          J1[a] = 2.0 * sv * L.Jc[1][a];
        }
        J0[9] = 2.0 * su;  J0[10] = 0.0;
        J1[9] = 0.0;       J1[10] = 2.0 * sv;
      }
    }
    if (DERIVS) {
      const unsigned peers = __match_any_sync(0xffffffffu, c8);
      const int rank = (c8 == 255) ? -1 : __popc(peers & ((1u << lane) - 1u));
      const int maxrank = __reduce_max_sync(0xffffffffu, rank);
      for (int r = 0; r <= maxrank; ++r) {
        if (rank == r) {
          double* row = my + c8 * SQC_VALS;
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 11; ++a)
#pragma unroll
            for (int b = a; b < 11; ++b, ++idx) row[idx] += fma(J0[a], J0[b], J1[a] * J1[b]);
#pragma unroll
          for (int a = 0; a < 11; ++a) row[66 + a] += fma(J0[a], f0, J1[a] * f1);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * (K + 1);
  if (DERIVS)
    for (int i = t; i < K; i += blockDim.x) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += s_acc[(size_t)w * K + i];
      out[i] = s;
    }
  const double tot = block_sum(ss, s_red);
  if (t == 0) out[K] = tot;
}

// part[block][73]: T[m][q] = sum M_m Q_q with M = sum_d s_d^2 Jp_d Jp_d^T (3x3 sym, 6 values),
// Q = Xh Xh^T (4x4 sym, 10 values), Xh = (X,1) the UNtransformed point: J^T J is the
// Kronecker-structured sum of M (x) Q; then g[k][j] = sum (Jp^T (s f))_k Xh_j; then sum f^2.
template <bool DERIVS>
__global__ void __launch_bounds__(SQ_THREADS)
k_sq_transform(const double* __restrict__ tab, const double* __restrict__ pts,
               const double2* __restrict__ uv, const uint8_t* __restrict__ cam,
               const int32_t* __restrict__ pt, const double* __restrict__ wgt, long long N, int C,
               const double* __restrict__ theta, double* __restrict__ part) {
  extern __shared__ double s_dyn[];
  __shared__ double s_red[SQ_THREADS / 32][SQT_VALS + 1];
  double* s_tab = s_dyn;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  load_tables_smem(tab, s_tab, C);
  double th[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) th[i] = theta[i];
  __syncthreads();
  double acc[DERIVS ? SQT_VALS : 1];
#pragma unroll
  for (int i = 0; i < (DERIVS ? SQT_VALS : 1); ++i) acc[i] = 0.0;
  double ss = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long o = (long long)blockIdx.x * blockDim.x + t; o < N; o += stride) {
    const int c8 = cam[o];
    const long long p = pt[o];
    const double2 ob = uv[o];
    const double w = wgt ? wgt[o] : 1.0;
    const double Xh[4] = {pts[3 * p], pts[3 * p + 1], pts[3 * p + 2], 1.0};
    double Xt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      Xt[k] = fma(th[4 * k], Xh[0], fma(th[4 * k + 1], Xh[1], fma(th[4 * k + 2], Xh[2], th[4 * k + 3])));
    ObsLin L;
    if (DERIVS) {
      obs_linearize<true>(s_tab + c8 * CAMTAB, Xt[0], Xt[1], Xt[2], ob.x, ob.y, 1.0, L);
    } else {
      double pu, pv;
      project_tab(s_tab + c8 * CAMTAB, Xt[0], Xt[1], Xt[2], pu, pv);
      L.ru = pu - ob.x;
      L.rv = pv - ob.y;
    }
    const double su = w * L.ru, sv = w * L.rv;
    const double f0 = su * L.ru, f1 = sv * L.rv;
    ss = fma(f0, f0, fma(f1, f1, ss));
    if (DERIVS) {
      double a0[3], a1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a0[k] = 2.0 * su * L.Jp[0][k];
        a1[k] = 2.0 * sv * L.Jp[1][k];
      }
      double M[SQT_M], Q[SQT_Q];
      int i = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = k; l < 3; ++l) M[i++] = fma(a0[k], a0[l], a1[k] * a1[l]);
      i = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int l = j; l < 4; ++l) Q[i++] = Xh[j] * Xh[l];
#pragma unroll
      for (int m = 0; m < SQT_M; ++m)
#pragma unroll
        for (int q = 0; q < SQT_Q; ++q) acc[m * SQT_Q + q] = fma(M[m], Q[q], acc[m * SQT_Q + q]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double v = fma(a0[k], f0, a1[k] * f1);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[60 + 4 * k + j] = fma(v, Xh[j], acc[60 + 4 * k + j]);
      }
    }
  }
  // block reduction, fixed order
  if (DERIVS) {
#pragma unroll
    for (int i = 0; i < SQT_VALS; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[wid][i] = v;
    }
  }
  {
    double v = ss;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[wid][SQT_VALS] = v;
  }
  __syncthreads();
  for (int i = t; i <= SQT_VALS; i += blockDim.x) {
    double s = 0.0;
    if (DERIVS || i == SQT_VALS)
      for (int w = 0; w < SQ_THREADS / 32; ++w) s += s_red[w][i];
    part[(size_t)blockIdx.x * (SQT_VALS + 1) + i] = s;
  }
}

}  // namespace lcba
