// Dense FP64 Cholesky solve of the reduced camera system (n = 11 C <= 704) on the device:
// blocked right-looking factorisation (32-wide panels) + blocked triangular solves.
// Replaces the LSMR call of scipy's TRF (scipy/optimize/_lsq/trf.py:494-495) for the camera
// block; a non-positive pivot raises fail[0] (the caller adds damping and retries).
#pragma once
#include "common.cuh"

namespace lcba {

constexpr int CH_NB = 32;

// Factor the nb x nb diagonal block at (k,k) and solve the panel below it.  One CTA.
__global__ void __launch_bounds__(256)
k_chol_panel(double* __restrict__ A, int n, int k, int* __restrict__ fail) {
  __shared__ double D[CH_NB][CH_NB + 1];
  const int t = threadIdx.x;
  const int nb = min(CH_NB, n - k);
  for (int e = t; e < CH_NB * CH_NB; e += blockDim.x) {
    const int i = e / CH_NB, j = e % CH_NB;
    D[i][j] = (i < nb && j < nb && j <= i) ? A[(size_t)(k + i) * n + k + j] : (i == j ? 1.0 : 0.0);
  }
  __syncthreads();
  // right-looking factorisation of the 32x32 diagonal block with the whole CTA:
  // thread (ty, tx) of an 8x32 grid updates rows ty, ty+8, ... of column tx
  {
    const int tx = t & 31, ty = t >> 5;
    for (int j = 0; j < nb; ++j) {
      double d = D[j][j];
      if (!(d > 0.0)) { if (t == 0) atomicOr(fail, 1); d = 1.0; }
      d = sqrt(d);
      __syncthreads();
      if (t == j) D[j][j] = d;
      if (t > j && t < nb) D[t][j] /= d;
      __syncthreads();
      if (tx > j && tx < nb) {
        const double lc = D[tx][j];
        for (int r = ty; r < nb; r += 8)
          if (r >= tx) D[r][tx] -= D[r][j] * lc;
      }
      __syncthreads();
    }
  }
  for (int e = t; e < nb * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    if (j <= i) A[(size_t)(k + i) * n + k + j] = D[i][j];
  }
  // panel: rows below, X L^T = A  =>  x[c] = (a[c] - sum_{m<c} x[m] L[c][m]) / L[c][c]
  for (int r = k + nb + t; r < n; r += blockDim.x) {
    double* row = A + (size_t)r * n + k;
    double x[CH_NB];
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
    for (int c = 0; c < CH_NB; ++c) {
      double s = x[c];
#pragma unroll
      for (int m = 0; m < c; ++m) s = fma(-x[m], D[c][m], s);
      x[c] = s / D[c][c];
    }
#pragma unroll
    for (int c = 0; c < CH_NB; ++c)
      if (c < nb) row[c] = x[c];
  }
}

// Trailing update A[i][j] -= sum_c L[i][k+c] L[j][k+c] for i >= j >= k+nb (32x32 tiles).
__global__ void __launch_bounds__(256)
k_chol_update(double* __restrict__ A, int n, int k) {
  const int base = k + CH_NB;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  __shared__ double Pi[CH_NB][CH_NB + 1], Pj[CH_NB][CH_NB + 1];
  const int t = threadIdx.x;
  for (int e = t; e < CH_NB * CH_NB; e += blockDim.x) {
    const int r = e / CH_NB, c = e % CH_NB;
    const int gi = base + ti * CH_NB + r, gj = base + tj * CH_NB + r;
    Pi[r][c] = (gi < n) ? A[(size_t)gi * n + k + c] : 0.0;
    Pj[r][c] = (gj < n) ? A[(size_t)gj * n + k + c] : 0.0;
  }
  __syncthreads();
  const int cj = t % CH_NB, r0 = t / CH_NB;   // 8 row groups
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int ri = r0 + 8 * m;
    const int gi = base + ti * CH_NB + ri, gj = base + tj * CH_NB + cj;
    if (gi < n && gj < n && gj <= gi) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) s = fma(Pi[ri][c], Pj[cj][c], s);
      A[(size_t)gi * n + gj] -= s;
    }
  }
}

// Solve L L^T x = b with the factor in the lower triangle of A.  One CTA of 256 threads,
// x and b in global memory (b is overwritten by y then x is written to out).
__global__ void __launch_bounds__(256)
k_chol_solve(const double* __restrict__ A, int n, const double* __restrict__ b,
             double* __restrict__ out) {
  extern __shared__ double y[];   // n
  const int t = threadIdx.x, lane = t & 31;
  for (int i = t; i < n; i += blockDim.x) y[i] = b[i];
  __syncthreads();
  // forward: L y = b
  for (int kb = 0; kb < n; kb += CH_NB) {
    const int nb = min(CH_NB, n - kb);
    if (t < 32) {
      double lrow[CH_NB];
#pragma unroll
      for (int j = 0; j < CH_NB; ++j)
        lrow[j] = (lane < nb && j <= lane && j < nb) ? A[(size_t)(kb + lane) * n + kb + j] : 0.0;
      double v = (lane < nb) ? y[kb + lane] : 0.0;
#pragma unroll
      for (int j = 0; j < CH_NB; ++j) {
        const double djj = __shfl_sync(0xffffffffu, lrow[j], j);
        double yj = __shfl_sync(0xffffffffu, v, j);
        yj = (j < nb) ? yj / djj : 0.0;
        if (lane == j) v = yj;
        else if (lane > j) v = fma(-lrow[j], yj, v);
      }
      if (lane < nb) y[kb + lane] = v;
    }
    __syncthreads();
    for (int i = kb + nb + t; i < n; i += blockDim.x) {
      const double* row = A + (size_t)i * n + kb;
      double s = y[i];
      for (int c = 0; c < nb; ++c) s = fma(-row[c], y[kb + c], s);
      y[i] = s;
    }
    __syncthreads();
  }
  // backward: L^T x = y
  const int nblk = (n + CH_NB - 1) / CH_NB;
  for (int bi = nblk - 1; bi >= 0; --bi) {
    const int kb = bi * CH_NB;
    const int nb = min(CH_NB, n - kb);
    if (t < 32) {
      double lcol[CH_NB];   // lcol[j] = L[kb+j][kb+lane], j >= lane
#pragma unroll
      for (int j = 0; j < CH_NB; ++j)
        lcol[j] = (lane < nb && j >= lane && j < nb) ? A[(size_t)(kb + j) * n + kb + lane] : 0.0;
      double v = (lane < nb) ? y[kb + lane] : 0.0;
#pragma unroll
      for (int jj = 0; jj < CH_NB; ++jj) {
        const int j = CH_NB - 1 - jj;
        // diagonal element of row j lives in lane j, lcol[j]
        const double djj = __shfl_sync(0xffffffffu, lcol[j], j);
        double xj = __shfl_sync(0xffffffffu, v, j);
        xj = (j < nb) ? xj / djj : 0.0;
        if (lane == j) v = xj;
        else if (lane < j) v = fma(-lcol[j], xj, v);
      }
      if (lane < nb) y[kb + lane] = v;
    }
    __syncthreads();
    for (int i = t; i < kb; i += blockDim.x) {
      double s = y[i];
      for (int c = 0; c < nb; ++c) s = fma(-A[(size_t)(kb + c) * n + i], y[kb + c], s);
      y[i] = s;
    }
    __syncthreads();
  }
  for (int i = t; i < n; i += blockDim.x) out[i] = y[i];
}

// A_out = A_in + mu * diag(d^2)   (copy for the in-place factorisation)
__global__ void k_copy_damped(const double* __restrict__ Ain, int n, double mu,
                              const double* __restrict__ d, double* __restrict__ Aout) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * n) return;
  double v = Ain[idx];
  const int r = (int)(idx / n), c = (int)(idx % n);
  if (r == c && mu != 0.0) v = fma(mu * d[r], d[r], v);
  Aout[idx] = v;
}

}  // namespace lcba
