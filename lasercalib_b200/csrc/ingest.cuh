// Ingest of the reference's wire format (PySBA.__init__, pySBA.py:28-59; observation
// order of scripts/get_points3d.py:74-86): validate, sort point-major / camera-ascending
// when needed, narrow int64 indices, build the per-point CSR offsets and visibility masks.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace lcba {

// flags[0] |= 1 index out of range ; |= 2 not sorted by (pt, cam)
// pt_off: first point of this shard (the caller may pass global point indices)
__global__ void k_make_keys(const long long* __restrict__ cam_idx,
                            const long long* __restrict__ pt_idx, long long N, int C, long long P,
                            long long pt_off, unsigned long long* __restrict__ keys,
                            int* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long long c = cam_idx[i], p = pt_idx[i] - pt_off;
  if (c < 0 || c >= C || p < 0 || p >= P) {
    atomicOr(flags, 1);
    keys[i] = 0;
    return;
  }
  const unsigned long long k = ((unsigned long long)p << 8) | (unsigned long long)c;
  keys[i] = k;
  if (i > 0) {
    const long long c0 = cam_idx[i - 1], p0 = pt_idx[i - 1] - pt_off;
    if (p0 > p || (p0 == p && c0 > c)) atomicOr(flags, 2);
  }
}

__global__ void k_iota(int32_t* __restrict__ v, long long N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) v[i] = (int32_t)i;
}

// keys sorted.  Narrow to (uint8 cam, int32 pt), gather uv / weights through perm (null =
// identity), mark CSR starts.  flags |= 4 on duplicate (camera, point).
__global__ void k_narrow_gather(const unsigned long long* __restrict__ keys,
                                const int32_t* __restrict__ perm,
                                const double2* __restrict__ uv_in, const double* __restrict__ w_in,
                                long long N, uint8_t* __restrict__ cam, int32_t* __restrict__ pt,
                                double2* __restrict__ uv, double* __restrict__ w,
                                int* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned long long k = keys[i];
  cam[i] = (uint8_t)(k & 0xff);
  pt[i] = (int32_t)(k >> 8);
  const long long src = perm ? (long long)perm[i] : i;
  uv[i] = uv_in[src];
  if (w) w[i] = w_in[src];
  if (i > 0 && keys[i - 1] == k) atomicOr(flags, 4);
}

// obs_start[p] = first sorted observation of point p (P+1 entries, empty ranges allowed)
__global__ void k_obs_start(const int32_t* __restrict__ pt, long long N, long long P,
                            uint32_t* __restrict__ obs_start) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > N) return;
  const long long lo = (i == 0) ? 0 : (long long)pt[i - 1] + 1;
  const long long hi = (i == N) ? P : (long long)pt[i];
  for (long long p = lo; p <= hi; ++p) obs_start[p] = (uint32_t)i;
}

// per-point visibility mask and the maximum observations per point
__global__ void k_point_masks(const uint8_t* __restrict__ cam, const uint32_t* __restrict__ obs_start,
                              long long P, unsigned long long* __restrict__ mask,
                              int* __restrict__ kmax) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const uint32_t a = obs_start[p], b = obs_start[p + 1];
  unsigned long long m = 0;
  for (uint32_t i = a; i < b; ++i) m |= 1ull << cam[i];
  mask[p] = m;
  atomicMax(kmax, (int)(b - a));
}

// ---- repeated (camera, point) rows ------------------------------------------------------
// The reference accepts any number of rows for one (camera, point) pair (each is one more
// residual pair; its own 3-dataset concatenation produces them, calibrate_camera.py:41-44).
// The observation-major passes handle them as they come.  The Schur-side passes work per
// distinct pair: rows of one pair have the same Jacobian up to their weight, so the pair
// enters U, W and Y with the weight sqrt(sum w^2).
__global__ void k_pair_first(const uint8_t* __restrict__ cam, const int32_t* __restrict__ pt,
                             long long N, int* __restrict__ first) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  first[i] = (i == 0 || cam[i] != cam[i - 1] || pt[i] != pt[i - 1]) ? 1 : 0;
}

__global__ void k_pair_weights(const int* __restrict__ first, const int* __restrict__ rank,
                               const double* __restrict__ w, long long N,
                               double* __restrict__ pair_w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || !first[i]) return;
  double s = 0.0;
  long long j = i;
  do {
    const double wj = w ? w[j] : 1.0;
    s = fma(wj, wj, s);
    ++j;
  } while (j < N && !first[j]);
  pair_w[rank[i]] = (j == i + 1) ? (w ? w[i] : 1.0) : sqrt(s);
}

__global__ void k_pair_start(const uint32_t* __restrict__ obs_start, const int* __restrict__ rank,
                             long long P, long long N, long long n_pairs,
                             uint32_t* __restrict__ pair_start) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > P) return;
  const uint32_t o = obs_start[p];
  pair_start[p] = (o < N) ? (uint32_t)rank[o] : (uint32_t)n_pairs;
}

}  // namespace lcba
