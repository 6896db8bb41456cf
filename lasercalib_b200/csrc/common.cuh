// Handle, error plumbing and small device utilities shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/lcba.h"
#include "model.cuh"

namespace lcba {

constexpr int SM_COUNT_B200 = 148;

#define LCBA_CUDA(h, call)                                                          \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      set_error(h, std::string(#call) + ": " + cudaGetErrorString(e_));             \
      return LCBA_E_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define LCBA_TRY(expr)                 \
  do {                                 \
    int rc_ = (expr);                  \
    if (rc_ != LCBA_OK) return rc_;    \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* red /* >= 32 doubles smem */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;
}

// first point p with obs_start[p] >= target  (obs_start has P+1 entries, non-decreasing)
__device__ __forceinline__ long long lower_bound_u32(const uint32_t* __restrict__ a, long long n,
                                                     unsigned long long target) {
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((unsigned long long)a[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace lcba
