// Do DFMA (FP64 pipe) and DMMA (mma.sync m8n8k4 f64) overlap on B200?  Half the warps of every
// CTA run a DFMA chain, the other half a DMMA chain; compare with each alone.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_mix(double* out, int iters, int mode) {   // mode 0: DFMA only, 1: DMMA only, 2: split by warp
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 1 || (mode == 2 && (warp & 1));
  double s = 0;
  if (!do_mma) {
    double acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], 1.0000001, 1e-9);
    for (int i = 0; i < 16; ++i) s += acc[i];
  } else {
    double c0[8] = {0}, c1[8] = {0};
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[2 * i]), "+d"(c0[2 * i + 1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c1[2 * i]), "+d"(c1[2 * i + 1]) : "d"(b), "d"(a));
      }
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, threads = 512, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 2 * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    k_mix<<<sms * 2, threads>>>(out, 100, mode);
    cudaEventRecord(e0);
    k_mix<<<sms * 2, threads>>>(out, iters, mode);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = threads / 32.0 * 2 * sms;
    const double f_fma = 2.0 * 16 * iters * 32, f_mma = 512.0 * 8 * iters;
    double flop = mode == 0 ? warps * f_fma : mode == 1 ? warps * f_mma : warps / 2 * (f_fma + f_mma);
    printf("mode %d: %.2f ms  %.2f TFLOP/s\n", mode, ms, flop / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
