// Measures the FP64 DFMA (CUDA-core) and DMMA (mma.sync m8n8k4 f64) peaks of the device:
// the second roofline denominator of the Schur kernel (MEASURED_PEAKS.json has HBM/bf16 only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
  double c0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[2 * i]), "+d"(c0[2 * i + 1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c1[2 * i]), "+d"(c1[2 * i + 1]) : "d"(b), "d"(a));
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    for (int bps : {1, 2, 4}) {
      if (threads * bps > 2048) continue;
      k_dfma<16><<<sms * bps, threads>>>(out, 100, 1.0000001, 1e-9);
      cudaEventRecord(e0);
      k_dfma<16><<<sms * bps, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flop = 2.0 * 16 * iters * (double)threads * bps * sms;
      printf("DFMA threads=%d blocks/SM=%d: %.2f ms  %.2f TFLOP/s\n", threads, bps, ms, flop / ms * 1e-9);
    }
  }
  for (int threads : {256, 512, 1024}) {
    k_dmma<<<sms * 2, threads>>>(out, 100);
    cudaEventRecord(e0);
    k_dmma<<<sms * 2, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per warp per mma: 8*8*4*2 flop; 8 mma per iteration
    const double flop = 512.0 * 8 * iters * (threads / 32.0) * 2 * sms;
    printf("DMMA threads=%d blocks/SM=2: %.2f ms  %.2f TFLOP/s\n", threads, ms, flop / ms * 1e-9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s, SMs %d, clock %d kHz\n", cudaGetErrorString(e), sms, p.clockRate);
  return e != cudaSuccess;
}
