// How many DISTINCT register operands can a DFMA stream sustain at full rate on B200?
// mode 0: acc[k] = fma(a[k], b[k], acc[k])   (3 distinct register operands, no reuse)
// mode 1: acc[k] = fma(a0,   b[k], acc[k])   (rank-1 update style: one operand reusable)
// mode 2: acc[k] = fma(a0,   b0,   acc[k])   (two operands reusable)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, const double* in, int iters) {
  double acc[16], a[16], b[16];
  for (int i = 0; i < 16; ++i) { acc[i] = in[i] * threadIdx.x; a[i] = in[16 + i] + 1e-9 * threadIdx.x; b[i] = in[32 + i] - 1e-9 * threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) acc[i] = fma(a[i], b[i], acc[i]);
      if (MODE == 1) acc[i] = fma(a[0], b[i], acc[i]);
      if (MODE == 2) acc[i] = fma(a[0], b[0], acc[i]);
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, threads = 512, iters = 20000;
  double *out, *in; cudaMalloc(&out, 8 * sms * 2 * threads); cudaMalloc(&in, 8 * 48);
  double h[48]; for (int i = 0; i < 48; ++i) h[i] = 1.0 + 1e-7 * i; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<sms * 2, threads>>>(out, in, iters);
      if (mode == 1) k<1><<<sms * 2, threads>>>(out, in, iters);
      if (mode == 2) k<2><<<sms * 2, threads>>>(out, in, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d: %.2f ms  %.2f TFLOP/s\n", mode, ms, 2.0 * 16 * iters * threads * 2.0 * sms / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
