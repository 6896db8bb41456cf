// Feasibility probe for a tensor-core (DMMA, mma.sync m8n8k4 f64) Schur consumer: every warp
// owns a 48 x 48 region of S = Y Y^T (4 x 4 cameras, 12 padded rows each), Y staged in shared
// memory as [kappa][rows] with a row stride = 4 (mod 16) doubles so that the fragment loads
// (lane -> row lane/4, kappa lane%4) are conflict free.  Per K-step (4 kappa): 12 LDS.64 and
// 36 DMMA per warp.  Checks the fragment layout against a host SYRK and reports TFLOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_syrk tools/dmma_syrk.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int NSLOT = 24, ROWS = NSLOT * 12, RP = ROWS + 4;   // 292 = 4 mod 16
constexpr int KST = 48;                                        // kappa per stage (16 points x 3)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// regions: per warp 8 slots (4 row cams, 4 col cams)
template <int NREG>
__global__ void __maxnreg__(NREG)
k_syrk(const double* __restrict__ Y /* [KST][RP] */, const int* __restrict__ regions, int reps,
       double* __restrict__ out /* [blocks][warps][48*48] */) {
  extern __shared__ double sY[];
  for (int i = threadIdx.x; i < KST * RP; i += blockDim.x) sY[i] = Y[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int* rg = regions + wid * 8;
  int offA[6], offB[6];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    const int rho = 8 * t + (lane >> 2);
    offA[t] = rg[rho / 12] * 12 + rho % 12 + (lane & 3) * RP;
    offB[t] = rg[4 + rho / 12] * 12 + rho % 12 + (lane & 3) * RP;
  }
  double acc[6][6][2];
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int u = 0; u < 6; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int k0 = 0; k0 < KST; k0 += 4) {
      const double* base = sY + k0 * RP;
      double a[6], b[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) { a[t] = base[offA[t]]; b[t] = base[offB[t]]; }
#pragma unroll
      for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int u = 0; u < 6; ++u) dmma(acc[t][u], a[t], b[u]);
    }
  }
  double* o = out + ((size_t)blockIdx.x * (blockDim.x >> 5) + wid) * 48 * 48;
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int row = 8 * t + (lane >> 2), col = 8 * u + 2 * (lane & 3);
      o[row * 48 + col] = acc[t][u][0];
      o[row * 48 + col + 1] = acc[t][u][1];
    }
}

int main(int argc, char** argv) {
  const int warps = argc > 1 ? atoi(argv[1]) : 8;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  std::vector<double> hY((size_t)KST * RP);
  srand(1);
  for (auto& v : hY) v = (rand() % 2001 - 1000) * 1e-3;
  std::vector<int> hreg(warps * 8);
  for (int w = 0; w < warps; ++w)
    for (int i = 0; i < 8; ++i) hreg[w * 8 + i] = (w * 5 + i * 3 + (i >= 4 ? 7 : 0)) % NSLOT;
  double *dY, *dout;
  int* dreg;
  cudaMalloc(&dY, hY.size() * 8);
  cudaMalloc(&dreg, hreg.size() * 4);
  cudaMalloc(&dout, (size_t)sms * warps * 48 * 48 * 8);
  cudaMemcpy(dY, hY.data(), hY.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dreg, hreg.data(), hreg.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)KST * RP * 8;
  cudaFuncSetAttribute(k_syrk<200>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // correctness: one repetition against the host
  k_syrk<200><<<sms, warps * 32, smem>>>(dY, dreg, 1, dout);
  std::vector<double> ho((size_t)warps * 48 * 48);
  cudaMemcpy(ho.data(), dout, ho.size() * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int w = 0; w < warps; ++w)
    for (int r = 0; r < 48; ++r)
      for (int c = 0; c < 48; ++c) {
        const int ra = hreg[w * 8 + r / 12] * 12 + r % 12, cb = hreg[w * 8 + 4 + c / 12] * 12 + c % 12;
        double s = 0;
        for (int k = 0; k < KST; ++k) s += hY[(size_t)k * RP + ra] * hY[(size_t)k * RP + cb];
        const double e = fabs(s - ho[((size_t)w * 48 + r) * 48 + c]);
        if (e > maxerr) maxerr = e;
      }
  printf("fragment layout check: max |err| = %.3e\n", maxerr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 4000;
  cudaEventRecord(e0);
  k_syrk<200><<<sms, warps * 32, smem>>>(dY, dreg, reps, dout);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 512.0 * 36 * (KST / 4) * (double)reps * warps * sms;
  printf("warps=%d: %.3f ms, %.2f TFLOP/s (DMMA incl. fragment loads), %.1f cycles per warp K-step at %d kHz\n",
         warps, ms, flop / ms * 1e-9, ms * 1e-3 * p.clockRate * 1e3 / ((double)reps * KST / 4), p.clockRate);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess || maxerr > 1e-9;
}
