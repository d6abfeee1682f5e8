// Microbenchmark: mma.sync m8n8k4 f64 (DMMA) issue rate per SM sub-partition, alone and interleaved
// with DFMA -- do they share the FP64 pipe on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NMMA, int NFMA>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double d[8][2];
  double x[8];
  for (int i = 0; i < 8; ++i) { d[i][0] = i; d[i][1] = -i; x[i] = threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NMMA) dmma(d[i][0], d[i][1], a, b);
      if (i < NFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
  long long h[8];
  const int iters = 2048;
#define RUN(M, F, threads) \
  k<M, F><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DMMA x%d + DFMA x%d per iteration, %d threads: %.1f cycles per iteration\n", M, F, threads, (double)h[0] / iters);
  RUN(1, 0, 32) RUN(4, 0, 32) RUN(8, 0, 32) RUN(0, 8, 32) RUN(8, 8, 32) RUN(4, 8, 32) RUN(2, 8, 32)
  RUN(8, 0, 128) RUN(8, 0, 256) RUN(0, 8, 256) RUN(8, 8, 256) RUN(2, 8, 256) RUN(1, 8, 256)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
