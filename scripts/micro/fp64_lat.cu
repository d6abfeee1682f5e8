// Microbenchmark: latency / issue interval of DFMA, DMUL and F2F on one warp, and FP64 throughput
// per SM with 1..8 warps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_lat.cu -o fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[CHAINS];
  for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void f2f_chain(float* out, long long* cyc, int iters) {
  float x = threadIdx.x * 1e-3f + 1.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double d = (double)x;          // F2F.F64.F32
    d = d * 1.0000001;             // DMUL
    x = (float)d;                  // F2F.F32.F64
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; float* fo;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096); cudaMalloc(&fo, 4096);
  long long h[8];
  const int iters = 4096;
#define RUN(C, threads) \
  dfma_chain<C><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DFMA chains=%d threads=%d: %.2f cycles per DFMA-step (%.2f per instr)\n", C, threads, (double)h[0] / iters, (double)h[0] / iters / C);
  RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(16, 32)
  RUN(8, 64) RUN(8, 128) RUN(8, 256) RUN(16, 256) RUN(8, 512)
  f2f_chain<<<1, 32>>>(fo, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("F2F.F64.F32 + DMUL + F2F.F32.F64 chain: %.2f cycles per round\n", (double)h[0] / iters);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
