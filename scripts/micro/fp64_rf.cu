// Microbenchmark: DFMA throughput with three DISTINCT register operands per instruction (the shape
// of the moment accumulation acc[p][c] += w[p] * m[c]) against the 2-register-operand chain of
// fp64_lat.cu -- does the register file / operand collector sustain 1 DFMA per 2 cycles?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_rf.cu -o fp64_rf
#include <cstdio>
#include <cuda_runtime.h>
template <int NW, int NM>
__global__ void outer(double* out, long long* cyc, int iters, const double* in) {
  double w[NW], m[NM], acc[NW][NM];
  for (int i = 0; i < NW; ++i) w[i] = in[threadIdx.x + i];
  for (int i = 0; i < NM; ++i) m[i] = in[threadIdx.x + 64 + i];
  for (int i = 0; i < NW; ++i)
    for (int j = 0; j < NM; ++j) acc[i][j] = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NW; ++i)
#pragma unroll
      for (int j = 0; j < NM; ++j) acc[i][j] = fma(w[i], m[j], acc[i][j]);
    // keep w, m changing so that nothing is hoisted (2 extra DADD per iteration)
    w[it & (NW - 1)] += 1e-9;
    m[it & (NM - 1)] += 1e-9;
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NW; ++i)
    for (int j = 0; j < NM; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// the same outer product, program order forced by volatile asm: consecutive DFMAs share w[i]
template <int NW, int NM>
__global__ void outer_asm(double* out, long long* cyc, int iters, const double* in) {
  double w[NW], m[NM], acc[NW][NM];
  for (int i = 0; i < NW; ++i) w[i] = in[threadIdx.x + i];
  for (int i = 0; i < NM; ++i) m[i] = in[threadIdx.x + 64 + i];
  for (int i = 0; i < NW; ++i)
    for (int j = 0; j < NM; ++j) acc[i][j] = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NW; ++i)
#pragma unroll
      for (int j = 0; j < NM; ++j)
        asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i][j]) : "d"(w[i]), "d"(m[j]));
    w[it & (NW - 1)] += 1e-9;
    m[it & (NM - 1)] += 1e-9;
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NW; ++i)
    for (int j = 0; j < NM; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double *out, *in; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0, 1 << 16);
  long long h[8];
  const int iters = 2048;
#define RUN(W, M, threads) \
  outer<W, M><<<1, threads>>>(out, cyc, iters, in); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("outer %dx%d threads=%d: %.2f cycles per DFMA (per warp: x%d warps per scheduler)\n", W, M, threads, (double)h[0] / iters / (W * M), (threads + 127) / 128);
  RUN(4, 4, 32) RUN(8, 4, 32) RUN(8, 4, 128) RUN(8, 4, 256) RUN(8, 4, 384) RUN(16, 4, 256) RUN(4, 8, 256) RUN(2, 2, 512)
#define RUNA(W, M, threads) \
  outer_asm<W, M><<<1, threads>>>(out, cyc, iters, in); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("outer_asm %dx%d threads=%d: %.2f cycles per DFMA (per warp: x%d warps per scheduler)\n", W, M, threads, (double)h[0] / iters / (W * M), (threads + 127) / 128);
  RUNA(8, 4, 128) RUNA(8, 4, 256) RUNA(16, 4, 256) RUNA(4, 8, 256)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
