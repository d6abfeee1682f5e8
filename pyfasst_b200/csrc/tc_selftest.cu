// Self-test of the tcgen05 building blocks (tc.cuh): D[128][N] = A B^T in tf32 (optionally
// 3xTF32) for every combination of K-major / MN-major operands.  One CTA; used by
// tests/test_tc_gpu.py to pin the descriptor / layout conventions the production kernels
// rely on.
#include "common.cuh"
#include "tc.cuh"

namespace pf {

// A global: a_mn ? [K][128] : [128][K];  B global: b_mn ? [K][N] : [N][K];  D: [128][N]
__global__ void __launch_bounds__(128)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                   int N, int K, int a_mn, int b_mn, int split3) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = 128 * 32 * 4, b_bytes = (uint32_t)N * 32 * 4;
  // the swizzled operand tiles need a 1024-byte aligned base (the swizzle is a function of
  // the shared-memory address bits)
  unsigned char* base = smem + ((1024 - (tc::smem_u32(smem) & 1023)) & 1023);
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;
  // MN-major tiles of a 32-row K chunk: per 32-wide MN group all 32 k rows are stacked
  const uint32_t lbo = 32 * 128, sbo = 512;

  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) tc::tmem_alloc(&tmem_base, ncols);
  if (tid == 0) {
    tc::mbar_init(&mbar, 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::idesc_tf32(128, N, a_mn, b_mn);

  uint32_t phase = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int idx = tid; idx < 128 * 32; idx += 128) {
      int r, k;
      float x;
      if (a_mn) { k = idx / 128; r = idx % 128; x = A[(size_t)(k0 + k) * 128 + r]; }
      else { r = idx / 32; k = idx % 32; x = A[(size_t)r * K + k0 + k]; }
      float hi, lo;
      tc::split_tf32(x, hi, lo);
      if (!split3) hi = x;
      const uint32_t off = a_mn ? tc::mnmajor_off(k, r, lbo, sbo) : tc::kmajor_off(r, k);
      *reinterpret_cast<float*>(a_hi + off) = hi;
      *reinterpret_cast<float*>(a_lo + off) = lo;
    }
    for (int idx = tid; idx < N * 32; idx += 128) {
      int n, k;
      float x;
      if (b_mn) { k = idx / N; n = idx % N; x = B[(size_t)(k0 + k) * N + n]; }
      else { n = idx / 32; k = idx % 32; x = B[(size_t)n * K + k0 + k]; }
      float hi, lo;
      tc::split_tf32(x, hi, lo);
      if (!split3) hi = x;
      const uint32_t off = b_mn ? tc::mnmajor_off(k, n, lbo, sbo) : tc::kmajor_off(n, k);
      *reinterpret_cast<float*>(b_hi + off) = hi;
      *reinterpret_cast<float*>(b_lo + off) = lo;
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      for (int j = 0; j < 4; ++j) {
        // K step j: 32 bytes further along a K-major row, 8 rows (1024 bytes) further down an
        // MN-major column
        const uint32_t ao = a_mn ? j * 1024 : j * 32, bo = b_mn ? j * 1024 : j * 32;
        const uint32_t pah = tc::smem_u32(a_hi) + ao, pal = tc::smem_u32(a_lo) + ao;
        const uint32_t pbh = tc::smem_u32(b_hi) + bo, pbl = tc::smem_u32(b_lo) + bo;
        const uint64_t dah = a_mn ? tc::smem_desc_mnmajor(pah, lbo, sbo) : tc::smem_desc_kmajor(pah);
        const uint64_t dal = a_mn ? tc::smem_desc_mnmajor(pal, lbo, sbo) : tc::smem_desc_kmajor(pal);
        const uint64_t dbh = b_mn ? tc::smem_desc_mnmajor(pbh, lbo, sbo) : tc::smem_desc_kmajor(pbh);
        const uint64_t dbl = b_mn ? tc::smem_desc_mnmajor(pbl, lbo, sbo) : tc::smem_desc_kmajor(pbl);
        tc::mma_tf32(tmem, dah, dbh, idesc, (k0 > 0 || j > 0) ? 1u : 0u);
        if (split3) {
          tc::mma_tf32(tmem, dah, dbl, idesc, 1u);
          tc::mma_tf32(tmem, dal, dbh, idesc, 1u);
        }
      }
      tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, phase);  // operands may be overwritten / D is complete
    phase ^= 1;
    __syncthreads();
  }
  tc::fence_after_thread_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tc::tmem_ld_wait();
    const int row = warp * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace pf

using namespace pf;

extern "C" int pf_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn,
                              int b_mn, int split3, void* stream) {
  PF_REQUIRE(N >= 32 && N <= 256 && N % 32 == 0, "pf_tc_selftest: N=%d must be a multiple of 32 <= 256", N);
  PF_REQUIRE(K >= 32 && K % 32 == 0, "pf_tc_selftest: K=%d must be a multiple of 32", K);
  const size_t smem = 2 * (128 * 32 * 4) + 2 * ((size_t)N * 32 * 4) + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) {
    set_error("pf_tc_selftest: %s", cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  tc_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(A, B, D, N, K, a_mn, b_mn, split3);
  return check_launch("tc_selftest_kernel");
}
