// K4 on the 5th-generation tensor cores (tcgen05, TMEM accumulators): the three dense
// contractions of the spectral M-step for float32 planes,
//     V[f,n]     = sum_k W[f,k] H[k,n]                          (spec_power_tc_kernel)
//     num[f,k]   = sum_n (hatW/P)[f,n] G[k,n]                   (fb_contract_tc_kernel)
//     num/den[k,n] = sum_f W[f,k] (O hatW/P'^2 | O/P')[f,n]      (tw_contract_tc_kernel)
// replacing FASST.comp_spat_comp_power (pyfasst/audioModel.py:430-498) and the contractions of
// FASST.update_spectral_components (audioModel.py:1521-1575, :1634-1727).
//
// On the CUDA cores these cost 32 + 32 + 96 FMAs per TF bin per source at K = 32 against 20
// bytes of traffic, i.e. they are FMA-bound (profiles/r01).  kind::tf32 MMAs with the 3xTF32
// split (x = hi + lo, A B ~= Ah Bh + Ah Bl + Al Bh, error ~2^-21 per product) keep float32-class
// accuracy -- the reference tolerances are 1e-4 on the factors -- and make all three
// memory-bound.  The A operands are *computed* (elementwise functions of the hatW / P planes),
// so they cannot come from TMA: threads load the planes, form the operand in registers, split
// it and store the hi / lo tiles into shared memory in the canonical UMMA layouts of tc.cuh;
// one elected thread issues the MMAs; completion is tracked with mbarriers (tcgen05.commit) so
// that the next tile is produced while the tensor core consumes the previous one.
#include "common.cuh"
#include "tc.cuh"

namespace pf {

constexpr float kEpsF = 1e-10f;

__device__ __forceinline__ float fast_rcpf(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * (2.0f - x * r);  // one Newton step: ~1 ulp
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_split4(unsigned char* hi_tile, unsigned char* lo_tile,
                                          uint32_t off, float4 x) {
  float4 h, l;
  tc::split_tf32(x.x, h.x, l.x);
  tc::split_tf32(x.y, h.y, l.y);
  tc::split_tf32(x.z, h.z, l.z);
  tc::split_tf32(x.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// ============================ FB update ===============================================
// D[128 f][32 k] += E1[128 f][32 n] G[32 k][32 n]^T per step of 32 frames; both operands are
// K-major (frames contiguous), which is how the planes and G lie in memory.
constexpr int FBT_THREADS = 256;
constexpr int FBT_ROWS = 128;
constexpr int FBT_KN = 32;  // frames per step (= one 128-byte swizzle row)

struct FbtStage {
  unsigned char a_hi[FBT_ROWS * FBT_KN * 4];
  unsigned char a_lo[FBT_ROWS * FBT_KN * 4];
  unsigned char b_hi[32 * FBT_KN * 4];
  unsigned char b_lo[32 * FBT_KN * 4];
};

__global__ void __launch_bounds__(FBT_THREADS, 2)
fb_contract_tc_kernel(const float* __restrict__ hatW, const float* __restrict__ Pp, long ld,
                      const float* __restrict__ G, long ldg, int k0, int K, int F, long N,
                      long chunk, int nsplit, double* __restrict__ num) {
  extern __shared__ __align__(1024) unsigned char fbt_smem[];
  __shared__ uint64_t mbar_free[2];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base;
  unsigned char* base = fbt_smem + ((1024 - (tc::smem_u32(fbt_smem) & 1023)) & 1023);
  FbtStage* stages = reinterpret_cast<FbtStage*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fblk = blockIdx.y * FBT_ROWS;
  const int split = blockIdx.x;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;
  const int nsteps = (int)((end - begin + FBT_KN - 1) / FBT_KN);

  if (warp == 0) tc::tmem_alloc(&tmem_base, 32);
  if (tid == 0) {
    tc::mbar_init(&mbar_free[0], 1);
    tc::mbar_init(&mbar_free[1], 1);
    tc::mbar_init(&mbar_done, 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::idesc_tf32(128, 32, 0, 0);

  // thread -> (row, 16-byte chunk) of the 128 x 32 plane tiles: 4 float4 per plane and step
  int prow[4], pchk[4];
  uint32_t poff[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + q * FBT_THREADS;
    prow[q] = i >> 3;
    pchk[q] = i & 7;
    poff[q] = (uint32_t)((prow[q] >> 3) * 1024 + (prow[q] & 7) * 128 + ((pchk[q] ^ (prow[q] & 7)) << 4));
  }
  // G tile 32 x 32: one float4 per thread
  const int grow = tid >> 3, gchk = tid & 7;
  const uint32_t goff = (uint32_t)((grow >> 3) * 1024 + (grow & 7) * 128 + ((gchk ^ (grow & 7)) << 4));
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  float4 hw_n[4], p_n[4], g_n;
  auto fetch = [&](int step) {
    const long nb = begin + (long)step * FBT_KN;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int f = fblk + prow[q];
      const long n = nb + pchk[q] * 4;
      const bool ok = (f < F) && (n + 4 <= ld);
      hw_n[q] = ok ? ldg4(hatW + (long)f * ld + n) : zero4;
      p_n[q] = ok ? ldg4(Pp + (long)f * ld + n) : zero4;
    }
    const long n = nb + gchk * 4;
    g_n = (k0 + grow < K && n + 4 <= ldg) ? ldg4(G + (long)(k0 + grow) * ldg + n) : zero4;
  };

  fetch(0);
  for (int s = 0; s < nsteps; ++s) {
    const int b = s & 1;
    float4 hw[4], p[4];
    const float4 g = g_n;
#pragma unroll
    for (int q = 0; q < 4; ++q) { hw[q] = hw_n[q]; p[q] = p_n[q]; }
    if (s + 1 < nsteps) fetch(s + 1);
    // the MMAs of step s-2 must have finished reading this ring slot
    if (s >= 2) tc::mbar_wait(&mbar_free[b], (uint32_t)(((s >> 1) - 1) & 1));
    FbtStage& st = stages[b];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 e;  // hatW / P^2 * O with O == P (audioModel.py:1565-1571)
      e.x = hw[q].x * fast_rcpf(fmaxf(p[q].x, kEpsF));
      e.y = hw[q].y * fast_rcpf(fmaxf(p[q].y, kEpsF));
      e.z = hw[q].z * fast_rcpf(fmaxf(p[q].z, kEpsF));
      e.w = hw[q].w * fast_rcpf(fmaxf(p[q].w, kEpsF));
      st_split4(st.a_hi, st.a_lo, poff[q], e);
    }
    st_split4(st.b_hi, st.b_lo, goff, g);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      const uint32_t ah = tc::smem_u32(st.a_hi), al = tc::smem_u32(st.a_lo);
      const uint32_t bh = tc::smem_u32(st.b_hi), bl = tc::smem_u32(st.b_lo);
#pragma unroll
      for (int j = 0; j < FBT_KN / 8; ++j) {
        const uint64_t dah = tc::smem_desc_kmajor(ah + j * 32), dal = tc::smem_desc_kmajor(al + j * 32);
        const uint64_t dbh = tc::smem_desc_kmajor(bh + j * 32), dbl = tc::smem_desc_kmajor(bl + j * 32);
        tc::mma_tf32(tmem, dah, dbh, idesc, (s > 0 || j > 0) ? 1u : 0u);
        tc::mma_tf32(tmem, dah, dbl, idesc, 1u);
        tc::mma_tf32(tmem, dal, dbh, idesc, 1u);
      }
      tc::mma_commit(&mbar_free[b]);
      if (s == nsteps - 1) tc::mma_commit(&mbar_done);
    }
  }
  // epilogue: TMEM -> registers -> partial numerators (warps 0-3 own TMEM lanes 32 w .. 32 w + 31)
  if (nsteps > 0) tc::mbar_wait(&mbar_done, 0);
  tc::fence_after_thread_sync();
  if (warp < 4) {
    uint32_t v[32];
    if (nsteps > 0) {
      tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
    const int f = fblk + warp * 32 + lane;
    if (f < F) {
      double* out = num + ((size_t)split * F + f) * K + k0;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k0 + k < K) out[k] = (double)__uint_as_float(v[k]);
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

}  // namespace pf

using namespace pf;

// host side: called from pf_nmf_fb_contract (nmf.cu) for float32 planes with P == O
int pf_fb_contract_tc(const float* hatW, const float* P, long ld, const float* G, long ldg, int K,
                      int F, long N, long chunk, int nsplit, double* num, cudaStream_t st) {
  const size_t smem = 2 * sizeof(FbtStage) + 1024;
  cudaError_t e = cudaFuncSetAttribute(fb_contract_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("fb_contract_tc_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  dim3 grid(nsplit, ceil_div(F, FBT_ROWS));
  for (int k0 = 0; k0 < K; k0 += 32) {
    fb_contract_tc_kernel<<<grid, FBT_THREADS, smem, st>>>(hatW, P, ld, G, ldg, k0, K, F, N, chunk,
                                                         nsplit, num);
    int rc = check_launch("fb_contract_tc_kernel");
    if (rc) return rc;
  }
  return PF_OK;
}
