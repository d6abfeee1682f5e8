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
#include <cuda.h>

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
  // chunk < 0: the splits of a row block take the 32-frame steps in turn (split x works on steps
  // x, x + nsplit, ...): together they read one contiguous region of every row at any moment
  const bool interleave = chunk < 0;
  const long step_stride = interleave ? (long)nsplit * FBT_KN : FBT_KN;
  const long begin = interleave ? (long)split * FBT_KN : (long)split * chunk;
  int nsteps;
  if (interleave) {
    const long total = (N + FBT_KN - 1) / FBT_KN;
    nsteps = split < total ? (int)((total - split + nsplit - 1) / nsplit) : 0;
  } else {
    long end = begin + chunk;
    if (end > N) end = N;
    nsteps = (int)((end - begin + FBT_KN - 1) / FBT_KN);
  }

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
    const long nb = begin + (long)step * step_stride;
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

// ============================ V = W H ====================================================
// D[128 f][256 n] = W[128 f][32 k] H[32 k][256 n]: A = W is K-major (k contiguous), B = H is
// MN-major (frames contiguous).  A CTA keeps its W block (hi / lo) in shared memory and walks
// frame tiles; the B tiles and the TMEM accumulators are double buffered so that the MMAs of
// tile t overlap the write-out of tile t-1.
constexpr int SPT_THREADS = 256;
constexpr int SPT_NT = 256;  // frames per tile (= UMMA N)
constexpr uint32_t SPT_LBO = 32 * 128, SPT_SBO = 512;

struct SptSmem {
  unsigned char a_hi[128 * 32 * 4];
  unsigned char a_lo[128 * 32 * 4];
  unsigned char b_hi[2][32 * SPT_NT * 4];
  unsigned char b_lo[2][32 * SPT_NT * 4];
  // per-warp 32 x 32 boxes of the write-out (128-byte rows in the TMA SWIZZLE_128B pattern)
  float tr[SPT_THREADS / 32][2][32 * 32];  // two boxes per warp: one is filled while the TMA reads the other
};

// The write-out goes through the TMA: a warp reads 32 rows x 32 columns of the accumulator from
// TMEM (thread = row), stores them as a swizzled 32 x 128-byte box (conflict-free 16-byte
// stores) and ONE lane hands the box to cp.async.bulk.tensor.2d (UTMASTG) -- no global store
// instructions, no bounds predicates (the tensor map clips rows >= F and columns >= ldv), and the
// copy out of shared memory overlaps the next TMEM load.
__global__ void __launch_bounds__(SPT_THREADS, 1)
spec_power_tc_kernel(const float* __restrict__ W, int ldw, const float* __restrict__ H, long ldh,
                     const __grid_constant__ CUtensorMap tmapV, long ldv, int F, int K, long N,
                     int tiles_per_cta) {
  extern __shared__ __align__(1024) unsigned char spt_smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base;
  unsigned char* base = spt_smem + ((1024 - (tc::smem_u32(spt_smem) & 1023)) & 1023);
  SptSmem& sm = *reinterpret_cast<SptSmem*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fblk = blockIdx.y * 128;
  const long ntiles = (ldv + SPT_NT - 1) / SPT_NT;
  // tiles_per_cta < 0: the CTAs of a row block take the frame tiles in turn (CTA x works on tiles
  // x, x + gridDim.x, ...), so that together they write one contiguous region of every row at any
  // moment; > 0: one contiguous run of tiles per CTA
  const bool interleave = tiles_per_cta < 0;
  const long t_step = interleave ? (long)gridDim.x : 1;
  const long t_begin = interleave ? (long)blockIdx.x : (long)blockIdx.x * tiles_per_cta;
  int nt;
  if (interleave) {
    nt = t_begin < ntiles ? (int)((ntiles - t_begin + t_step - 1) / t_step) : 0;
  } else {
    long t_end = t_begin + tiles_per_cta;
    if (t_end > ntiles) t_end = ntiles;
    nt = (int)(t_end - t_begin);
  }

  if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
  if (tid == 0) {
    tc::mbar_init(&mbar[0], 1);
    tc::mbar_init(&mbar[1], 1);
    tc::fence_mbar_init();
  }
  // W block: 128 rows x 32 k, zero padded (k >= K, f >= F)
  for (int i = tid; i < 128 * 32; i += SPT_THREADS) {
    const int r = i >> 5, k = i & 31;
    const int f = fblk + r;
    const float x = (f < F && k < K) ? W[(long)f * ldw + k] : 0.f;
    float hi, lo;
    tc::split_tf32(x, hi, lo);
    const uint32_t off = tc::kmajor_off(r, k);
    *reinterpret_cast<float*>(sm.a_hi + off) = hi;
    *reinterpret_cast<float*>(sm.a_lo + off) = lo;
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::idesc_tf32(128, SPT_NT, 0, 1);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // thread -> float4 (k, n4) of the 32 x 256 H tile: 8 per thread
  float4 h_n[8];
  auto fetch = [&](long tile) {
    const long nb = tile * SPT_NT;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = tid + q * SPT_THREADS;  // 0 .. 2047
      const int k = i >> 6, c = i & 63;
      const long n = nb + c * 4;
      h_n[q] = (k < K && n + 4 <= ldh) ? ldg4(H + (long)k * ldh + n) : zero4;
    }
  };
  // write-out of one accumulator (see above)
  int box = 0;
  auto write_out = [&](int b, long tile) {
    // warps 0-3: columns [0,128); warps 4-7: columns [128,256) of accumulator b
    const int wq = warp & 3, half = warp >> 2;
    const long nb = tile * SPT_NT + half * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * SPT_NT + half * 128 + c0), v);
      tc::tmem_ld_wait();
      float* tr = sm.tr[warp][box];
      if (lane == 0) tc::tma_store_wait_read<1>();  // the box before the previous one has left shared memory
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(tr) + lane * 128 +
                                   ((c ^ (lane & 7)) << 4)) =
            make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                        __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {  // (always a group, possibly empty: the wait above counts groups)
        if (nb + c0 < ldv) tc::tma_store_2d(&tmapV, tc::smem_u32(tr), (int)(nb + c0), fblk + wq * 32);
        tc::tma_store_commit();
      }
      box ^= 1;
    }
  };

  if (nt > 0) fetch(t_begin);
  for (int t = 0; t < nt; ++t) {
    const int b = t & 1;
    // B tile t -> smem slot b (free: the MMAs of tile t-2 were waited for in the epilogue of t-2)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = tid + q * SPT_THREADS;
      const int k = i >> 6, c = i & 63;
      st_split4(sm.b_hi[b], sm.b_lo[b], tc::mnmajor_off(k, c * 4, SPT_LBO, SPT_SBO), h_n[q]);
    }
    if (t + 1 < nt) fetch(t_begin + (t + 1) * t_step);
    tc::fence_proxy_async();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      const uint32_t ah = tc::smem_u32(sm.a_hi), al = tc::smem_u32(sm.a_lo);
      const uint32_t bh = tc::smem_u32(sm.b_hi[b]), bl = tc::smem_u32(sm.b_lo[b]);
      const uint32_t d = tmem + (uint32_t)(b * SPT_NT);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t dah = tc::smem_desc_kmajor(ah + j * 32), dal = tc::smem_desc_kmajor(al + j * 32);
        const uint64_t dbh = tc::smem_desc_mnmajor(bh + j * 1024, SPT_LBO, SPT_SBO);
        const uint64_t dbl = tc::smem_desc_mnmajor(bl + j * 1024, SPT_LBO, SPT_SBO);
        tc::mma_tf32(d, dah, dbh, idesc, j > 0 ? 1u : 0u);
        tc::mma_tf32(d, dah, dbl, idesc, 1u);
        tc::mma_tf32(d, dal, dbh, idesc, 1u);
      }
      tc::mma_commit(&mbar[b]);
    }
    if (t > 0) {  // write tile t-1 while the tensor core works on tile t
      tc::mbar_wait(&mbar[b ^ 1], (uint32_t)(((t - 1) >> 1) & 1));
      tc::fence_after_thread_sync();
      write_out(b ^ 1, t_begin + (t - 1) * t_step);
    }
  }
  if (nt > 0) {
    const int b = (nt - 1) & 1;
    tc::mbar_wait(&mbar[b], (uint32_t)(((nt - 1) >> 1) & 1));
    tc::fence_after_thread_sync();
    write_out(b, t_begin + (nt - 1) * t_step);
  }
  if (lane == 0) tc::tma_store_wait<0>();  // this warp's boxes are written before the CTA retires
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ============================ TW update ==================================================
// num^T[128 n][32 k] += E1^T[128 n][16 f] W[16 f][32 k] (and den with E2) per step of 16
// frequency rows: both operands are MN-major (A: frames contiguous, B: components contiguous).
// P' = max(W' H, eps) -- the component's power with the updated W (audioModel.py:1639-1645) --
// is read from a plane produced by spec_power_tc_kernel just before.
constexpr int TWT_THREADS = 256;
constexpr int TWT_FR = 16;   // frequency rows per step (two K = 8 MMAs)
constexpr int TWT_NT = 128;  // frames per CTA (= UMMA M)
constexpr uint32_t TWT_LBO = TWT_FR * 128, TWT_SBO = 512;

struct TwtStage {
  unsigned char a1_hi[TWT_FR * TWT_NT * 4];
  unsigned char a1_lo[TWT_FR * TWT_NT * 4];
  unsigned char a2_hi[TWT_FR * TWT_NT * 4];
  unsigned char a2_lo[TWT_FR * TWT_NT * 4];
  unsigned char b_hi[TWT_FR * 32 * 4];
  unsigned char b_lo[TWT_FR * 32 * 4];
};

__global__ void __launch_bounds__(TWT_THREADS, 2)
tw_contract_tc_kernel(const float* __restrict__ hatW, const float* __restrict__ Op,
                      const float* __restrict__ Pn, long ld, const float* __restrict__ W, int ldw,
                      int K, int F, long N, int fchunk, int fsplit, double* __restrict__ num,
                      double* __restrict__ den, long ldo) {
  extern __shared__ __align__(1024) unsigned char twt_smem[];
  __shared__ uint64_t mbar_free[2];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base;
  unsigned char* base = twt_smem + ((1024 - (tc::smem_u32(twt_smem) & 1023)) & 1023);
  TwtStage* stages = reinterpret_cast<TwtStage*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long nb = (long)blockIdx.x * TWT_NT;
  const int split = blockIdx.y;
  const int fb = split * fchunk;
  int fe = fb + fchunk;
  if (fe > F) fe = F;
  const int nsteps = (fe - fb + TWT_FR - 1) / TWT_FR;

  if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
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
  const uint32_t idesc = tc::idesc_tf32(128, 32, 1, 1);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // thread -> float4 (f, n4) of the 16 x 128 plane tiles: 2 per plane and step
  int prow[2], pcol[2];
  uint32_t poff[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = tid + q * TWT_THREADS;  // 0 .. 511
    prow[q] = i >> 5;
    pcol[q] = (i & 31) * 4;
    poff[q] = tc::mnmajor_off(prow[q], pcol[q], TWT_LBO, TWT_SBO);
  }
  // W tile 16 x 32: threads 0..127 hold 4 consecutive k of one row
  const int wrow = tid >> 3, wk = (tid & 7) * 4;
  const uint32_t woff = tc::mnmajor_off(wrow, wk, TWT_LBO, TWT_SBO);

  float4 hw_n[2], o_n[2], p_n[2], w_n;
  auto fetch = [&](int step) {
    const int f0 = fb + step * TWT_FR;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = f0 + prow[q];
      const long n = nb + pcol[q];
      const bool ok = (f < fe) && (n + 4 <= ld);
      const long off = (long)f * ld + n;
      hw_n[q] = ok ? ldg4(hatW + off) : zero4;
      o_n[q] = ok ? ldg4(Op + off) : zero4;
      p_n[q] = ok ? ldg4(Pn + off) : zero4;
    }
    w_n = zero4;
    if (tid < 128) {
      const int f = f0 + wrow;
      if (f < fe) {
        const float* wr = W + (long)f * ldw;
        w_n.x = (wk + 0 < K) ? __ldg(wr + wk + 0) : 0.f;
        w_n.y = (wk + 1 < K) ? __ldg(wr + wk + 1) : 0.f;
        w_n.z = (wk + 2 < K) ? __ldg(wr + wk + 2) : 0.f;
        w_n.w = (wk + 3 < K) ? __ldg(wr + wk + 3) : 0.f;
      }
    }
  };
  auto e12 = [&](float hw, float o, float p, float& e1, float& e2) {
    // rows beyond the shard (all-zero loads) contribute nothing: W is zero there as well
    const float oc = fmaxf(o, kEpsF);
    const float rp = fast_rcpf(fmaxf(p, kEpsF));
    e2 = oc * rp;               // other / P'            (audioModel.py:1694-1701)
    e1 = oc * (hw * rp * rp);   // other * hat_W / P'^2  (audioModel.py:1714-1720)
  };

  if (nsteps > 0) fetch(0);
  for (int s = 0; s < nsteps; ++s) {
    const int b = s & 1;
    float4 hw[2], o[2], p[2];
    const float4 w = w_n;
#pragma unroll
    for (int q = 0; q < 2; ++q) { hw[q] = hw_n[q]; o[q] = o_n[q]; p[q] = p_n[q]; }
    if (s + 1 < nsteps) fetch(s + 1);
    if (s >= 2) tc::mbar_wait(&mbar_free[b], (uint32_t)(((s >> 1) - 1) & 1));
    TwtStage& st = stages[b];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float4 e1, e2;
      e12(hw[q].x, o[q].x, p[q].x, e1.x, e2.x);
      e12(hw[q].y, o[q].y, p[q].y, e1.y, e2.y);
      e12(hw[q].z, o[q].z, p[q].z, e1.z, e2.z);
      e12(hw[q].w, o[q].w, p[q].w, e1.w, e2.w);
      st_split4(st.a1_hi, st.a1_lo, poff[q], e1);
      st_split4(st.a2_hi, st.a2_lo, poff[q], e2);
    }
    if (tid < 128) st_split4(st.b_hi, st.b_lo, woff, w);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      const uint32_t a1h = tc::smem_u32(st.a1_hi), a1l = tc::smem_u32(st.a1_lo);
      const uint32_t a2h = tc::smem_u32(st.a2_hi), a2l = tc::smem_u32(st.a2_lo);
      const uint32_t bh = tc::smem_u32(st.b_hi), bl = tc::smem_u32(st.b_lo);
#pragma unroll
      for (int j = 0; j < TWT_FR / 8; ++j) {
        const uint64_t dbh = tc::smem_desc_mnmajor(bh + j * 1024, TWT_LBO, TWT_SBO);
        const uint64_t dbl = tc::smem_desc_mnmajor(bl + j * 1024, TWT_LBO, TWT_SBO);
        const uint64_t d1h = tc::smem_desc_mnmajor(a1h + j * 1024, TWT_LBO, TWT_SBO);
        const uint64_t d1l = tc::smem_desc_mnmajor(a1l + j * 1024, TWT_LBO, TWT_SBO);
        const uint64_t d2h = tc::smem_desc_mnmajor(a2h + j * 1024, TWT_LBO, TWT_SBO);
        const uint64_t d2l = tc::smem_desc_mnmajor(a2l + j * 1024, TWT_LBO, TWT_SBO);
        const uint32_t acc = (s > 0 || j > 0) ? 1u : 0u;
        tc::mma_tf32(tmem, d1h, dbh, idesc, acc);
        tc::mma_tf32(tmem, d1h, dbl, idesc, 1u);
        tc::mma_tf32(tmem, d1l, dbh, idesc, 1u);
        tc::mma_tf32(tmem + 32, d2h, dbh, idesc, acc);
        tc::mma_tf32(tmem + 32, d2h, dbl, idesc, 1u);
        tc::mma_tf32(tmem + 32, d2l, dbh, idesc, 1u);
      }
      tc::mma_commit(&mbar_free[b]);
      if (s == nsteps - 1) tc::mma_commit(&mbar_done);
    }
  }
  if (nsteps > 0) tc::mbar_wait(&mbar_done, 0);
  tc::fence_after_thread_sync();
  if (warp < 4) {
    const long n = nb + warp * 32 + lane;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t v[32];
      if (nsteps > 0) {
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(which * 32), v);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      double* out = which ? den : num;
      if (n < ldo) {
        const bool live = n < N;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < K) out[((size_t)split * K + k) * ldo + n] = live ? (double)__uint_as_float(v[k]) : 0.0;
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

// ============================ TW update, P' formed in the kernel =========================
// The same contractions, but P' = W' H is not read from a plane: a first MMA forms
//     P'[128 n][16 f] = H^T[128 n][32 k] W'^T[32 k][16 f]
// in TMEM (A = the CTA's H tile, MN-major, staged once; B = the step's W' rows, K-major), every
// thread reads the P' values of ITS elements back with tcgen05.ld, forms the two operands and
// the main MMAs follow as above.  Per source and iteration this removes the spec_power launch
// (4 B/bin written) and the P' read (4 B/bin): 8 instead of 16 B/bin for the TW update.
// Thread mapping (dictated by TMEM: a warp reads the 32 lanes (w & 3) * 32 ..): thread = one
// frame n = 32 (w & 3) + lane and the 8 frequency rows (w >> 2) * 8 .. + 7 of the step.
// The P' MMA of step s + 1 is issued before the main MMAs of step s (double-buffered P').
constexpr int TWF_THREADS = 256;
constexpr int TWF_FR = 16;
constexpr int TWF_NT = 128;
constexpr uint32_t TWF_LBO = TWF_FR * 128, TWF_SBO = 512;   // a1 / a2 / b tiles (K = 16 rows)
constexpr uint32_t TWF_HLBO = 32 * 128;                     // H tile (K = 32 rows)
constexpr int TWF_PB = 64;   // rows per P' MMA block (= its UMMA N): 4 steps share 12 MMAs

struct TwfSmem {
  // operand tiles of the main MMAs (ONE stage: the MMAs of a step are short and have finished
  // by the time the next step's operands are formed)
  unsigned char a1_hi[TWF_FR * TWF_NT * 4];
  unsigned char a1_lo[TWF_FR * TWF_NT * 4];
  unsigned char a2_hi[TWF_FR * TWF_NT * 4];
  unsigned char a2_lo[TWF_FR * TWF_NT * 4];
  unsigned char b_hi[TWF_FR * 32 * 4];
  unsigned char b_lo[TWF_FR * 32 * 4];
  unsigned char h_hi[32 * TWF_NT * 4];   // H tile, MN-major [32 k][128 n]
  unsigned char h_lo[32 * TWF_NT * 4];
  unsigned char wp_hi[TWF_PB * 32 * 4];  // W' rows of the next P' block, K-major [64 f][32 k]
  unsigned char wp_lo[TWF_PB * 32 * 4];
  float ptile[2][TWF_FR][TWF_NT];        // P' of a step, [f][n] (transposed out of TMEM)
};

// Variants measured on configs[1] (profiles/r01): thread = one frame x 8 rows straight out of
// TMEM (4-byte accesses) 229 us per launch; the same with fewer address instructions 240 us; with
// an mbarrier hand-over to a dedicated MMA warp 293 us; this one -- P' transposed through
// shared memory one step ahead so that the threads keep 16-byte accesses -- 220 us; that one
// with the MMA issue moved to a ninth warp 322 us (9 warps of 112 registers: three of them land
// on one scheduler's 16 K registers, so only ONE CTA fits per SM -- the same holds for every
// 288-thread variant above).
__global__ void __launch_bounds__(TWF_THREADS, 2)
tw_contract_fused_tc_kernel(const float* __restrict__ hatW, const float* __restrict__ Op, long ld,
                            const float* __restrict__ W, int ldw, const float* __restrict__ H,
                            long ldh, int K, int F, long N, int fchunk, int fsplit,
                            double* __restrict__ num, double* __restrict__ den, long ldo) {
  extern __shared__ __align__(1024) unsigned char twf_smem[];
  __shared__ uint64_t s_desc[28];
  __shared__ uint64_t mbar_free;
  __shared__ uint64_t mbar_p[2];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base;
  unsigned char* base = twf_smem + ((1024 - (tc::smem_u32(twf_smem) & 1023)) & 1023);
  TwfSmem& sm = *reinterpret_cast<TwfSmem*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long nb = (long)blockIdx.x * TWF_NT;
  const int split = blockIdx.y;
  const int fb = split * fchunk;
  int fe = fb + fchunk;
  if (fe > F) fe = F;
  const int nsteps = (fe - fb + TWF_FR - 1) / TWF_FR;

  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  if (tid == 0) {
    tc::mbar_init(&mbar_free, 1);
    tc::mbar_init(&mbar_p[0], 1);
    tc::mbar_init(&mbar_p[1], 1);
    tc::mbar_init(&mbar_done, 1);
    tc::fence_mbar_init();
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // H tile 32 k x 128 n (zero beyond K and beyond the row): 4 float4 per thread
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + q * TWF_THREADS;  // 0 .. 1023
    const int k = i >> 5, c = (i & 31) * 4;
    const long n = nb + c;
    const float4 h = (k < K && n + 4 <= ldh) ? ldg4(H + (long)k * ldh + n) : zero4;
    st_split4(sm.h_hi, sm.h_lo, tc::mnmajor_off(k, c, TWF_HLBO, TWF_SBO), h);
  }
  // W rows: threads 0..127 hold 4 consecutive k of one of the 16 rows of a step
  const int wrow = tid >> 3, wk = (tid & 7) * 4;
  const uint32_t woff = tc::mnmajor_off(wrow, wk, TWF_LBO, TWF_SBO);   // main B (f = K index)
  auto load_w = [&](int step) {
    float4 w = zero4;
    if (tid < 128 && step < nsteps) {
      const int f = fb + step * TWF_FR + wrow;
      if (f < fe) {
        const float* wr = W + (long)f * ldw;
        w.x = (wk + 0 < K) ? __ldg(wr + wk + 0) : 0.f;
        w.y = (wk + 1 < K) ? __ldg(wr + wk + 1) : 0.f;
        w.z = (wk + 2 < K) ? __ldg(wr + wk + 2) : 0.f;
        w.w = (wk + 3 < K) ? __ldg(wr + wk + 3) : 0.f;
      }
    }
    return w;
  };
  // W' rows of one P' block (64 rows x 32 k = 512 float4: two per thread), K-major tile
  auto stage_wp_block = [&](int blk) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = tid + q * TWF_THREADS;  // 0 .. 511
      const int r = i >> 3, kk = (i & 7) * 4;
      const int f = fb + blk * TWF_PB + r;
      float4 w = zero4;
      if (f < fe) {
        const float* wr = W + (long)f * ldw;
        w.x = (kk + 0 < K) ? __ldg(wr + kk + 0) : 0.f;
        w.y = (kk + 1 < K) ? __ldg(wr + kk + 1) : 0.f;
        w.z = (kk + 2 < K) ? __ldg(wr + kk + 2) : 0.f;
        w.w = (kk + 3 < K) ? __ldg(wr + kk + 3) : 0.f;
      }
      float4 hi, lo;
      tc::split_tf32(w.x, hi.x, lo.x); tc::split_tf32(w.y, hi.y, lo.y);
      tc::split_tf32(w.z, hi.z, lo.z); tc::split_tf32(w.w, hi.w, lo.w);
      const uint32_t off = tc::kmajor_off(r, kk);
      *reinterpret_cast<float4*>(sm.wp_hi + off) = hi;
      *reinterpret_cast<float4*>(sm.wp_lo + off) = lo;
    }
  };
  // plane elements of this thread: float4 (f, n4) of the 16 x 128 tiles, 2 per plane and step
  int prow[2], pcol[2];
  uint32_t poff[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = tid + q * TWF_THREADS;  // 0 .. 511
    prow[q] = i >> 5;
    pcol[q] = (i & 31) * 4;
    poff[q] = tc::mnmajor_off(prow[q], pcol[q], TWF_LBO, TWF_SBO);
  }
  float4 hw_n[2], o_n[2];
  auto fetch = [&](int step) {
    const int f0 = fb + step * TWF_FR;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = f0 + prow[q];
      const long n = nb + pcol[q];
      const bool ok = (f < fe) && (n + 4 <= ld);
      const long off = (long)f * ld + n;
      hw_n[q] = ok ? ldg4(hatW + off) : zero4;
      o_n[q] = ok ? ldg4(Op + off) : zero4;
    }
  };
  // TMEM -> ptile: this thread's TMEM lane is frame nl, it moves rows fhalf * 8 .. + 7
  const int nl = (warp & 3) * 32 + lane, fhalf = warp >> 2;
  const uint32_t tmem_lane = (uint32_t)((warp & 3) * 32) << 16;

  const int nblocks = (nsteps + 3) / 4;
  float4 w_cur = load_w(0), w_n1 = load_w(1);
  stage_wp_block(0);
  tc::fence_proxy_async();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::idesc_tf32(128, 32, 1, 1);
  const uint32_t idesc_p = tc::idesc_tf32(128, TWF_PB, 1, 0);
  // every shared-memory descriptor is loop invariant (one operand stage): thread 0 builds them
  // once; issuing a step is then 28 shared loads + 36 MMAs instead of ~400 ALU instructions on
  // the critical path of warp 0
  if (tid == 0) {
    const uint32_t hh = tc::smem_u32(sm.h_hi), hl = tc::smem_u32(sm.h_lo);
    const uint32_t wh = tc::smem_u32(sm.wp_hi), wl = tc::smem_u32(sm.wp_lo);
    for (int j = 0; j < 4; ++j) {
      s_desc[j] = tc::smem_desc_mnmajor(hh + j * 1024, TWF_HLBO, TWF_SBO);
      s_desc[4 + j] = tc::smem_desc_mnmajor(hl + j * 1024, TWF_HLBO, TWF_SBO);
      s_desc[8 + j] = tc::smem_desc_kmajor(wh + j * 32);
      s_desc[12 + j] = tc::smem_desc_kmajor(wl + j * 32);
    }
    const uint32_t a1h = tc::smem_u32(sm.a1_hi), a1l = tc::smem_u32(sm.a1_lo);
    const uint32_t a2h = tc::smem_u32(sm.a2_hi), a2l = tc::smem_u32(sm.a2_lo);
    const uint32_t bh = tc::smem_u32(sm.b_hi), bl = tc::smem_u32(sm.b_lo);
    for (int j = 0; j < TWF_FR / 8; ++j) {
      s_desc[16 + 6 * j + 0] = tc::smem_desc_mnmajor(bh + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 1] = tc::smem_desc_mnmajor(bl + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 2] = tc::smem_desc_mnmajor(a1h + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 3] = tc::smem_desc_mnmajor(a1l + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 4] = tc::smem_desc_mnmajor(a2h + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 5] = tc::smem_desc_mnmajor(a2l + j * 1024, TWF_LBO, TWF_SBO);
    }
  }
  // P'(block) -> TMEM columns 64 + 64 (block & 1) .. + 63; W' rows are in wp_hi / wp_lo
  auto issue_p = [&](int blk) {
    const int step = blk;  // (mbarrier slot = block parity)
    const uint32_t d = tmem + 64u + (uint32_t)TWF_PB * (uint32_t)(blk & 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t dah = s_desc[j], dal = s_desc[4 + j], dbh = s_desc[8 + j], dbl = s_desc[12 + j];
      tc::mma_tf32(d, dah, dbh, idesc_p, j > 0 ? 1u : 0u);
      tc::mma_tf32(d, dah, dbl, idesc_p, 1u);
      tc::mma_tf32(d, dal, dbh, idesc_p, 1u);
    }
    tc::mma_commit(&mbar_p[step & 1]);
  };
  // P' of a step: wait for the MMA of its block, move this thread's 8 values from TMEM to
  // ptile[step & 1]
  auto p_to_smem = [&](int step) {
    const int blk = step >> 2, bb = blk & 1;
    tc::mbar_wait(&mbar_p[bb], (uint32_t)((blk >> 1) & 1));
    tc::fence_after_thread_sync();
    uint32_t pv[8];
    tc::tmem_ld_32x8(tmem + tmem_lane + 64u + (uint32_t)TWF_PB * (uint32_t)bb +
                         16u * (uint32_t)(step & 3) + 8u * (uint32_t)fhalf, pv);
    tc::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) sm.ptile[step & 1][fhalf * 8 + q][nl] = __uint_as_float(pv[q]);
  };
  if (nsteps > 0) {
    fetch(0);
    if (tid == 0) issue_p(0);
    p_to_smem(0);                       // (also: the P' MMA of block 0 has finished reading wp)
    tc::fence_before_thread_sync();
    __syncthreads();                    // ptile[0] complete
  }
  for (int s = 0; s < nsteps; ++s) {
    float4 hw[2], o[2], p[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) { hw[q] = hw_n[q]; o[q] = o_n[q]; }
    if (s + 1 < nsteps) fetch(s + 1);
    const float4 w_n2 = load_w(s + 2);
#pragma unroll
    for (int q = 0; q < 2; ++q)
      p[q] = *reinterpret_cast<const float4*>(&sm.ptile[s & 1][prow[q]][pcol[q]]);
    float4 e1[2], e2[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      // rows beyond the shard (all-zero loads) contribute nothing: W is zero there as well
      auto e12 = [&](float hwv, float ov, float pvv, float& a, float& c) {
        const float oc = fmaxf(ov, kEpsF);
        const float rp = fast_rcpf(fmaxf(pvv, kEpsF));
        c = oc * rp;                // other / P'            (audioModel.py:1694-1701)
        a = oc * (hwv * rp * rp);   // other * hat_W / P'^2  (audioModel.py:1714-1720)
      };
      e12(hw[q].x, o[q].x, p[q].x, e1[q].x, e2[q].x);
      e12(hw[q].y, o[q].y, p[q].y, e1[q].y, e2[q].y);
      e12(hw[q].z, o[q].z, p[q].z, e1[q].z, e2[q].z);
      e12(hw[q].w, o[q].w, p[q].w, e1[q].w, e2[q].w);
    }
    // the main MMAs of the previous step must have finished reading the operand tiles
    if (s >= 1) tc::mbar_wait(&mbar_free, (uint32_t)((s - 1) & 1));
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      st_split4(sm.a1_hi, sm.a1_lo, poff[q], e1[q]);
      st_split4(sm.a2_hi, sm.a2_lo, poff[q], e2[q]);
    }
    if (tid < 128) st_split4(sm.b_hi, sm.b_lo, woff, w_cur);
    if (s + 1 < nsteps) p_to_smem(s + 1);  // its block's MMA was issued >= 2 steps ago
    // second step of a block: stage the W' rows of the next block (wp is free: the MMA of the
    // current block completed before its first rows were read) -- its MMA follows this barrier
    const bool next_block = (s & 3) == 1 && (s >> 2) + 1 < nblocks;
    if (next_block) stage_wp_block((s >> 2) + 1);
    w_cur = w_n1;
    w_n1 = w_n2;
    tc::fence_proxy_async();
    tc::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      if (next_block) issue_p((s >> 2) + 1);
#pragma unroll
      for (int j = 0; j < TWF_FR / 8; ++j) {
        const uint64_t dbh = s_desc[16 + 6 * j + 0], dbl = s_desc[16 + 6 * j + 1];
        const uint64_t d1h = s_desc[16 + 6 * j + 2], d1l = s_desc[16 + 6 * j + 3];
        const uint64_t d2h = s_desc[16 + 6 * j + 4], d2l = s_desc[16 + 6 * j + 5];
        const uint32_t acc = (s > 0 || j > 0) ? 1u : 0u;
        tc::mma_tf32(tmem, d1h, dbh, idesc, acc);
        tc::mma_tf32(tmem, d1h, dbl, idesc, 1u);
        tc::mma_tf32(tmem, d1l, dbh, idesc, 1u);
        tc::mma_tf32(tmem + 32, d2h, dbh, idesc, acc);
        tc::mma_tf32(tmem + 32, d2h, dbl, idesc, 1u);
        tc::mma_tf32(tmem + 32, d2l, dbh, idesc, 1u);
      }
      tc::mma_commit(&mbar_free);
      if (s == nsteps - 1) tc::mma_commit(&mbar_done);
    }
  }
  if (nsteps > 0) tc::mbar_wait(&mbar_done, 0);
  tc::fence_after_thread_sync();
  if (warp < 4) {
    const long no = nb + warp * 32 + lane;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t v[32];
      if (nsteps > 0) {
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(which * 32), v);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      double* out = which ? den : num;
      if (no < ldo) {
        const bool live = no < N;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < K) out[((size_t)split * K + k) * ldo + no] = live ? (double)__uint_as_float(v[k]) : 0.0;
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ---- the same kernel as two half-CTA pipelines (round 2) ------------------------------------------
// The 16 rows of a step belong to two groups of 4 warps: group h forms and stores the operands of
// rows 8h..8h+7, synchronises with a NAMED barrier over its 128 threads only, its first thread
// issues the MMAs of its 8 rows (K = 8) into the group's OWN accumulators and commits to the
// group's own completion barrier.  The groups drift apart, so the barrier / issue / MMA latency
// of one overlaps the operand work of the other (per-line stall profile of the single-pipeline
// kernel: 15.6 % block barrier, 13.4 % wait for the operand stage,
// profiles/r02/ncu_tw_contract_fused_tc_kernel.txt).  Shared between the groups: the H tile, the
// W' rows and the P' MMA of a 64-row block (staged by all threads, handed to thread 0 through an
// mbarrier with 256 arrivals, which also proves that every thread has read the P' buffer that the
// new block overwrites).  The two accumulator pairs are added in a fixed order in the write-out.
// 209 -> 191 us per launch (profiles/r02/ncu_tw_contract_fused2_tc_kernel.txt).  Not kept: warps
// that own 8 rows x their own 32 frames, transpose P' through a private tile and hand over through
// mbarriers only (no named barrier): 1.091 -> 1.114 ms for the spectral phase.
// (original notes:) thread = one frame x 8 rows straight out of
// TMEM (4-byte accesses) 229 us per launch; the same with fewer address instructions 240 us; with
// an mbarrier hand-over to a dedicated MMA warp 293 us; this one -- P' transposed through
// shared memory one step ahead so that the threads keep 16-byte accesses -- 220 us; that one
// with the MMA issue moved to a ninth warp 322 us (9 warps of 112 registers: three of them land
// on one scheduler's 16 K registers, so only ONE CTA fits per SM -- the same holds for every
// 288-thread variant above).
__global__ void __launch_bounds__(TWF_THREADS, 2)
tw_contract_fused2_tc_kernel(const float* __restrict__ hatW, const float* __restrict__ Op, long ld,
                            const float* __restrict__ W, int ldw, const float* __restrict__ H,
                            long ldh, int K, int F, long N, int fchunk, int fsplit,
                            double* __restrict__ num, double* __restrict__ den, long ldo) {
  extern __shared__ __align__(1024) unsigned char twf_smem[];
  __shared__ uint64_t s_desc[28];
  __shared__ uint64_t mbar_free[2];  // per group: the MMAs of its previous step have read its tiles
  __shared__ uint64_t mbar_p[2];
  __shared__ uint64_t mbar_wp;       // 256 arrivals: the W' rows of the next P' block are staged
  __shared__ uint64_t mbar_done;     // 2 arrivals: the last MMAs of both groups
  __shared__ uint32_t tmem_base;
  unsigned char* base = twf_smem + ((1024 - (tc::smem_u32(twf_smem) & 1023)) & 1023);
  TwfSmem& sm = *reinterpret_cast<TwfSmem*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long nb = (long)blockIdx.x * TWF_NT;
  const int split = blockIdx.y;
  const int fb = split * fchunk;
  int fe = fb + fchunk;
  if (fe > F) fe = F;
  const int nsteps = (fe - fb + TWF_FR - 1) / TWF_FR;

  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  if (tid == 0) {
    tc::mbar_init(&mbar_free[0], 1);
    tc::mbar_init(&mbar_free[1], 1);
    tc::mbar_init(&mbar_p[0], 1);
    tc::mbar_init(&mbar_p[1], 1);
    tc::mbar_init(&mbar_wp, TWF_THREADS);
    tc::mbar_init(&mbar_done, 2);
    tc::fence_mbar_init();
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // H tile 32 k x 128 n (zero beyond K and beyond the row): 4 float4 per thread
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = tid + q * TWF_THREADS;  // 0 .. 1023
    const int k = i >> 5, c = (i & 31) * 4;
    const long n = nb + c;
    const float4 h = (k < K && n + 4 <= ldh) ? ldg4(H + (long)k * ldh + n) : zero4;
    st_split4(sm.h_hi, sm.h_lo, tc::mnmajor_off(k, c, TWF_HLBO, TWF_SBO), h);
  }
  // W rows: the first 64 threads of a group hold 4 consecutive k of one of its 8 rows of a step
  const int grp = warp >> 2, gt = tid & 127;
  const int wrow = 8 * grp + (gt >> 3), wk = (tid & 7) * 4;
  const uint32_t woff = tc::mnmajor_off(wrow, wk, TWF_LBO, TWF_SBO);   // main B (f = K index)
  auto load_w = [&](int step) {
    float4 w = zero4;
    if (gt < 64 && step < nsteps) {
      const int f = fb + step * TWF_FR + wrow;
      if (f < fe) {
        const float* wr = W + (long)f * ldw;
        w.x = (wk + 0 < K) ? __ldg(wr + wk + 0) : 0.f;
        w.y = (wk + 1 < K) ? __ldg(wr + wk + 1) : 0.f;
        w.z = (wk + 2 < K) ? __ldg(wr + wk + 2) : 0.f;
        w.w = (wk + 3 < K) ? __ldg(wr + wk + 3) : 0.f;
      }
    }
    return w;
  };
  // W' rows of one P' block (64 rows x 32 k = 512 float4: two per thread), K-major tile
  auto stage_wp_block = [&](int blk) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = tid + q * TWF_THREADS;  // 0 .. 511
      const int r = i >> 3, kk = (i & 7) * 4;
      const int f = fb + blk * TWF_PB + r;
      float4 w = zero4;
      if (f < fe) {
        const float* wr = W + (long)f * ldw;
        w.x = (kk + 0 < K) ? __ldg(wr + kk + 0) : 0.f;
        w.y = (kk + 1 < K) ? __ldg(wr + kk + 1) : 0.f;
        w.z = (kk + 2 < K) ? __ldg(wr + kk + 2) : 0.f;
        w.w = (kk + 3 < K) ? __ldg(wr + kk + 3) : 0.f;
      }
      float4 hi, lo;
      tc::split_tf32(w.x, hi.x, lo.x); tc::split_tf32(w.y, hi.y, lo.y);
      tc::split_tf32(w.z, hi.z, lo.z); tc::split_tf32(w.w, hi.w, lo.w);
      const uint32_t off = tc::kmajor_off(r, kk);
      *reinterpret_cast<float4*>(sm.wp_hi + off) = hi;
      *reinterpret_cast<float4*>(sm.wp_lo + off) = lo;
    }
  };
  // plane elements of this thread: float4 (f, n4) of the 16 x 128 tiles, 2 per plane and step
  int prow[2], pcol[2];
  uint32_t poff[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {  // group h: rows 8h .. 8h+7, two per warp
    prow[q] = 8 * grp + 2 * (warp & 3) + q;
    pcol[q] = lane * 4;
    poff[q] = tc::mnmajor_off(prow[q], pcol[q], TWF_LBO, TWF_SBO);
  }
  float4 hw_n[2], o_n[2];
  auto fetch = [&](int step) {
    const int f0 = fb + step * TWF_FR;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f = f0 + prow[q];
      const long n = nb + pcol[q];
      const bool ok = (f < fe) && (n + 4 <= ld);
      const long off = (long)f * ld + n;
      hw_n[q] = ok ? ldg4(hatW + off) : zero4;
      o_n[q] = ok ? ldg4(Op + off) : zero4;
    }
  };
  // TMEM -> ptile: this thread's TMEM lane is frame nl, it moves rows fhalf * 8 .. + 7
  const int nl = (warp & 3) * 32 + lane, fhalf = warp >> 2;
  const uint32_t tmem_lane = (uint32_t)((warp & 3) * 32) << 16;

  const int nblocks = (nsteps + 3) / 4;
  float4 w_cur = load_w(0), w_n1 = load_w(1);
  stage_wp_block(0);
  tc::fence_proxy_async();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = tc::idesc_tf32(128, 32, 1, 1);
  const uint32_t idesc_p = tc::idesc_tf32(128, TWF_PB, 1, 0);
  // every shared-memory descriptor is loop invariant (one operand stage): thread 0 builds them
  // once; issuing a step is then 28 shared loads + 36 MMAs instead of ~400 ALU instructions on
  // the critical path of warp 0
  if (tid == 0) {
    const uint32_t hh = tc::smem_u32(sm.h_hi), hl = tc::smem_u32(sm.h_lo);
    const uint32_t wh = tc::smem_u32(sm.wp_hi), wl = tc::smem_u32(sm.wp_lo);
    for (int j = 0; j < 4; ++j) {
      s_desc[j] = tc::smem_desc_mnmajor(hh + j * 1024, TWF_HLBO, TWF_SBO);
      s_desc[4 + j] = tc::smem_desc_mnmajor(hl + j * 1024, TWF_HLBO, TWF_SBO);
      s_desc[8 + j] = tc::smem_desc_kmajor(wh + j * 32);
      s_desc[12 + j] = tc::smem_desc_kmajor(wl + j * 32);
    }
    const uint32_t a1h = tc::smem_u32(sm.a1_hi), a1l = tc::smem_u32(sm.a1_lo);
    const uint32_t a2h = tc::smem_u32(sm.a2_hi), a2l = tc::smem_u32(sm.a2_lo);
    const uint32_t bh = tc::smem_u32(sm.b_hi), bl = tc::smem_u32(sm.b_lo);
    for (int j = 0; j < TWF_FR / 8; ++j) {
      s_desc[16 + 6 * j + 0] = tc::smem_desc_mnmajor(bh + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 1] = tc::smem_desc_mnmajor(bl + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 2] = tc::smem_desc_mnmajor(a1h + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 3] = tc::smem_desc_mnmajor(a1l + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 4] = tc::smem_desc_mnmajor(a2h + j * 1024, TWF_LBO, TWF_SBO);
      s_desc[16 + 6 * j + 5] = tc::smem_desc_mnmajor(a2l + j * 1024, TWF_LBO, TWF_SBO);
    }
  }
  // P'(block) -> TMEM columns 64 + 64 (block & 1) .. + 63; W' rows are in wp_hi / wp_lo
  auto issue_p = [&](int blk) {
    const int step = blk;  // (mbarrier slot = block parity)
    const uint32_t d = tmem + 64u + (uint32_t)TWF_PB * (uint32_t)(blk & 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t dah = s_desc[j], dal = s_desc[4 + j], dbh = s_desc[8 + j], dbl = s_desc[12 + j];
      tc::mma_tf32(d, dah, dbh, idesc_p, j > 0 ? 1u : 0u);
      tc::mma_tf32(d, dah, dbl, idesc_p, 1u);
      tc::mma_tf32(d, dal, dbh, idesc_p, 1u);
    }
    tc::mma_commit(&mbar_p[step & 1]);
  };
  // P' of a step: wait for the MMA of its block, move this thread's 8 values from TMEM to
  // ptile[step & 1]
  auto p_to_smem = [&](int step) {
    const int blk = step >> 2, bb = blk & 1;
    tc::mbar_wait(&mbar_p[bb], (uint32_t)((blk >> 1) & 1));
    tc::fence_after_thread_sync();
    uint32_t pv[8];
    tc::tmem_ld_32x8(tmem + tmem_lane + 64u + (uint32_t)TWF_PB * (uint32_t)bb +
                         16u * (uint32_t)(step & 3) + 8u * (uint32_t)fhalf, pv);
    tc::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) sm.ptile[step & 1][fhalf * 8 + q][nl] = __uint_as_float(pv[q]);
  };
  __syncthreads();  // the descriptors built by thread 0 are read by both issuing threads
  const bool issuer = gt == 0;           // threads 0 and 128
  const uint32_t dacc = tmem + 192u * (uint32_t)grp;  // the group's accumulators: num, den (+32)
  if (nsteps > 0) {
    fetch(0);
    if (tid == 0) issue_p(0);
    p_to_smem(0);                       // (also: the P' MMA of block 0 has finished reading wp)
    tc::fence_before_thread_sync();
    tc::named_bar_sync(1 + grp, 128);   // the group's rows of ptile[0] are complete
  }
  for (int s = 0; s < nsteps; ++s) {
    float4 hw[2], o[2], p[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) { hw[q] = hw_n[q]; o[q] = o_n[q]; }
    if (s + 1 < nsteps) fetch(s + 1);
    const float4 w_n2 = load_w(s + 2);
#pragma unroll
    for (int q = 0; q < 2; ++q)
      p[q] = *reinterpret_cast<const float4*>(&sm.ptile[s & 1][prow[q]][pcol[q]]);
    float4 e1[2], e2[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      // rows beyond the shard (all-zero loads) contribute nothing: W is zero there as well
      auto e12 = [&](float hwv, float ov, float pvv, float& a, float& c) {
        const float oc = fmaxf(ov, kEpsF);
        const float rp = fast_rcpf(fmaxf(pvv, kEpsF));
        c = oc * rp;                // other / P'            (audioModel.py:1694-1701)
        a = oc * (hwv * rp * rp);   // other * hat_W / P'^2  (audioModel.py:1714-1720)
      };
      e12(hw[q].x, o[q].x, p[q].x, e1[q].x, e2[q].x);
      e12(hw[q].y, o[q].y, p[q].y, e1[q].y, e2[q].y);
      e12(hw[q].z, o[q].z, p[q].z, e1[q].z, e2[q].z);
      e12(hw[q].w, o[q].w, p[q].w, e1[q].w, e2[q].w);
    }
    // the group's MMAs of the previous step must have finished reading its operand rows
    if (s >= 1) tc::mbar_wait(&mbar_free[grp], (uint32_t)((s - 1) & 1));
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      st_split4(sm.a1_hi, sm.a1_lo, poff[q], e1[q]);
      st_split4(sm.a2_hi, sm.a2_lo, poff[q], e2[q]);
    }
    if (gt < 64) st_split4(sm.b_hi, sm.b_lo, woff, w_cur);
    if (s + 1 < nsteps) p_to_smem(s + 1);  // its block's MMA was issued >= 2 steps ago
    // first step of a block: stage the W' rows of the next block (wp is free: the MMA of the
    // current block completed before its first rows were read, one step ago); thread 0 issues its
    // MMA once all 256 threads have arrived -- which also means that nobody reads the P' buffer it
    // overwrites any more (last read two steps ago)
    const bool next_block = (s & 3) == 0 && (s >> 2) + 1 < nblocks;
    if (next_block) stage_wp_block((s >> 2) + 1);
    w_cur = w_n1;
    w_n1 = w_n2;
    tc::fence_proxy_async();
    tc::fence_before_thread_sync();
    if (next_block) tc::mbar_arrive(&mbar_wp);
    tc::named_bar_sync(1 + grp, 128);
    if (issuer) {
      tc::fence_after_thread_sync();
      if (grp == 0 && next_block) {
        tc::mbar_wait(&mbar_wp, (uint32_t)((s >> 2) & 1));
        tc::fence_after_thread_sync();
        issue_p((s >> 2) + 1);
      }
      const uint64_t dbh = s_desc[16 + 6 * grp + 0], dbl = s_desc[16 + 6 * grp + 1];
      const uint64_t d1h = s_desc[16 + 6 * grp + 2], d1l = s_desc[16 + 6 * grp + 3];
      const uint64_t d2h = s_desc[16 + 6 * grp + 4], d2l = s_desc[16 + 6 * grp + 5];
      const uint32_t acc = s > 0 ? 1u : 0u;
      tc::mma_tf32(dacc, d1h, dbh, idesc, acc);
      tc::mma_tf32(dacc, d1h, dbl, idesc, 1u);
      tc::mma_tf32(dacc, d1l, dbh, idesc, 1u);
      tc::mma_tf32(dacc + 32, d2h, dbh, idesc, acc);
      tc::mma_tf32(dacc + 32, d2h, dbl, idesc, 1u);
      tc::mma_tf32(dacc + 32, d2l, dbh, idesc, 1u);
      tc::mma_commit(&mbar_free[grp]);
      if (s == nsteps - 1) tc::mma_commit(&mbar_done);
    }
  }
  if (nsteps > 0) tc::mbar_wait(&mbar_done, 0);
  tc::fence_after_thread_sync();
  if (warp < 4) {
    const long no = nb + warp * 32 + lane;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      double* out = which ? den : num;
#pragma unroll 1
      for (int k0 = 0; k0 < 32; k0 += 8) {  // 8 columns of both accumulators at a time
        uint32_t v[8], v2[8];
        if (nsteps > 0) {
          tc::tmem_ld_32x8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(which * 32 + k0), v);
          tc::tmem_ld_32x8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(192 + which * 32 + k0), v2);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = v2[i] = 0u;
        }
        if (no < ldo) {
          const bool live = no < N;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // rows 0-7 of every step, then rows 8-15: a fixed order
            if (k0 + k < K)
              out[((size_t)split * K + k0 + k) * ldo + no] =
                  live ? (double)__uint_as_float(v[k]) + (double)__uint_as_float(v2[k]) : 0.0;
        }
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace pf

using namespace pf;

// TW update contractions for float32 planes, K <= 32, P' = W' H formed in the kernel
int pf_tw_contract_fused_tc(const float* hatW, const float* O, long ld, const float* W, int ldw,
                            const float* H, long ldh, int K, int F, long N, int fchunk, int fsplit,
                            double* num, double* den, long ldo, cudaStream_t st) {
  const size_t smem = sizeof(TwfSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(tw_contract_fused_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("tw_contract_fused_tc_kernel: %zu bytes of shared memory: %s", smem,
              cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  dim3 grid(ceil_div(N, TWF_NT), fsplit);
  // PYFASST_TW_HALVES=0: the single-pipeline kernel (one block barrier per step)
  static const bool halves = [] {
    const char* e2 = getenv("PYFASST_TW_HALVES");
    return e2 == nullptr || atoi(e2) != 0;
  }();
  if (halves) {
    e = cudaFuncSetAttribute(tw_contract_fused2_tc_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("tw_contract_fused2_tc_kernel: %s", cudaGetErrorString(e));
      return PF_ERR_CUDA;
    }
    tw_contract_fused2_tc_kernel<<<grid, TWF_THREADS, smem, st>>>(hatW, O, ld, W, ldw, H, ldh, K, F,
                                                                N, fchunk, fsplit, num, den, ldo);
    return check_launch("tw_contract_fused2_tc_kernel");
  }
  tw_contract_fused_tc_kernel<<<grid, TWF_THREADS, smem, st>>>(hatW, O, ld, W, ldw, H, ldh, K, F, N,
                                                             fchunk, fsplit, num, den, ldo);
  return check_launch("tw_contract_fused_tc_kernel");
}

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
  // PYFASST_FBT_INTERLEAVE=0: one contiguous run of frames per split (the older mapping)
  const char* env = getenv("PYFASST_FBT_INTERLEAVE");
  const bool interleave = env == nullptr || atoi(env) != 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    fb_contract_tc_kernel<<<grid, FBT_THREADS, smem, st>>>(hatW, P, ld, G, ldg, k0, K, F, N,
                                                         interleave ? -chunk : chunk, nsplit, num);
    int rc = check_launch("fb_contract_tc_kernel");
    if (rc) return rc;
  }
  return PF_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*pf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                       const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static pf_encode_tiled_fn tensor_map_encoder() {
  static pf_encode_tiled_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (pf_encode_tiled_fn)p;
  }();
  return fn;
}

// V = W H for float32 planes, K <= 32 (called from pf_spec_power)
int pf_spec_power_tc(const float* W, int ldw, const float* H, long ldh, float* V, long ldv, int F,
                     int K, long N, cudaStream_t st) {
  // tensor map of the output plane [F rows][ldv columns] float32, boxes of 32 x 32, 128-byte swizzle
  pf_encode_tiled_fn encode = tensor_map_encoder();
  if (encode == nullptr) {
    set_error("spec_power_tc_kernel: cuTensorMapEncodeTiled is not available from this driver");
    return PF_ERR_CUDA;
  }
  CUtensorMap tmapV;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)ldv, (cuuint64_t)F};
    const cuuint64_t strides[1] = {(cuuint64_t)ldv * sizeof(float)};
    const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    const CUresult r = encode(&tmapV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)V, dims, strides,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("spec_power_tc_kernel: cuTensorMapEncodeTiled failed (%d) for V=%p ldv=%ld F=%d",
                (int)r, (void*)V, ldv, F);
      return PF_ERR_CUDA;
    }
  }
  const size_t smem = sizeof(SptSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(spec_power_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("spec_power_tc_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  const long ntiles = (ldv + SPT_NT - 1) / SPT_NT;
  const int fblocks = ceil_div(F, 128);
  // one CTA per SM (it owns all 512 TMEM columns): split the frame tiles over ~148 CTAs
  long splits = 148 / fblocks;  // never more CTAs than SMs: a second wave would double the time
  if (splits > ntiles) splits = ntiles;
  if (splits < 1) splits = 1;
  const int per = (int)((ntiles + splits - 1) / splits);
  dim3 grid(ceil_div(ntiles, per), fblocks);
  // PYFASST_SPT_INTERLEAVE=0: contiguous runs of tiles per CTA (the older mapping)
  const char* env = getenv("PYFASST_SPT_INTERLEAVE");
  const bool interleave = env == nullptr || atoi(env) != 0;
  spec_power_tc_kernel<<<grid, SPT_THREADS, smem, st>>>(W, ldw, H, ldh, tmapV, ldv, F, K, N,
                                                        interleave ? -per : per);
  return check_launch("spec_power_tc_kernel");
}

// TW update contractions for float32 planes, K <= 32; Pn = W' H (from pf_spec_power_tc)
int pf_tw_contract_tc(const float* hatW, const float* O, const float* Pn, long ld, const float* W,
                      int ldw, int K, int F, long N, int fchunk, int fsplit, double* num,
                      double* den, long ldo, cudaStream_t st) {
  const size_t smem = 2 * sizeof(TwtStage) + 1024;
  cudaError_t e = cudaFuncSetAttribute(tw_contract_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("tw_contract_tc_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  dim3 grid(ceil_div(N, TWT_NT), fsplit);
  tw_contract_tc_kernel<<<grid, TWT_THREADS, smem, st>>>(hatW, O, Pn, ld, W, ldw, K, F, N, fchunk,
                                                       fsplit, num, den, ldo);
  return check_launch("tw_contract_tc_kernel");
}
