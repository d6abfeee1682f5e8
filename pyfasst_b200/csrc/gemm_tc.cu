// General float32 GEMM on the 5th-generation tensor cores (tcgen05 kind::tf32 with the
// 3xTF32 split, TMEM accumulators): C[M x N] = op(A) op(B), row-major operands.
//
// This is the dense workhorse of the SIMM source/filter model (pyfasst/SeparateLeadStereo/
// SIMM/SIMM.py: every np.dot of the update loops :303-393 and :613-941 -- contractions with
// the NF0 ~ 480 column F0 dictionary WF0 are large enough to be tensor-pipe work).  x = hi + lo
// with hi = tf32(x): A B ~= Ah Bh + Ah Bl + Al Bh keeps float32-class accuracy (~2^-21 per
// product), which the parity tolerance (1e-4 on the factors) needs and plain tf32 (2^-11) does
// not give.
//
// The hi / lo operand tiles cannot come from TMA (they are computed from the loaded values):
// the 256 threads of a CTA load a K chunk of 32, split it and store it into shared memory in the
// canonical UMMA layouts of tc.cuh -- K-major for an operand that is contiguous along K,
// MN-major (SWIZZLE_128B_BASE32B) for one contiguous along M / N -- while one elected thread
// issues the MMAs of the previous chunk; ring slots are recycled through mbarriers
// (tcgen05.commit).  One CTA owns a 128 x BN tile of C.
#include "common.cuh"
#include "tc.cuh"

namespace pf {

constexpr int GT_THREADS = 256;
constexpr int GT_BM = 128;
constexpr int GT_BK = 32;

__device__ __forceinline__ float4 gt_ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void gt_st_split4(unsigned char* hi_tile, unsigned char* lo_tile,
                                             uint32_t off, float4 x) {
  float4 h, l;
  tc::split_tf32(x.x, h.x, l.x);
  tc::split_tf32(x.y, h.y, l.y);
  tc::split_tf32(x.z, h.z, l.z);
  tc::split_tf32(x.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// One operand tile of ROWS (M or N extent) x 32 k.
//  KMAJOR : global element (r, k) at P[(r0 + r) * ld + k0 + k]   (contiguous along k)
//  !KMAJOR: global element (r, k) at P[(k0 + k) * ld + r0 + r]   (contiguous along r)
template <int ROWS, bool KMAJOR>
struct OperandTile {
  static constexpr int NV = ROWS * GT_BK / 4 / GT_THREADS;  // float4 per thread
  static constexpr uint32_t LBO = GT_BK * 128, SBO = 512;   // MN-major: k rows stacked per group
  float4 v[NV];

  __device__ __forceinline__ void fetch(const float* __restrict__ P, long ld, long r0, long rmax,
                                        long k0, long kmax, int tid) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int i = tid + q * GT_THREADS;
      if (KMAJOR) {
        const int r = i >> 3, c = i & 7;
        const long row = r0 + r, k = k0 + c * 4;
        v[q] = (row < rmax && k + 4 <= kmax) ? gt_ldg4(P + row * ld + k) : zero4;
      } else {
        constexpr int VPR = ROWS / 4;  // float4 per k row
        const int k = i / VPR, c = i % VPR;
        const long kk = k0 + k, col = r0 + c * 4;
        v[q] = (kk < kmax && col + 4 <= ld) ? gt_ldg4(P + kk * ld + col) : zero4;
      }
    }
  }
  __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo, int tid) const {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int i = tid + q * GT_THREADS;
      uint32_t off;
      if (KMAJOR) {
        const int r = i >> 3, c = i & 7;
        off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
      } else {
        constexpr int VPR = ROWS / 4;
        const int k = i / VPR, c = i % VPR;
        off = tc::mnmajor_off(k, c * 4, LBO, SBO);
      }
      gt_st_split4(hi, lo, off, v[q]);
    }
  }
  // descriptor of K step j (8 of the 32 k) of a tile at shared address `saddr`
  __device__ __forceinline__ static uint64_t desc(uint32_t saddr, int j) {
    return KMAJOR ? tc::smem_desc_kmajor(saddr + j * 32)
                  : tc::smem_desc_mnmajor(saddr + j * 1024, LBO, SBO);
  }
};

template <int BN>
struct GtStage {
  unsigned char a_hi[GT_BM * GT_BK * 4];
  unsigned char a_lo[GT_BM * GT_BK * 4];
  unsigned char b_hi[BN * GT_BK * 4];
  unsigned char b_lo[BN * GT_BK * 4];
};

template <int BN, int STAGES>
struct GtSmem {
  GtStage<BN> stage[STAGES];  // the write-out reuses stage[0] for its per-warp transpose tiles
};
// CTAs per SM: two when the operand ring fits twice in shared memory (one-stage ring of the
// short-K products, or the narrow BN = 64 tile) -- one CTA's write-out then overlaps the
// other's loads and MMAs; TMEM: 2 x BN <= 512 columns.
template <int BN, int STAGES>
struct GtOcc {
  static constexpr int value = (sizeof(GtSmem<BN, STAGES>) + 1024 <= 110 * 1024) ? 2 : 1;
};
static_assert(sizeof(GtStage<64>) >= (GT_THREADS / 32) * 32 * 36 * sizeof(float), "transpose tiles");

// TA: A is stored [K][M] (A^T given); TB: B is stored [N][K] (B^T given)
template <bool TA, bool TB, int BN, int STAGES>
__global__ void __launch_bounds__(GT_THREADS, GtOcc<BN, STAGES>::value)
gemm_tf32x3_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                   float* __restrict__ C, long ldc, int M, int N, int K, int kper, long cstride) {
  extern __shared__ __align__(1024) unsigned char gt_smem[];
  __shared__ uint64_t mbar_free[2];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base;
  unsigned char* base = gt_smem + ((1024 - (tc::smem_u32(gt_smem) & 1023)) & 1023);
  GtSmem<BN, STAGES>& sm = *reinterpret_cast<GtSmem<BN, STAGES>*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long m0 = (long)blockIdx.y * GT_BM, n0 = (long)blockIdx.x * BN;
  // split K: CTA z contracts k in [kbeg, kend) into the partial product C + z * cstride
  const long kbeg = (long)blockIdx.z * kper;
  const long kend = kbeg + kper < (long)K ? kbeg + kper : (long)K;
  C += (long)blockIdx.z * cstride;
  const int nchunks = kend > kbeg ? (int)((kend - kbeg + GT_BK - 1) / GT_BK) : 0;
  constexpr uint32_t TCOLS = BN < 32 ? 32 : BN;

  if (warp == 0) tc::tmem_alloc(&tmem_base, TCOLS);
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
  const uint32_t idesc = tc::idesc_tf32(GT_BM, BN, TA ? 1 : 0, TB ? 0 : 1);

  OperandTile<GT_BM, !TA> ta;  // A is K-major unless its transpose is what is stored
  OperandTile<BN, TB> tb;      // B is K-major when B^T ([N][K]) is what is stored
  ta.fetch(A, lda, m0, M, kbeg, kend, tid);
  tb.fetch(B, ldb, n0, N, kbeg, kend, tid);
  for (int s = 0; s < nchunks; ++s) {
    const int b = s % STAGES;
    GtStage<BN>& st = sm.stage[b];
    if (s >= STAGES) tc::mbar_wait(&mbar_free[b], (uint32_t)((s / STAGES - 1) & 1));
    ta.store(st.a_hi, st.a_lo, tid);
    tb.store(st.b_hi, st.b_lo, tid);
    if (s + 1 < nchunks) {
      ta.fetch(A, lda, m0, M, kbeg + (long)(s + 1) * GT_BK, kend, tid);
      tb.fetch(B, ldb, n0, N, kbeg + (long)(s + 1) * GT_BK, kend, tid);
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      const uint32_t ah = tc::smem_u32(st.a_hi), al = tc::smem_u32(st.a_lo);
      const uint32_t bh = tc::smem_u32(st.b_hi), bl = tc::smem_u32(st.b_lo);
#pragma unroll
      for (int j = 0; j < GT_BK / 8; ++j) {
        const uint64_t dah = OperandTile<GT_BM, !TA>::desc(ah, j), dal = OperandTile<GT_BM, !TA>::desc(al, j);
        const uint64_t dbh = OperandTile<BN, TB>::desc(bh, j), dbl = OperandTile<BN, TB>::desc(bl, j);
        tc::mma_tf32(tmem, dah, dbh, idesc, (s > 0 || j > 0) ? 1u : 0u);
        tc::mma_tf32(tmem, dah, dbl, idesc, 1u);
        tc::mma_tf32(tmem, dal, dbh, idesc, 1u);
      }
      tc::mma_commit(&mbar_free[b]);
      if (s == nchunks - 1) tc::mma_commit(&mbar_done);
    }
  }
  if (nchunks > 0) tc::mbar_wait(&mbar_done, 0);
  tc::fence_after_thread_sync();

  // write-out: warp w reads TMEM lanes 32 (w % 4) .. +31 (rows of C) and a share of the
  // columns, transposes 32 x 32 blocks through shared memory and stores 128-byte row segments
  // (all MMAs have completed: the operand ring is free)
  float* tr = reinterpret_cast<float*>(&sm.stage[0]) + warp * (32 * 36);
  const int wq = warp & 3, half = warp >> 2;
  constexpr int CPW = BN / 2 < 32 ? 32 : BN / 2;  // columns per warp group
  if (half * CPW < BN) {
#pragma unroll 1
    for (int c0 = 0; c0 < CPW; c0 += 32) {
      uint32_t v[32];
      if (nchunks > 0) {
        tc::tmem_ld_32x32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * CPW + c0), v);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(tr + lane * 36 + i) =
            make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                        __uint_as_float(v[i + 3]));
      __syncwarp();
      const long col = n0 + half * CPW + c0 + (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const long row = m0 + wq * 32 + r;
        const float4 x = *reinterpret_cast<const float4*>(tr + r * 36 + (lane & 7) * 4);
        if (row < M) {
          float* out = C + row * ldc + col;
          if (col + 4 <= N) {
            *reinterpret_cast<float4*>(out) = x;
          } else {
            if (col + 0 < N) out[0] = x.x;
            if (col + 1 < N) out[1] = x.y;
            if (col + 2 < N) out[2] = x.z;
          }
        }
      }
      __syncwarp();
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

template <bool TA, bool TB, int BN, int STAGES>
static int launch_gemm_stages(const float* A, long lda, const float* B, long ldb, float* C,
                              long ldc, int M, int N, int K, int ksplit, int kper, long cstride,
                              cudaStream_t st) {
  const size_t smem = sizeof(GtSmem<BN, STAGES>) + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<TA, TB, BN, STAGES>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("gemm_tf32x3_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  dim3 grid(ceil_div(N, BN), ceil_div(M, GT_BM), ksplit);
  gemm_tf32x3_kernel<TA, TB, BN, STAGES><<<grid, GT_THREADS, smem, st>>>(A, lda, B, ldb, C, ldc, M,
                                                                         N, K, kper, cstride);
  return check_launch("gemm_tf32x3_kernel");
}

template <bool TA, bool TB, int BN>
static int launch_gemm(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                       int M, int N, int K, int ksplit, int kper, long cstride, cudaStream_t st) {
  // short contractions (at most two K chunks per CTA: the K = R products of the SIMM
  // accompaniment model) take the one-stage ring and two CTAs per SM
  const long kspan = ksplit > 1 ? kper : K;
  if (kspan <= 2 * GT_BK)
    return launch_gemm_stages<TA, TB, BN, 1>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride,
                                             st);
  return launch_gemm_stages<TA, TB, BN, 2>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride,
                                           st);
}

template <bool TA, bool TB>
static int dispatch_bn(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                       int M, int N, int K, int ksplit, int kper, long cstride, cudaStream_t st) {
  if (N > 128)
    return launch_gemm<TA, TB, 256>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
  if (N > 64)
    return launch_gemm<TA, TB, 128>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
  return launch_gemm<TA, TB, 64>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
}

static int dispatch_gemm(const float* A, long lda, int transA, const float* B, long ldb, int transB,
                         float* C, long ldc, int M, int N, int K, int ksplit, int kper,
                         long cstride, cudaStream_t st) {
  if (transA) {
    if (transB)
      return dispatch_bn<true, true>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
    return dispatch_bn<true, false>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
  }
  if (transB)
    return dispatch_bn<false, true>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
  return dispatch_bn<false, false>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper, cstride, st);
}

// out[i] = sum_z part[z][i] in a fixed order (float4 per thread, double accumulation)
// (the GEMM writes N of the ldw columns of a partial product: the padding is masked to zero)
__global__ void gemm_splitk_reduce_kernel(const float* __restrict__ part, int ksplit, long count4,
                                          long stride, int N, int ldw, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int z = 0; z < ksplit; ++z) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)z * stride) + i);
    a += v.x; b += v.y; c += v.z; d += v.w;
  }
  const int col = (int)((i * 4) % ldw);
  reinterpret_cast<float4*>(out)[i] =
      make_float4(col + 0 < N ? (float)a : 0.f, col + 1 < N ? (float)b : 0.f,
                  col + 2 < N ? (float)c : 0.f, col + 3 < N ? (float)d : 0.f);
}

static int check_gemm_args(const char* who, const float* A, long lda, int transA, const float* B,
                           long ldb, int transB, const float* C, long ldc, int M, int N, int K) {
  if (!(M > 0 && N > 0 && K > 0)) {
    set_error("%s: empty problem %d x %d x %d", who, M, N, K);
    return PF_ERR_ARG;
  }
  if (lda % 4 || ldb % 4 || ldc % 4) {
    set_error("%s: leading dimensions must be multiples of 4 (lda=%ld ldb=%ld ldc=%ld)", who, lda,
              ldb, ldc);
    return PF_ERR_ARG;
  }
  if ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) != 0) {
    set_error("%s: operands must be 16-byte aligned", who);
    return PF_ERR_ARG;
  }
  if (!(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N)) {
    set_error("%s: leading dimension smaller than the row length", who);
    return PF_ERR_ARG;
  }
  // an operand that is contiguous along K is read in float4 along K
  if (!((transA && !transB) || K % 4 == 0)) {
    set_error("%s: K=%d must be a multiple of 4 (zero padded) unless both operands are "
              "contiguous along M / N", who, K);
    return PF_ERR_ARG;
  }
  return PF_OK;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_gemm_tf32x3(const float* A, int64_t lda, int transA, const float* B, int64_t ldb,
                              int transB, float* C, int64_t ldc, int M, int N, int K,
                              void* stream) {
  int rc = check_gemm_args("pf_gemm_tf32x3", A, lda, transA, B, ldb, transB, C, ldc, M, N, K);
  if (rc) return rc;
  return dispatch_gemm(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, 1, K, 0, as_stream(stream));
}

extern "C" int pf_gemm_splitk_plan(int M, int N, int K, int* ksplit, int64_t* workspace_bytes) {
  PF_REQUIRE(M > 0 && N > 0 && K > 0, "pf_gemm_splitk_plan: empty problem %d x %d x %d", M, N, K);
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  const long tiles = (long)ceil_div(N, bn) * ceil_div(M, GT_BM);
  // ONE wave of co-resident CTAs (two per SM on a 148-SM part: 296 slots), never a CTA more --
  // rounding the split count UP (33 splits x 9 row tiles = 297 CTAs for the F x R products of the
  // SIMM model) left one CTA to run alone after the wave --, but at least 8 K chunks per CTA
  long slots = 2L * 148;
  if (const char* e = getenv("PYFASST_GEMM_SPLIT_SLOTS")) {  // tuning override
    const long v = atol(e);
    if (v > 0) slots = v;
  }
  long want = slots / tiles;
  const long most = (K + 8L * GT_BK - 1) / (8L * GT_BK);
  if (want > most) want = most;
  if (want < 1) want = 1;
  *ksplit = (int)want;
  const long ldc = (N + 3) / 4 * 4;
  *workspace_bytes = want > 1 ? (int64_t)want * M * ldc * sizeof(float) : 0;
  return PF_OK;
}

extern "C" int pf_gemm_tf32x3_splitk(const float* A, int64_t lda, int transA, const float* B,
                                     int64_t ldb, int transB, float* C, int64_t ldc, int M, int N,
                                     int K, float* workspace, int64_t workspace_bytes,
                                     void* stream) {
  int rc = check_gemm_args("pf_gemm_tf32x3_splitk", A, lda, transA, B, ldb, transB, C, ldc, M, N, K);
  if (rc) return rc;
  int ksplit;
  int64_t need;
  pf_gemm_splitk_plan(M, N, K, &ksplit, &need);
  cudaStream_t st = as_stream(stream);
  if (ksplit == 1)
    return dispatch_gemm(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, 1, K, 0, st);
  PF_REQUIRE(workspace != nullptr && workspace_bytes >= need && (((uintptr_t)workspace) & 15) == 0,
             "pf_gemm_tf32x3_splitk: workspace of %ld bytes needed (16-byte aligned), got %ld",
             (long)need, (long)workspace_bytes);
  const long ldw = (N + 3) / 4 * 4;
  long kper = ((long)K + ksplit - 1) / ksplit;
  kper = (kper + GT_BK - 1) / GT_BK * GT_BK;
  ksplit = (int)((K + kper - 1) / kper);
  const long stride = (long)M * ldw;
  rc = dispatch_gemm(A, lda, transA, B, ldb, transB, workspace, ldw, M, N, K, ksplit, (int)kper,
                     stride, st);
  if (rc) return rc;
  if (ldc == ldw) {
    const long count4 = stride / 4;
    gemm_splitk_reduce_kernel<<<ceil_div(count4, 256), 256, 0, st>>>(workspace, ksplit, count4,
                                                                     stride, N, (int)ldw, C);
    return check_launch("gemm_splitk_reduce_kernel");
  }
  set_error("pf_gemm_tf32x3_splitk: ldc=%ld must equal N rounded up to 4 (%ld)", (long)ldc, ldw);
  return PF_ERR_ARG;
}
