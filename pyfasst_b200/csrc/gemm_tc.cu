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
// threads load a K chunk of 32, split it and store it into shared memory in the canonical UMMA
// layouts of tc.cuh -- K-major for an operand that is contiguous along K, MN-major
// (SWIZZLE_128B_BASE32B) for one contiguous along M / N; ring slots are recycled through mbarriers
// (tcgen05.commit).  One CTA owns a 128 x BN tile of C.  Two kernels:
//   gemm_tf32x3_ws_kernel  long contractions (more than two K chunks per CTA): 16 producer warps
//                          and a dedicated MMA warp, full / free mbarriers per ring slot, no
//                          CTA-wide barrier in the loop, one CTA per SM
//   gemm_tf32x3_kernel     short contractions (the K = R products): 8 warps, a barrier per chunk,
//                          one elected thread issues the MMAs; two CTAs per SM so that one CTA's
//                          write-out overlaps the other's loads (also the PYFASST_GEMM_WS=0 path)
#include "common.cuh"
#include "tc.cuh"

namespace pf {

constexpr int GT_THREADS = 256;
constexpr int GT_BM = 128;
constexpr int GT_BK = 32;

// Operand loads do not allocate in L1: the operand ring leaves ~30 KB of the SM's 256 KB to L1,
// and a load that allocates holds a line there while it is in flight -- with __ldg every variant
// of this kernel (1 or 2 CTAs per SM, 1..4 chunks of register prefetch, 2..4 ring stages) streamed
// an F x N plane at the same ~2.5 TB/s, L2-resident or not (profiles/r02/gemm_shapes_experiments.txt).
__device__ __forceinline__ float4 gt_ldg4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
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
// Thread t owns the float4 i = t + 256 q, q < NV, of every chunk: everything that does not change
// from chunk to chunk -- its first global address, the stride between its float4, the bounds
// predicates, the shared-memory offsets (affine in q) -- is computed ONCE by init(); a chunk that
// lies inside [kbeg, kend) then costs one load, 4 cvt, 4 subtractions and two 16-byte stores per
// float4 (ncu of round 1's per-chunk index arithmetic: 45 instructions per float4, the producers
// -- not the tensor core -- set the pace: profiles/r02/ncu_gemm_tf32x3_kernel_*.txt).
template <int ROWS, bool KMAJOR, int NT = GT_THREADS>
struct OperandTile {
  static constexpr int NV = ROWS * GT_BK / 4 / NT;          // float4 per thread
  static constexpr uint32_t LBO = GT_BK * 128, SBO = 512;   // MN-major: k rows stacked per group
  static constexpr int VPR = ROWS / 4;                      // !KMAJOR: float4 per k row
  static constexpr int KPQ = NT / VPR;                      // !KMAJOR: k rows between q and q + 1
  static_assert(NV >= 1 && (KMAJOR || KPQ % 4 == 0), "k & 3 must not depend on q");
  // shared-memory byte offset of float4 q = off0 + q * QOFF (K-major: NT / 8 rows further)
  static constexpr uint32_t QOFF = KMAJOR ? (uint32_t)(NT / 64) * 1024u : (uint32_t)(KPQ / 4) * SBO;

  struct Lane {
    const float* p;    // address of float4 q = 0 in the chunk at k = kbeg
    long qstride;      // floats between float4 q and q + 1
    long kstride;      // floats between consecutive chunks
    uint32_t off0;     // shared-memory offset of float4 q = 0
    uint32_t rowmask;  // KMAJOR: bit q = row of float4 q is inside the matrix; !KMAJOR: all or none
    int kk;            // KMAJOR: first k of the thread's float4 within a chunk; !KMAJOR: k row of q = 0
  };

  __device__ __forceinline__ static Lane init(const float* __restrict__ P, long ld, long r0,
                                              long rmax, long kbeg, int tid) {
    Lane L;
    if (KMAJOR) {
      const int r = tid >> 3, c = tid & 7;
      L.p = P + (r0 + r) * ld + kbeg + c * 4;
      L.qstride = (long)(NT / 8) * ld;
      L.kstride = GT_BK;
      L.off0 = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
      L.rowmask = 0;
#pragma unroll
      for (int q = 0; q < NV; ++q)
        if (r0 + r + (NT / 8) * q < rmax) L.rowmask |= 1u << q;
      L.kk = c * 4;
    } else {
      const int k = tid / VPR, c = tid % VPR;
      L.p = P + (kbeg + k) * ld + r0 + c * 4;
      L.qstride = (long)KPQ * ld;
      L.kstride = (long)GT_BK * ld;
      L.off0 = tc::mnmajor_off(k, c * 4, LBO, SBO);
      L.rowmask = (r0 + c * 4 + 4 <= ld) ? 0xFFFFFFFFu : 0u;
      L.kk = k;
    }
    return L;
  }

  float4 v[NV];

  // chunk number `chunk` (from kbeg); klim = kend - (kbeg + 32 chunk): 32 or more inside
  __device__ __forceinline__ void fetch(const Lane& L, int chunk, long klim) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = L.p + chunk * L.kstride;
    if (klim >= GT_BK) {
#pragma unroll
      for (int q = 0; q < NV; ++q)
        v[q] = ((L.rowmask >> q) & 1u) ? gt_ldg4(p + q * L.qstride) : zero4;
    } else {
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const bool ok = KMAJOR ? (L.kk + 4 <= klim) : (L.kk + q * KPQ < klim);
        v[q] = (((L.rowmask >> q) & 1u) && ok) ? gt_ldg4(p + q * L.qstride) : zero4;
      }
    }
  }
  __device__ __forceinline__ void store(const Lane& L, unsigned char* hi, unsigned char* lo) const {
#pragma unroll
    for (int q = 0; q < NV; ++q) gt_st_split4(hi, lo, L.off0 + q * QOFF, v[q]);
  }
  // descriptor of the first K step (8 of the 32 k) of a tile at shared address `saddr`; step j:
  // + j * DSTEP (the start-address field counts 16 bytes: 32 bytes / 1024 bytes further)
  static constexpr uint64_t DSTEP = KMAJOR ? 2 : 64;
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
// (PF register sets of prefetched operand chunks, 16 + BN / 8 registers each on top of ~64 --
// beyond 128 registers per thread only one CTA fits)
template <int BN, int STAGES, int PF>
struct GtOcc {
  static constexpr int value =
      (sizeof(GtSmem<BN, STAGES>) + 1024 <= 110 * 1024 && PF * (16 + BN / 8) <= 64) ? 2 : 1;
};
static_assert(sizeof(GtStage<64>) >= (GT_THREADS / 32) * 32 * 36 * sizeof(float), "transpose tiles");

// Write-out of a CTA's 128 x BN accumulator by its first NWARPS warps: warp w reads TMEM lanes
// 32 (w % 4) .. +31 (rows of the tile) and the columns of group w / 4.
//  TC = false: C[row][col] -- 32 x 32 blocks are transposed through shared memory (`scratch`: the
//              operand ring, free once all MMAs have completed) and stored as 128-byte row segments
//  TC = true : C[col][row] (the transposed product) -- a lane holds one row, so for every column
//              the warp's 32 rows are 128 contiguous bytes: stored straight from the registers
template <int BN, bool TC, int NWARPS = 8>
__device__ __forceinline__ void gt_write_out(float* scratch, uint32_t tmem, bool have_acc,
                                             float* __restrict__ C, long ldc, int M, int N, long m0,
                                             long n0, int warp, int lane) {
  float* tr = scratch + warp * (32 * 36);
  const int wq = warp & 3, half = warp >> 2;
  constexpr int CPW = BN / (NWARPS / 4) < 32 ? 32 : BN / (NWARPS / 4);  // columns per warp group
  if (half * CPW >= BN) return;
#pragma unroll 1
  for (int c0 = 0; c0 < CPW; c0 += 32) {
    uint32_t v[32];
    if (have_acc) {
      tc::tmem_ld_32x32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * CPW + c0), v);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    }
    if (TC) {
      const long row = m0 + wq * 32 + lane;
      const long col0 = n0 + half * CPW + c0;
      if (row < M) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < N) C[(col0 + i) * ldc + row] = __uint_as_float(v[i]);
      }
      continue;
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

// TA: A is stored [K][M] (A^T given); TB: B is stored [N][K] (B^T given)
// PF: operand chunks in flight per thread (register sets).  One set exposes the whole DRAM latency
// in every K step (the loads of chunk s + 1 are issued one barrier before they are needed) and
// asks DRAM for 128 bytes of a row at a time; PF sets are loaded PF chunks ahead, and the PF
// consecutive chunks of one row are requested back to back (PF x 128 contiguous bytes).
// Tiles are numbered along x with the dimension that has FEWER tiles running fastest: the CTAs
// that run at the same time then share the strip of the large operand (C_f0 = WF0^T [num|den]:
// the 4 row tiles of one column tile read the same 1025 x 256 block of the 848 MB plane pair
// from L2 instead of re-reading the planes from HBM 4 times).
template <bool TA, bool TB, int BN, int STAGES, int PF>
__global__ void __launch_bounds__(GT_THREADS, GtOcc<BN, STAGES, PF>::value)
gemm_tf32x3_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                   float* __restrict__ C, long ldc, int M, int N, int K, int kper, long cstride,
                   int mtiles, int ntiles) {
  extern __shared__ __align__(1024) unsigned char gt_smem[];
  __shared__ uint64_t mbar_free[2];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base;
  unsigned char* base = gt_smem + ((1024 - (tc::smem_u32(gt_smem) & 1023)) & 1023);
  GtSmem<BN, STAGES>& sm = *reinterpret_cast<GtSmem<BN, STAGES>*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int mt = mtiles <= ntiles ? tile % mtiles : tile / ntiles;
  const int nt = mtiles <= ntiles ? tile / mtiles : tile % ntiles;
  const long m0 = (long)mt * GT_BM, n0 = (long)nt * BN;
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

  OperandTile<GT_BM, !TA> ta[PF];  // A is K-major unless its transpose is what is stored
  OperandTile<BN, TB> tb[PF];      // B is K-major when B^T ([N][K]) is what is stored
  const auto la = OperandTile<GT_BM, !TA>::init(A, lda, m0, M, kbeg, tid);
  const auto lb = OperandTile<BN, TB>::init(B, ldb, n0, N, kbeg, tid);
#pragma unroll
  for (int p = 0; p < PF; ++p)
    if (p < nchunks) {
      ta[p].fetch(la, p, kend - kbeg - (long)p * GT_BK);
      tb[p].fetch(lb, p, kend - kbeg - (long)p * GT_BK);
    }
  for (int s0 = 0; s0 < nchunks; s0 += PF) {
#pragma unroll
    for (int p = 0; p < PF; ++p) {
      const int s = s0 + p;
      if (s >= nchunks) break;  // (uniform over the CTA)
      const int b = s % STAGES;
      GtStage<BN>& st = sm.stage[b];
      if (s >= STAGES) tc::mbar_wait(&mbar_free[b], (uint32_t)((s / STAGES - 1) & 1));
      ta[p].store(la, st.a_hi, st.a_lo);
      tb[p].store(lb, st.b_hi, st.b_lo);
      if (s + PF < nchunks) {
        ta[p].fetch(la, s + PF, kend - kbeg - (long)(s + PF) * GT_BK);
        tb[p].fetch(lb, s + PF, kend - kbeg - (long)(s + PF) * GT_BK);
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
  }
  if (nchunks > 0) tc::mbar_wait(&mbar_done, 0);
  tc::fence_after_thread_sync();

  gt_write_out<BN, false>(reinterpret_cast<float*>(&sm.stage[0]), tmem, nchunks > 0, C, ldc, M, N,
                          m0, n0, warp, lane);
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}


// ---- the same product with a dedicated MMA warp (round 2) ----------------------------------------
// ncu of the kernel above on the SIMM shapes (profiles/r02/ncu_gemm_tf32x3_kernel_*.txt): the top
// stall is the CTA barrier of every K step -- the elected thread is ALSO a producer, so each step
// costs store + barrier + the serial issue of 12 MMAs (tensor pipe 24 % on the skinny split-K
// products, which should be HBM-bound, 54 % on the large ones).  Here warps 0..7 only produce:
// they store chunk s, make it visible to the async proxy and ARRIVE on full[s % 2] (256
// arrivals); warp 8 waits for it, issues the MMAs and commits to free[s % 2] -- no CTA-wide
// barrier inside the loop, the producers run up to two chunks ahead of the tensor core.
// TC: the product is stored transposed (C[col][row]); the host uses it to run a product with a
// SHORT M (C = A^T B, M <= 64: C_hm = WM^T [planes], M = 40) as C^T = B^T A with the long
// dimension on the 128 MMA rows instead of padding M to 128 (tensor-bound on zeros: 54 % busy).
// 16 producer warps: the producers are latency-bound at two warps per scheduler (ncu: issue slots
// 39 % busy, ~240 instructions per warp and chunk) -- four per scheduler with half the float4 each
constexpr int GW_PRODUCERS = 512;
constexpr int GW_THREADS = GW_PRODUCERS + 32;

// operand ring of the dedicated-MMA-warp kernel: as many stages as fit in 192 KB (4 x 48 KB at
// BN = 64, 3 x 64 KB at 128, 2 x 96 KB at 256) -- the round trip arrive -> issue -> MMA -> commit
// of a chunk is ~1000 cycles, with two stages the producers of the narrow tiles waited for it in
// every other chunk
template <int BN>
struct GwStages {
  static constexpr int value = BN == 64 ? 4 : (BN == 128 ? 3 : 2);
};

template <bool TA, bool TB, int BN, int PF, bool TC>
__global__ void __launch_bounds__(GW_THREADS, 1)
gemm_tf32x3_ws_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                      float* __restrict__ C, long ldc, int M, int N, int K, int kper, long cstride,
                      int mtiles, int ntiles) {
  constexpr int STAGES = GwStages<BN>::value;
  using TileA = OperandTile<GT_BM, !TA, GW_PRODUCERS>;  // A is K-major unless its transpose is what is stored
  using TileB = OperandTile<BN, TB, GW_PRODUCERS>; // B is K-major when B^T ([N][K]) is what is stored
  extern __shared__ __align__(1024) unsigned char gt_smem[];
  __shared__ uint64_t mbar_full[STAGES];
  __shared__ uint64_t mbar_free[STAGES];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base;
  unsigned char* base = gt_smem + ((1024 - (tc::smem_u32(gt_smem) & 1023)) & 1023);
  GtSmem<BN, STAGES>& sm = *reinterpret_cast<GtSmem<BN, STAGES>*>(base);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int mt = mtiles <= ntiles ? tile % mtiles : tile / ntiles;
  const int nt = mtiles <= ntiles ? tile / mtiles : tile % ntiles;
  const long m0 = (long)mt * GT_BM, n0 = (long)nt * BN;
  const long kbeg = (long)blockIdx.z * kper;
  const long kend = kbeg + kper < (long)K ? kbeg + kper : (long)K;
  C += (long)blockIdx.z * cstride;
  const int nchunks = kend > kbeg ? (int)((kend - kbeg + GT_BK - 1) / GT_BK) : 0;
  constexpr uint32_t TCOLS = BN < 32 ? 32 : BN;

  if (warp == 0) tc::tmem_alloc(&tmem_base, TCOLS);
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < STAGES; ++b) {
      tc::mbar_init(&mbar_full[b], GW_PRODUCERS);
      tc::mbar_init(&mbar_free[b], 1);
    }
    tc::mbar_init(&mbar_done, 1);
    tc::fence_mbar_init();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;

  // slot b is used for chunks b, b + STAGES, ..: `ph` = parity of the use in progress
  if (warp == GW_PRODUCERS / 32) {
    // ---- MMA warp: every lane follows the chunks, lane 0 issues
    const uint32_t idesc = tc::idesc_tf32(GT_BM, BN, TA ? 1 : 0, TB ? 0 : 1);
    int b = 0;
    uint32_t ph = 0;
    for (int s = 0; s < nchunks; ++s) {
      GtStage<BN>& st = sm.stage[b];
      tc::mbar_wait(&mbar_full[b], ph);
      if (lane == 0) {
        tc::fence_after_thread_sync();
        const uint64_t dah = TileA::desc(tc::smem_u32(st.a_hi), 0), dal = TileA::desc(tc::smem_u32(st.a_lo), 0);
        const uint64_t dbh = TileB::desc(tc::smem_u32(st.b_hi), 0), dbl = TileB::desc(tc::smem_u32(st.b_lo), 0);
#pragma unroll
        for (int j = 0; j < GT_BK / 8; ++j) {
          const uint64_t ja = j * TileA::DSTEP, jb = j * TileB::DSTEP;
          tc::mma_tf32(tmem, dah + ja, dbh + jb, idesc, (s > 0 || j > 0) ? 1u : 0u);
          tc::mma_tf32(tmem, dah + ja, dbl + jb, idesc, 1u);
          tc::mma_tf32(tmem, dal + ja, dbh + jb, idesc, 1u);
        }
        tc::mma_commit(&mbar_free[b]);
        if (s == nchunks - 1) tc::mma_commit(&mbar_done);
      }
      __syncwarp();
      if (++b == STAGES) {
        b = 0;
        ph ^= 1u;
      }
    }
  } else {
    // ---- producers
    TileA ta[PF];
    TileB tb[PF];
    const auto la = TileA::init(A, lda, m0, M, kbeg, tid);
    const auto lb = TileB::init(B, ldb, n0, N, kbeg, tid);
    const long kspan = kend - kbeg;
#pragma unroll
    for (int p = 0; p < PF; ++p)
      if (p < nchunks) {
        ta[p].fetch(la, p, kspan - (long)p * GT_BK);
        tb[p].fetch(lb, p, kspan - (long)p * GT_BK);
      }
    int b = 0;
    uint32_t ph = 0;
    for (int s0 = 0; s0 < nchunks; s0 += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        const int s = s0 + p;
        if (s >= nchunks) break;
        GtStage<BN>& st = sm.stage[b];
        if (s >= STAGES) tc::mbar_wait(&mbar_free[b], ph ^ 1u);
        ta[p].store(la, st.a_hi, st.a_lo);
        tb[p].store(lb, st.b_hi, st.b_lo);
        if (s + PF < nchunks) {
          ta[p].fetch(la, s + PF, kspan - (long)(s + PF) * GT_BK);
          tb[p].fetch(lb, s + PF, kspan - (long)(s + PF) * GT_BK);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&mbar_full[b]);
        if (++b == STAGES) {
          b = 0;
          ph ^= 1u;
        }
      }
    }
    if (nchunks > 0) tc::mbar_wait(&mbar_done, 0);
    tc::fence_after_thread_sync();
    gt_write_out<BN, TC, GW_PRODUCERS / 32>(reinterpret_cast<float*>(&sm.stage[0]), tmem, nchunks > 0, C, ldc, M, N,
                         m0, n0, warp, lane);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

template <bool TA, bool TB, int BN, int STAGES, int PF>
static int launch_gemm_stages(const float* A, long lda, const float* B, long ldb, float* C,
                              long ldc, int M, int N, int K, int ksplit, int kper, long cstride,
                              cudaStream_t st) {
  const size_t smem = sizeof(GtSmem<BN, STAGES>) + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<TA, TB, BN, STAGES, PF>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("gemm_tf32x3_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  const int mtiles = ceil_div(M, GT_BM), ntiles = ceil_div(N, BN);
  dim3 grid((unsigned)((long)mtiles * ntiles), 1, ksplit);
  gemm_tf32x3_kernel<TA, TB, BN, STAGES, PF><<<grid, GT_THREADS, smem, st>>>(
      A, lda, B, ldb, C, ldc, M, N, K, kper, cstride, mtiles, ntiles);
  return check_launch("gemm_tf32x3_kernel");
}

// PYFASST_GEMM_WS=0: the barrier-per-chunk kernel for long contractions too (round 1)
static bool gemm_use_ws() {
  static const bool v = [] {
    const char* e = getenv("PYFASST_GEMM_WS");
    return !(e && e[0] == '0');
  }();
  return v;
}

template <bool TA, bool TB, int BN, int PF, bool TC>
static int launch_gemm_ws(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                          int M, int N, int K, int ksplit, int kper, long cstride, cudaStream_t st) {
  const size_t smem = sizeof(GtSmem<BN, GwStages<BN>::value>) + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_ws_kernel<TA, TB, BN, PF, TC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("gemm_tf32x3_ws_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  const int mtiles = ceil_div(M, GT_BM), ntiles = ceil_div(N, BN);
  dim3 grid((unsigned)((long)mtiles * ntiles), 1, ksplit);
  gemm_tf32x3_ws_kernel<TA, TB, BN, PF, TC><<<grid, GW_THREADS, smem, st>>>(
      A, lda, B, ldb, C, ldc, M, N, K, kper, cstride, mtiles, ntiles);
  return check_launch("gemm_tf32x3_ws_kernel");
}

template <bool TA, bool TB, int BN>
static int launch_gemm(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                       int M, int N, int K, int ksplit, int kper, long cstride, cudaStream_t st) {
  // short contractions (at most two K chunks per CTA: the K = R products of the SIMM
  // accompaniment model) take the one-stage ring and two CTAs per SM
  const long kspan = ksplit > 1 ? kper : K;
  if (kspan <= 2 * GT_BK)  // (one register set: two CTAs per SM also at BN = 256)
    return launch_gemm_stages<TA, TB, BN, 1, 1>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper,
                                                cstride, st);
  // long contractions: dedicated MMA warp, one CTA per SM, 4 (BN = 64) / 2 chunks in flight
  if (gemm_use_ws())
    return launch_gemm_ws<TA, TB, BN, BN == 64 ? 4 : 2, false>(A, lda, B, ldb, C, ldc, M, N, K,
                                                               ksplit, kper, cstride, st);
  return launch_gemm_stages<TA, TB, BN, 2, 2>(A, lda, B, ldb, C, ldc, M, N, K, ksplit, kper,
                                              cstride, st);
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
  // C = A^T B with a short M (and a long N) as C^T = B^T A on the 128 MMA rows, stored transposed:
  // measured SLOWER than padding M to 128 (C_hm, M = 40: 666 against 531 us -- the 128-column
  // operand rows of the narrow tile are 512-byte runs, the 256-column ones 1 KB), so only on
  // request (PYFASST_GEMM_TC=1; parity-tested)
  static const bool use_tc = [] {
    const char* e = getenv("PYFASST_GEMM_TC");
    return e && e[0] == '1';
  }();
  if (use_tc && transA && !transB && M <= 64 && N >= 4 * GT_BM && ksplit == 1 && K > 2 * GT_BK &&
      gemm_use_ws())
    return launch_gemm_ws<true, false, 64, 4, true>(B, ldb, A, lda, C, ldc, N, M, K, 1, K, 0, st);
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
  // (the long-K kernel of the narrow tile runs one CTA per SM)
  long slots = (bn == 64 && gemm_use_ws()) ? 148 : 2L * 148;
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
