// tcgen05 (5th-generation tensor core) building blocks for sm_100a: shared-memory matrix
// descriptors, the instruction descriptor of kind::tf32, TMEM allocation, MMA issue /
// commit, TMEM -> register loads and mbarrier helpers.  Inline PTX only.
//
// Canonical shared-memory operand layouts used here (fp32 elements read as tf32, 128-byte
// swizzle, tile base 1024-byte aligned):
//
//   K-major  [rows][32 k]  : byte(r, k) = (r/8)*1024 + (r%8)*128 + (((k/4) ^ (r%8)) * 16) + (k%4)*4
//                            one MMA (K = 8) reads 32 bytes of each row: descriptor start
//                            address advanced by 32 bytes per K step; SBO = 1024 (8-row groups)
//   MN-major [k][mn]       : tf32 operands that are contiguous along M / N must use the
//                            SWIZZLE_128B_BASE32B layout (32-byte swizzle granules, 4-row atoms):
//                            byte(k, m) = (m/32)*LBO + (k/4)*SBO + (k%4)*128
//                                         + ((((m%32)/8) ^ (k%4)) * 32) + (m%8)*4
//                            LBO = distance between 32-element groups along MN, SBO = distance
//                            between 4-row groups along K; one MMA (K = 8) reads two K groups
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- descriptors ---------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (Blackwell version field = 1).
// layout: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (MN-major tf32 tiles)
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr) {
  return smem_desc(saddr, 16, 1024, LAYOUT_SW128);
}
__device__ __forceinline__ uint64_t smem_desc_mnmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return smem_desc(saddr, lbo, sbo, LAYOUT_SW128_BASE32B);
}

// 32-bit instruction descriptor: D = F32, A = B = TF32, dense, M x N, operand majors
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row r, k) in a K-major SW128 tile of 32 fp32 per row
__device__ __forceinline__ uint32_t kmajor_off(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4);
}
// byte offset of element (k, m) in an MN-major SW128_BASE32B tile
__device__ __forceinline__ uint32_t mnmajor_off(int k, int m, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)((m >> 5) * lbo + (k >> 2) * sbo + (k & 3) * 128 +
                    (((((m & 31) >> 3) ^ (k & 3)) & 3) << 5) + (m & 7) * 4);
}

// ---- TMEM ------------------------------------------------------------------------------
// one full warp allocates `ncols` (power of two >= 32) columns; the base address is written
// to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (the MMA reads smem there)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- MMA -------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(mbar))
               : "memory");
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count)
               : "memory");
}
// plain arrival of the calling thread (release semantics at CTA scope)
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// named barrier over `nthreads` threads (a multiple of 32) of the CTA, id 1..15
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(mbar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  while (!mbar_try_wait(mbar, parity)) {
  }
}

// ---- TMEM -> registers ---------------------------------------------------------------------
// 32 lanes x 32 columns: thread t of warp w (w = warp index % 4) receives lane 32 w + t,
// columns [col, col + 32) of its lane in v[0..31].  taddr = base + (lane << 16) + column.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 columns (same lane mapping)
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- TMA (bulk tensor) stores: shared -> global through a CUtensorMap ------------------------
// box written by cp.async.bulk.tensor.2d: `smem_addr` holds the box row-major in the tensor map's
// swizzle (SWIZZLE_128B with 128-byte rows: 16-byte chunk c of row r sits at chunk c ^ (r & 7);
// tile base 1024-byte aligned); coordinates (x = innermost = column, y = row); parts of the box
// outside the tensor are not written.  Issued by ONE thread, which also commits / waits its groups.
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_addr, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :: "l"(tmap), "r"(smem_addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all but the newest N groups of this thread have finished READING shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
// ... have completed (their writes are done)
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}

// ---- 3xTF32 split ----------------------------------------------------------------------------
// x = hi + lo with hi = tf32-rounded x (round to nearest, ties away from zero) and lo = x - hi
// (exact in fp32); A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi has a relative error of ~2^-21 per product,
// i.e. fp32-class.  The rounding is done on the bit pattern -- add half an ulp of tf32 to the
// magnitude, clear the 13 low mantissa bits: exactly `cvt.rna.tf32.f32` for every finite x --
// because ptxas EMULATES that conversion on sm_100a (no CVT in the SASS: ~10 ISETP / FSETP / SEL /
// LOP3 / VIADD per element for its NaN / infinity cases), which made the operand producers of
// every tcgen05 kernel here instruction-bound (ncu of gemm_tf32x3_ws_kernel: 40 instructions per
// float4, profiles/r02/ncu_gemm_tf32x3_ws_kernel_D_q.txt).  Two integer operations instead.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

}  // namespace tc
}  // namespace pf
