// Shared helpers for the pyfasst_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pyfasst_b200.h"

namespace pf {

// ---- error plumbing: no exception crosses the C ABI -------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PF_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      pf::set_error(__VA_ARGS__);        \
      return PF_ERR_ARG;                 \
    }                                    \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- vector access: VEC consecutive frames per thread -----------------------
template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  typedef float4 type;
  static constexpr int N = 4;
};
template <>
struct VecOf<double> {
  typedef double2 type;
  static constexpr int N = 2;
};

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, T (&out)[VecOf<T>::N]);
template <>
__device__ __forceinline__ void load_vec<float>(const float* p, float (&out)[4]) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec<double>(const double* p, double (&out)[2]) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  out[0] = v.x; out[1] = v.y;
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const T (&in)[VecOf<T>::N]);
template <>
__device__ __forceinline__ void store_vec<float>(float* p, const float (&in)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <>
__device__ __forceinline__ void store_vec<double>(double* p, const double (&in)[2]) {
  *reinterpret_cast<double2*>(p) = make_double2(in[0], in[1]);
}

// ---- math wrappers so one kernel body serves float and double ----------------
__device__ __forceinline__ float pf_log(float x) { return logf(x); }
__device__ __forceinline__ double pf_log(double x) { return log(x); }
__device__ __forceinline__ float pf_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double pf_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float pf_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ double pf_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float pf_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double pf_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float pf_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double pf_abs(double a) { return fabs(a); }
__device__ __forceinline__ float pf_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double pf_sqrt(double a) { return sqrt(a); }

// 1/x: MUFU.RCP seed plus one (float) or two (double) Newton steps instead of the IEEE
// division with its range check and slow-path call; relative error ~1e-7 / < 1e-14.  The
// arguments here are clamped away from zero.
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ double fast_rcp(double x) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)x));
  double r = (double)r0;
  r = r * (2.0 - x * r);
  r = r * (2.0 - x * r);
  return r;
}
// the same with the seed from MUFU.RCP64H on the high word (~2^-20 relative; no float round trip:
// two conversions less, the same two Newton steps)
__device__ __forceinline__ double fast_rcp_h(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * (2.0 - x * r);
  r = r * (2.0 - x * r);
  return r;
}

// ---- warp / block reductions (fixed order => deterministic) -------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = v > w ? v : w;
  }
  return v;
}

// ---- cp.async (LDGSTS): global -> shared without staging registers -------------
// 16-byte copy, L2 only (.cg); src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src),
               "r"(src_bytes));
}
// 4- or 8-byte copy (.ca), for small strided tiles
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gmem_src,
                                               int src_bytes) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(dst), "l"(gmem_src),
               "n"(BYTES), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING));
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace pf
