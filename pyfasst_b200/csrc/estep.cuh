// Shared pieces of the stereo E-step kernels (estep.cu: fused kernel, Wiener filter, host entry
// points; estep_ws.cu: warp-specialised kernel).
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace pf {

#ifndef PF_ESTEP_THREADS
#define PF_ESTEP_THREADS 128
#endif
#ifndef PF_ESTEP_MINB
#define PF_ESTEP_MINB 2
#endif
constexpr int ESTEP_THREADS = PF_ESTEP_THREADS;
constexpr int ESTEP_MINB = PF_ESTEP_MINB;  // CTAs per SM the register allocation aims for
#ifndef PF_ESTEP_MINB_SMEMIO
#define PF_ESTEP_MINB_SMEMIO 3
#endif
constexpr int ESTEP_MINB_SMEMIO = PF_ESTEP_MINB_SMEMIO;  // same, shared-memory-resident I/O variant
constexpr int MAXJ = 6;
constexpr int MAXR = 16;
constexpr int PF_F32_FASTMATH = 2;
constexpr int ESTEP_DEFAULT_VARIANT = 35;  // packed moments + hardware conversions + interleaved splits
#ifndef PF_ESTEP_DEPTH
#define PF_ESTEP_DEPTH 3
#endif
constexpr int ESTEP_DEPTH = PF_ESTEP_DEPTH;  // passes in flight in the cp.async ring (OPT bit 2)  // float storage AND float per-bin algebra (experiments)

__host__ __device__ constexpr int npairs(int J) { return J * (J + 1) / 2; }
// accumulators per frequency: S (4 per pair), T (8 per source), sv (J), ll (1)
__host__ __device__ constexpr int nacc(int J) { return 4 * npairs(J) + 8 * J + J + 1; }
// coefficients per frequency: R_j (4 per source), D_jk (per pair)
__host__ __device__ constexpr int ncoef(int J) { return 4 * J + npairs(J); }

struct SubMap {
  int src_of_sub[MAXR];  // spatial component of each sub-source (rank column)
  double invrank[MAXJ];
};

// 1/x: for double, a float reciprocal refined by two Newton steps (4 DFMA) instead of
// the ~20-instruction IEEE division; relative error < 1e-14.
// MUFU.RCP plus one Newton step (relative error ~1e-7, like the IEEE division, but without its
// range check and slow-path call; the arguments here are clamped to >= 1e-10)
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ double fast_rcp(double x) {
  float r0;  // MUFU.RCP (2^-23 relative): no IEEE-division slow path to branch to
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)x));
  double r = (double)r0;
  r = r * (2.0 - x * r);
  r = r * (2.0 - x * r);
  return r;
}

// float -> double without the (quarter-rate, XU pipe) F2F conversion: re-bias the exponent with
// integer ops.  Exact for normal numbers; zero and denormals map to |x| < 1.2e-38, which is
// far below every eps clamp of this path.  The E-step is limited by the XU pipe otherwise
// (profiles/r01: 23 conversions per bin).
__device__ __forceinline__ double widen(float x) {
  const unsigned f = __float_as_uint(x);
  const unsigned hi = (((f & 0x7fffffffu) >> 3) + 0x38000000u) | (f & 0x80000000u);
  return __hiloint2double((int)hi, (int)(f << 29));
}
__device__ __forceinline__ double widen(double x) { return x; }
// the hardware conversion (F2F.F64.F32, XU pipe): one issue slot instead of five
__device__ __forceinline__ double widen_hw(float x) { return (double)x; }
__device__ __forceinline__ double widen_hw(double x) { return x; }

// ---- packed float32 pairs (FFMA2 / FMUL2 / FADD2 of sm_100): the per-frequency moment sums are
// 40% of the instructions of the E-step; two accumulators share one instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// acc += a * b / acc += a  (in-out operand: the accumulator keeps its register pair)
__device__ __forceinline__ void fma2_acc(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void add2_acc(f32x2& acc, f32x2 a) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a));
}

// The moment accumulators of one thread: S (4 per source pair), T (8 per source), sv (J), in
// the order of the `partial` array.  Scalar version (any type) and packed float32 version.
template <typename T, int J, bool PACK>
struct Moments {
  static constexpr int NP = J * (J + 1) / 2;
  static constexpr int COUNT = 4 * NP + 9 * J;
  T acc[COUNT];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < COUNT; ++i) acc[i] = (T)0;
  }
  __device__ __forceinline__ void add(const T (&pr)[NP], const T (&vt)[J], T t00, T t11, T t01r,
                                      T t01i, const T (&u)[8]) {
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      acc[4 * p + 0] += pr[p] * t00;
      acc[4 * p + 1] += pr[p] * t11;
      acc[4 * p + 2] += pr[p] * t01r;
      acc[4 * p + 3] += pr[p] * t01i;
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[4 * NP + 8 * j + e] += vt[j] * u[e];
      acc[4 * NP + 8 * J + j] += vt[j];
    }
  }
  __device__ __forceinline__ T get(int i) const { return acc[i]; }
};

template <int J>
struct Moments<float, J, true> {
  static constexpr int NP = J * (J + 1) / 2;
  static constexpr int COUNT = 4 * NP + 9 * J;
  static constexpr int NPK = (COUNT + 1) / 2;
  f32x2 acc[NPK];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < NPK; ++i) acc[i] = 0ull;
  }
  __device__ __forceinline__ void add(const float (&pr)[NP], const float (&vt)[J], float t00,
                                      float t11, float t01r, float t01i, const float (&u)[8]) {
    const f32x2 ta = pack2(t00, t11), tb = pack2(t01r, t01i);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const f32x2 pp = pack2(pr[p], pr[p]);  // broadcast operand: no instruction
      fma2_acc(acc[2 * p + 0], pp, ta);
      fma2_acc(acc[2 * p + 1], pp, tb);
    }
    const f32x2 u0 = pack2(u[0], u[1]), u1 = pack2(u[2], u[3]);
    const f32x2 u2 = pack2(u[4], u[5]), u3 = pack2(u[6], u[7]);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const f32x2 vv = pack2(vt[j], vt[j]);
      fma2_acc(acc[2 * NP + 4 * j + 0], vv, u0);
      fma2_acc(acc[2 * NP + 4 * j + 1], vv, u1);
      fma2_acc(acc[2 * NP + 4 * j + 2], vv, u2);
      fma2_acc(acc[2 * NP + 4 * j + 3], vv, u3);
    }
#pragma unroll
    for (int j = 0; j < J; j += 2)
      add2_acc(acc[2 * NP + 4 * J + j / 2], pack2(vt[j], j + 1 < J ? vt[j + 1] : 0.f));
  }
  __device__ __forceinline__ float get(int i) const {
    float a, b;
    unpack2(acc[i >> 1], a, b);
    return (i & 1) ? b : a;
  }
};

// Per-bin algebra shared by the E-step and the Wiener filter.
// Sigma = s2 I + sum_j v_j R_j ; returns Sigma^-1 (i00, i11, i01) in the compute type C and
// det Sigma / the pair products v_j v_k in the type D.  The determinant is expanded into
// non-negative terms (no s00*s11 - |s01|^2 cancellation), so it is safe in float32; only the
// entries of Sigma (which the adjugate later cancels against x) need the type C.  A common
// relative error of 1/det scales Sigma^-1 and y together and is harmless.
template <typename C, typename D, int J>
__device__ __forceinline__ void sigma_inverse(const C (&vj)[J], const D (&vd)[J],
                                              const C* __restrict__ coef,
                                              const D* __restrict__ dcoef, C s2,
                                              D (&pr)[J * (J + 1) / 2], D& det, C& i00, C& i11,
                                              C& i01r, C& i01i) {
  C s00 = s2, s11 = s2, s01r = (C)0, s01i = (C)0;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    s00 += vj[j] * coef[4 * j + 0];
    s11 += vj[j] * coef[4 * j + 1];
    s01r += vj[j] * coef[4 * j + 2];
    s01i += vj[j] * coef[4 * j + 3];
  }
  const D d2 = (D)s2;
  det = d2 * ((D)s00 + ((D)s11 - d2));
  int p = 0;
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int k = j; k < J; ++k) {
      pr[p] = vd[j] * vd[k];
      det += pr[p] * dcoef[p];
      ++p;
    }
  det = pf_max(det, (D)1e-10);  // Q5 clamp (det >= 0 here, so sign(det+eps) = +1)
  const C idet = (C)fast_rcp(det);
  i00 = s11 * idet;
  i11 = s00 * idet;
  i01r = -s01r * idet;
  i01i = -s01i * idet;
}


// ---- warp-specialised float32 kernel (estep_ws.cu) --------------------------------
// true when pf_estep_stereo / pf_estep_plan use it for this shape
bool estep_use_ws(int J, long N, int dtype);
// frames one CTA covers per pass, and the passes a CTA should get
long estep_ws_pass();
long estep_ws_passes_per_cta();
int dispatch_estep_ws(int J, const void* X, const void* V, const double* coef,
                      const double* noise, const SubMap& map, void* hatW, double* partial, int F,
                      long N, long ld, long chunk, int nsplit, cudaStream_t st);

}  // namespace pf
