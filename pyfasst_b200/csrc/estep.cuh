// Shared pieces of the stereo E-step and Wiener kernels (estep.cu).
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace pf {

constexpr int ESTEP_THREADS = 128;
constexpr int ESTEP_DEPTH = 3;  // passes in flight in the cp.async ring of the E-step
constexpr int MAXJ = 6;
constexpr int MAXR = 16;

__host__ __device__ constexpr int npairs(int J) { return J * (J + 1) / 2; }
// accumulators per frequency and CTA: S (4 per source pair), Z (4 per source), sv (J),
// clamp corrections (8 per source), ll (1)
__host__ __device__ constexpr int nacc(int J) { return 4 * npairs(J) + 13 * J + 1; }
// coefficients per frequency: R_j (4 per source), D_jk (per pair; Wiener filter only)
__host__ __device__ constexpr int ncoef(int J) { return 4 * J + npairs(J); }

struct SubMap {
  int src_of_sub[MAXR];  // spatial component of each sub-source (rank column)
  double invrank[MAXJ];
};

// Per-bin algebra of the Wiener filter.
// Sigma = s2 I + sum_j v_j R_j ; returns Sigma^-1 (i00, i11, i01) in the compute type C and
// det Sigma / the pair products v_j v_k in the type D.  The determinant is expanded into
// non-negative terms (no s00*s11 - |s01|^2 cancellation), so it is safe in float32; only the
// entries of Sigma (which the adjugate later cancels against x) need the type C.
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

}  // namespace pf
