// K2 -- fused per-bin E-step of the FASST GEM loop (stereo), sm_100a.
//
// Replaces the reference's FASST.compute_suff_stat (pyfasst/audioModel.py:580-764,
// with inv_herm_mat_2d, pyfasst/tools/signalTools.py:132-196) and the rank-mean
// of hat_Ws in GEM_iteration (audioModel.py:408-414).
//
// The reference walks Rtot^2 complex F x N planes.  Here one pass over the data
// forms, per time-frequency bin,
//     Sigma = sum_j v_j R_j + s2 I,   y = Sigma^-1 x,   M = y y^H - Sigma^-1
// (R_j = sum_{r in j} a_r a_r^H per frequency), writes the posterior source
// power  hatW_j = | v_j + v_j^2 tr(M R_j) / rank_j |, and accumulates per
// frequency the moments
//     S_jk = sum_n v_j v_k M,   T_j = sum_n v_j x y^H,   sv_j = sum_n v_j,
//     ll   = sum_n log(det Sigma * pi) + x^H Sigma^-1 x
// which a second, per-frequency kernel contracts with the mixing vectors into
// hat_Rss / hat_Rxs exactly as the reference defines them:
//     hat_Rss[r1,r2] = a_r1^H S_{j1 j2} a_r2 / N + delta_{r1 r2} sv_{j1} / N
//     hat_Rxs[:, r]  = T_j a_r / N.
// tests/kernel_model.py states the same algebra in numpy and
// tests/test_oracle_golden.py / test_kernel_model.py check it against the oracle.
//
// Layout: SoA planes, frames contiguous: X[4][F][ld] (re0, im0, re1, im1),
// V[J][F][ld], hatW[J][F][ld]; one CTA owns one frequency and a run of frames,
// a thread owns VEC consecutive frames (float4 / double2 accesses).
#include <utility>
#include <vector>

#include "estep.cuh"

#ifndef ESTEP_MINB_RR
#define ESTEP_MINB_RR 2
#endif
namespace pf {

// ---- per-frequency coefficients from the mixing matrix ----------------------
// A: complex128 [R][2][F] (mix_matrix of retrieve_subsrc_params, audioModel.py:562-576)
// coef[f] = { R_j (4 reals per source), D_jk (mixed discriminants, one per source pair) }
__global__ void spat_coef_kernel(const double2* __restrict__ A, SubMap map, int R, int J,
                                 int F, double* __restrict__ coef) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double Rj[MAXJ][4];
  for (int j = 0; j < J; ++j) Rj[j][0] = Rj[j][1] = Rj[j][2] = Rj[j][3] = 0.0;
  for (int r = 0; r < R; ++r) {
    int j = map.src_of_sub[r];
    double2 a0 = A[((size_t)r * 2 + 0) * F + f];
    double2 a1 = A[((size_t)r * 2 + 1) * F + f];
    Rj[j][0] += a0.x * a0.x + a0.y * a0.y;
    Rj[j][1] += a1.x * a1.x + a1.y * a1.y;
    Rj[j][2] += a0.x * a1.x + a0.y * a1.y;  // Re a0 conj(a1)
    Rj[j][3] += a0.y * a1.x - a0.x * a1.y;  // Im a0 conj(a1)
  }
  double* c = coef + (size_t)f * ncoef(J);
  for (int j = 0; j < J; ++j)
    for (int e = 0; e < 4; ++e) c[4 * j + e] = Rj[j][e];
  int p = 4 * J;
  for (int j = 0; j < J; ++j)
    for (int k = j; k < J; ++k) {
      double d;
      if (j == k)
        d = Rj[j][0] * Rj[j][1] - Rj[j][2] * Rj[j][2] - Rj[j][3] * Rj[j][3];
      else
        d = Rj[j][0] * Rj[k][1] + Rj[j][1] * Rj[k][0] -
            2.0 * (Rj[j][2] * Rj[k][2] + Rj[j][3] * Rj[k][3]);
      c[p++] = fmax(d, 0.0);  // mixed discriminants of PSD matrices are >= 0
    }
}

// ---- the fused per-bin kernel -------------------------------------------------
// T: storage type of the planes (float: VEC = 4 frames per thread and pass; double: 2).  The
// per-bin algebra AND the per-frequency moment sums run in float64 whatever T is:
//  * Sigma^-1 has entries ~ 1/noise and a^H (y y^H - Sigma^-1) a cancels them down by
//    cond(Sigma); in float32 that costs eps * cond ~ 1e-2 at 50-60 dB bins;
//  * the same cancellation happens AFTER the sum over the frames when the moments S_jk are
//    contracted with the mixing vectors, and hat_Rss is then inverted by the spatial M-step
//    (cond(hat_Rss) ~ 1e3..1e4 for the nearly collinear sub-sources of a fresh rank-2 model):
//    float32 moment sums (round 1) gave A to 2.5e-4 after ONE iteration and a log-likelihood
//    trajectory off by 4e-4 on the reference's own tamy.wav (tests/test_tamy_gpu.py,
//    profiles/r02/estep_precision.txt); float64 sums give 1e-8.
// To keep the float64 work per bin small the cross moments T_j = sum_n v_j x y^H (8 reals per
// source) are NOT accumulated: with Sigma y = x,
//     x y^H = Sigma M + I      =>   T_j = sv_j I + s2 Z_j + sum_l R_l S_lj,   Z_j = sum_n v_j M
// (4 reals per source; exact to ~1e-16 * cond(Sigma) in float64).  Where the determinant clamp
// (Q5, signalTools.py:186-188) is active, Sigma^-1 is not the inverse of Sigma:
// Sigma Sigma_c^-1 = kappa I with kappa = det / det_c, so x y^H = Sigma M + I + (1 - kappa)
// (x y^H - I); that correction is accumulated for the clamped bins only (a slow path: on real
// music the clamp is active in ~2 % of the bins at the end of the annealing, in none of the
// synthetic benchmark's), in per-thread slots of shared memory.
// Inputs are staged by a cp.async ring (ESTEP_DEPTH passes ahead; every thread copies and later
// reads back ITS OWN 16 bytes per plane, so no barrier is needed, only wait_group on its own
// groups) and STAY in the ring slot: a bin reads its 4 + J scalars when its turn comes and
// writes hat_W back over V, so that no input / output / prefetch vector is live across the
// algebra.  Lane l visits the VEC bins of its vector rotated by l / 8: the scalar accesses at a
// 16-byte stride are then bank-conflict free.
// RR (real R): every mixing vector is real (instantaneous mixing, audioModel.py:2349-2393), so
// R_j and Sigma are real symmetric and the instantaneous mixing update only takes the REAL parts
// of hat_Rss / hat_Rxs (np.real(np.mean(...)), audioModel.py:818-820): Im M01 is neither needed
// for hat_W (tr(M R_j) with real R_j) nor accumulated -- 112 instead of 143 FP64 operations per
// bin, 46 instead of 61 float64 sums.  The imaginary parts of the statistics come out as zero
// (GEM loop of all-instantaneous models only; compute_suff_stat takes RR = false).
template <typename T, int J, bool RR>
__global__ void __launch_bounds__(ESTEP_THREADS, RR ? ESTEP_MINB_RR : 2)
estep_stereo_kernel(const T* __restrict__ X, const T* __restrict__ V,
                    const double* __restrict__ coef, const double* __restrict__ noise,
                    SubMap map, T* __restrict__ hatW, double* __restrict__ partial, int F,
                    long N, long ld, int nsplit) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NP = npairs(J);
  constexpr int NA = nacc(J);
  constexpr int NC = ncoef(J);
  constexpr int NPL = 4 + J;
  constexpr int NM = 4 * NP + 5 * J;  // S, Z, sv: the register-resident sums
  constexpr double kLogPi = 1.1447298858494002;  // log(pi): Q4, log(det*pi)
  constexpr double kEps = 1e-10;                 // ref: audioModel.py:61, signalTools.py:11
  // RR kernel, float32 planes: the reciprocal seed from MUFU.RCP64H and the float logarithms of
  // a pass summed in float32 (0.640 -> 0.631 ms; the same two changes SLOW the general kernel,
  // 0.764 -> 0.83 ms, profiles/r02/estep_inst_experiment.txt: it keeps its round-1 forms)
  constexpr bool LLP = RR && sizeof(T) == 4;

  const int f = blockIdx.y;
  const int split = blockIdx.x;
  __shared__ double s_coef[4 * J], s_coef2[4 * J];
  __shared__ double s_red[ESTEP_THREADS / 32][NA];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  // [ESTEP_DEPTH][NPL][ESTEP_THREADS] 16-byte vectors (the ring), [J][ESTEP_THREADS] vectors
  // (hat_W of the current pass), then the clamp-correction slots [8 J][ESTEP_THREADS] doubles
  typedef typename VecOf<T>::type VT;
  VT* s_ring = reinterpret_cast<VT*>(s_dyn);
  VT* s_outv = s_ring + (size_t)ESTEP_DEPTH * NPL * ESTEP_THREADS;
  double* s_corr = reinterpret_cast<double*>(s_outv + (size_t)J * ESTEP_THREADS);
  T* so = reinterpret_cast<T*>(s_outv + threadIdx.x);
  if (threadIdx.x < 4 * J) {
    const double c = coef[(size_t)f * NC + threadIdx.x];
    s_coef[threadIdx.x] = c;
    s_coef2[threadIdx.x] = (threadIdx.x & 2) ? 2.0 * c : c;  // R00, R11, 2 Re R01, 2 Im R01
  }
#pragma unroll
  for (int i = 0; i < 8 * J; ++i) s_corr[i * ESTEP_THREADS + threadIdx.x] = 0.0;
  __syncthreads();
  const double s2 = noise[f];
  T invrank[J];
#pragma unroll
  for (int j = 0; j < J; ++j) invrank[j] = (T)map.invrank[j];

  double mom[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) mom[i] = 0.0;
  double acc_ll = 0.0;
  int ndead = 0;
  bool any_clamped = false;
  auto ll_term = [&](double det) -> double {  // log(det * pi)
    if (sizeof(T) == 8) return log(det) + kLogPi;
    return (double)(__logf((float)det) + (float)kLogPi);
  };
  // one bin of the slow path (see above): recomputed from the inputs still in the ring slot
  auto clamp_correction = [&](const T* sb, int es) {
    const double a0r = (double)sb[0 * ESTEP_THREADS * VEC + es];
    const double a0i = (double)sb[1 * ESTEP_THREADS * VEC + es];
    const double a1r = (double)sb[2 * ESTEP_THREADS * VEC + es];
    const double a1i = (double)sb[3 * ESTEP_THREADS * VEC + es];
    double vj[J], s00 = s2, s11 = s2, s01r = 0.0, s01i = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      vj[j] = (double)sb[(4 + j) * ESTEP_THREADS * VEC + es];
      s00 += vj[j] * s_coef[4 * j + 0];
      s11 += vj[j] * s_coef[4 * j + 1];
      s01r += vj[j] * s_coef[4 * j + 2];
      s01i += vj[j] * s_coef[4 * j + 3];
    }
    const double det_raw = s00 * s11 - s01r * s01r - s01i * s01i;
    const double idet = fast_rcp(kEps);
    const double y0r = (s11 * a0r - s01r * a1r + s01i * a1i) * idet;
    const double y0i = (s11 * a0i - s01r * a1i - s01i * a1r) * idet;
    const double y1r = (s00 * a1r - s01r * a0r - s01i * a0i) * idet;
    const double y1i = (s00 * a1i - s01r * a0i + s01i * a0r) * idet;
    const double k1 = 1.0 - det_raw * idet;
    const double u[8] = {a0r * y0r + a0i * y0i - 1.0, a0i * y0r - a0r * y0i,
                         a0r * y1r + a0i * y1i,       a0i * y1r - a0r * y1i,
                         a1r * y0r + a1i * y0i,       a1i * y0r - a1r * y0i,
                         a1r * y1r + a1i * y1i - 1.0, a1i * y1r - a1r * y1i};
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double c = k1 * vj[j];
#pragma unroll
      for (int i = 0; i < 8; ++i) s_corr[(8 * j + i) * ESTEP_THREADS + threadIdx.x] += c * u[i];
    }
  };

  const long plane = (long)F * ld;
  const long row = (long)f * ld;
  // the CTAs of a row take its passes in turn (CTA s works on passes s, s + nsplit, ...): they
  // are launched together, so at any moment they touch ONE contiguous region of every plane
  const long stride = (long)ESTEP_THREADS * VEC * nsplit;
  const long first = (long)split * ESTEP_THREADS * VEC + (long)threadIdx.x * VEC;
  const long end = N;
  auto ring_issue = [&](long n, int slot) {
    if (n < end) {
      VT* dst = s_ring + (size_t)slot * NPL * ESTEP_THREADS + threadIdx.x;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl)
        cp_async16(dst + pl * ESTEP_THREADS, X + pl * plane + row + n, 16);
#pragma unroll
      for (int j = 0; j < J; ++j)
        cp_async16(dst + (4 + j) * ESTEP_THREADS, V + j * plane + row + n, 16);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int d = 0; d < ESTEP_DEPTH; ++d) ring_issue(first + d * stride, d);
  int slot = 0;
  const int rot = (threadIdx.x >> 3) & (VEC - 1);
  for (long n0 = first; n0 < end; n0 += stride) {
    T* sb = reinterpret_cast<T*>(s_ring + (size_t)slot * NPL * ESTEP_THREADS + threadIdx.x);
    cp_async_wait<ESTEP_DEPTH - 1>();
    if (n0 + VEC > end) {  // frames beyond the end of the row: zero inputs (own slot, no barrier)
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (n0 + e >= end) {
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) sb[pl * ESTEP_THREADS * VEC + e] = (T)0;
          ++ndead;
        }
    }
    unsigned cmask = 0;  // steps whose bin has the determinant clamp active
    float ll_pass = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int es = (e + rot) & (VEC - 1);  // the bin of the vector this step works on
      const double a0r = (double)sb[0 * ESTEP_THREADS * VEC + es];
      const double a0i = (double)sb[1 * ESTEP_THREADS * VEC + es];
      const double a1r = (double)sb[2 * ESTEP_THREADS * VEC + es];
      const double a1i = (double)sb[3 * ESTEP_THREADS * VEC + es];
      T vt[J];
      double vj[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        vt[j] = sb[(4 + j) * ESTEP_THREADS * VEC + es];
        vj[j] = (double)vt[j];
      }
      // Sigma_x = sum_j v_j R_j + s2 I (audioModel.py:613-654)
      double s00 = s2, s11 = s2, s01r = 0.0, s01i = 0.0;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        s00 += vj[j] * s_coef[4 * j + 0];
        s11 += vj[j] * s_coef[4 * j + 1];
        s01r += vj[j] * s_coef[4 * j + 2];
        if (!RR) s01i += vj[j] * s_coef[4 * j + 3];
      }
      // det as the reference forms it, with its clamp (signalTools.py:183-188: for
      // |det| < eps, sign(det + eps) max(|det|, eps) = +eps)
      const double det_raw = RR ? s00 * s11 - s01r * s01r : s00 * s11 - s01r * s01r - s01i * s01i;
      const bool clamped = fabs(det_raw) < kEps;
      cmask |= clamped ? (1u << e) : 0u;
      const double det = clamped ? kEps : det_raw;
      const double idet = RR ? fast_rcp_h(det) : fast_rcp(det);
      // y' = adj(Sigma) x (y = y' / det) and P = y' y'^H do not wait for the 1/det chain
      const double y0r = RR ? s11 * a0r - s01r * a1r : s11 * a0r - s01r * a1r + s01i * a1i;
      const double y0i = RR ? s11 * a0i - s01r * a1i : s11 * a0i - s01r * a1i - s01i * a1r;
      const double y1r = RR ? s00 * a1r - s01r * a0r : s00 * a1r - s01r * a0r - s01i * a0i;
      const double y1i = RR ? s00 * a1i - s01r * a0i : s00 * a1i - s01r * a0i + s01i * a0r;
      const double p00 = y0r * y0r + y0i * y0i;
      const double p11 = y1r * y1r + y1i * y1i;
      const double p01r = y0r * y1r + y0i * y1i;
      const double p01i = RR ? 0.0 : y0i * y1r - y0r * y1i;
      // log-likelihood integrand log(det*pi) + x^H Sigma^-1 x (audioModel.py:660-664).  Frames
      // beyond the end of the row have x = 0, v = 0: their term log(det0 pi) is counted in
      // `ndead` and taken out after the loop -- no branch in the unrolled body
      const double quad = (a0r * y0r + a0i * y0i + a1r * y1r + a1i * y1i) * idet;
      if (LLP) {  // the float logarithms of a pass are summed in float32
        ll_pass += __logf((float)det);
        acc_ll += quad;
      } else {
        acc_ll += ll_term(det) + quad;
      }
      // M = y y^H - Sigma^-1 = (P / det - adj(Sigma)) / det
      const double m00 = fma(idet, p00, -s11) * idet;
      const double m11 = fma(idet, p11, -s00) * idet;
      const double m01r = fma(idet, p01r, s01r) * idet;
      const double m01i = RR ? 0.0 : fma(idet, p01i, s01i) * idet;
      // posterior source power (audioModel.py:727-729, :408-414)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const T q = RR ? (T)(s_coef2[4 * j + 0] * m00 + s_coef2[4 * j + 1] * m11 +
                             s_coef2[4 * j + 2] * m01r)
                       : (T)(s_coef2[4 * j + 0] * m00 + s_coef2[4 * j + 1] * m11 +
                             s_coef2[4 * j + 2] * m01r + s_coef2[4 * j + 3] * m01i);
        so[j * ESTEP_THREADS * VEC + es] = pf_abs(vt[j] + vt[j] * vt[j] * (q * invrank[j]));
      }
      // S_jk += v_j v_k M ; Z_j += v_j M ; sv_j += v_j
      {
        int p = 0;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
          for (int k = j; k < J; ++k) {
            const double pr = vj[j] * vj[k];
            mom[4 * p + 0] += pr * m00;
            mom[4 * p + 1] += pr * m11;
            mom[4 * p + 2] += pr * m01r;
            if (!RR) mom[4 * p + 3] += pr * m01i;
            ++p;
          }
#pragma unroll
        for (int j = 0; j < J; ++j) {
          mom[4 * NP + 4 * j + 0] += vj[j] * m00;
          mom[4 * NP + 4 * j + 1] += vj[j] * m11;
          mom[4 * NP + 4 * j + 2] += vj[j] * m01r;
          if (!RR) mom[4 * NP + 4 * j + 3] += vj[j] * m01i;
          mom[4 * NP + 4 * J + j] += vj[j];
        }
      }
    }
    if (LLP) acc_ll += (double)(ll_pass + (float)VEC * (float)kLogPi);
#pragma unroll
    for (int j = 0; j < J; ++j)
      *reinterpret_cast<VT*>(hatW + j * plane + row + n0) = s_outv[j * ESTEP_THREADS + threadIdx.x];
    if (cmask != 0) {  // rare slow path, kept out of the unrolled body (no scheduling barrier there)
      any_clamped = true;
#pragma unroll 1
      for (int e = 0; e < VEC; ++e)
        if ((cmask >> e) & 1u) clamp_correction(sb, (e + rot) & (VEC - 1));
    }
    ring_issue(n0 + ESTEP_DEPTH * stride, slot);  // refill the slot just consumed
    slot = slot + 1 == ESTEP_DEPTH ? 0 : slot + 1;
  }
  cp_async_wait<0>();
  {  // the padding frames' log(det0 pi), formed exactly as in the loop (x = 0, v = 0)
    const double det0 = s2 * s2;
    acc_ll -= (double)ndead * ll_term(fabs(det0) < kEps ? kEps : det0);
  }

  // fixed-order block reduction in double (H8: deterministic, no atomics)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NM; ++i) {
    const double d = warp_sum(mom[i]);
    if (lane == 0) s_red[warp][i] = d;
  }
  {
    const double d = warp_sum(acc_ll);
    if (lane == 0) s_red[warp][NA - 1] = d;
  }
  if (__syncthreads_or(any_clamped)) {
    for (int i = 0; i < 8 * J; ++i) {
      const double d = warp_sum(s_corr[i * ESTEP_THREADS + threadIdx.x]);
      if (lane == 0) s_red[warp][NM + i] = d;
    }
  } else if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8 * J; ++i) s_red[warp][NM + i] = 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NA; i += ESTEP_THREADS) {
    double d = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < ESTEP_THREADS / 32; ++w2) d += s_red[w2][i];
    partial[((size_t)f * nsplit + split) * NA + i] = d;
  }
}

// ---- per-frequency contraction with the mixing vectors -------------------------
// hat_Rss: complex128 [F][R][R], hat_Rxs: complex128 [F][2][R], ll_f: [F]
// partial: [F][nsplit][ S (4 per pair) | Z (4 per source) | sv (J) | clamp corr. (8 per source) | ll ]
__global__ void estep_finalize_kernel(const double* __restrict__ partial,
                                      const double2* __restrict__ A,
                                      const double* __restrict__ coef,
                                      const double* __restrict__ noise, SubMap map, int R, int J,
                                      int F, long N, int nsplit, double2* __restrict__ hat_Rss,
                                      double2* __restrict__ hat_Rxs, double* __restrict__ ll_f,
                                      int real_only) {
  const int f = blockIdx.x;
  const int NA = nacc(J), NP = npairs(J), NC = ncoef(J);
  __shared__ double s_acc[nacc(MAXJ)];
  __shared__ double s_T[MAXJ][8];  // T_j = sum_n v_j x y^H, [c][c2] complex, row-major
  __shared__ double2 s_a[MAXR][2];
  for (int i = threadIdx.x; i < NA; i += blockDim.x) {
    double d = 0.0;
    for (int s = 0; s < nsplit; ++s) d += partial[((size_t)f * nsplit + s) * NA + i];
    s_acc[i] = d;
  }
  for (int i = threadIdx.x; i < 2 * R; i += blockDim.x)
    s_a[i >> 1][i & 1] = A[((size_t)(i >> 1) * 2 + (i & 1)) * F + f];
  __syncthreads();
  const double invN = 1.0 / (double)N;
  if (threadIdx.x == 0) ll_f[f] = s_acc[NA - 1];
  // T_j = sv_j I + s2 Z_j + sum_l R_l S_lj + (clamp corrections); x y^H = Sigma M + I
  if (threadIdx.x < J) {
    const int j = threadIdx.x;
    const double s2 = noise[f];
    const double* z = s_acc + 4 * NP + 4 * j;
    const double sv = s_acc[4 * NP + 4 * J + j];
    double t[8] = {sv + s2 * z[0], 0.0, s2 * z[2], s2 * z[3], s2 * z[2], -s2 * z[3],
                   sv + s2 * z[1], 0.0};
    for (int l = 0; l < J; ++l) {
      const int a = l < j ? l : j, b = l < j ? j : l;
      const int p = a * J - a * (a - 1) / 2 + (b - a);  // pair (a <= b), row-major upper triangle
      const double m00 = s_acc[4 * p + 0], m11 = s_acc[4 * p + 1];
      const double mr = s_acc[4 * p + 2], mi = s_acc[4 * p + 3];
      const double* c = coef + (size_t)f * NC + 4 * l;  // R_l: R00, R11, Re R01, Im R01
      const double r00 = c[0], r11 = c[1], rr = c[2], ri = c[3];
      // (R S)00 = r00 m00 + r01 conj(m01); (R S)01 = r00 m01 + r01 m11
      t[0] += r00 * m00 + rr * mr + ri * mi;
      t[1] += ri * mr - rr * mi;
      t[2] += r00 * mr + rr * m11;
      t[3] += r00 * mi + ri * m11;
      // (R S)10 = conj(r01) m00 + r11 conj(m01); (R S)11 = conj(r01) m01 + r11 m11
      t[4] += rr * m00 + r11 * mr;
      t[5] += -ri * m00 - r11 * mi;
      t[6] += rr * mr + ri * mi + r11 * m11;
      t[7] += rr * mi - ri * mr;
    }
    for (int i = 0; i < 8; ++i) s_T[j][i] = t[i] + s_acc[4 * NP + 5 * J + 8 * j + i];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * R; idx += blockDim.x) {
    const int r1 = idx / R, r2 = idx % R;
    if (r1 > r2) continue;
    int j1 = map.src_of_sub[r1], j2 = map.src_of_sub[r2];
    if (j1 > j2) { int t = j1; j1 = j2; j2 = t; }
    // pair index of (j1 <= j2) in row-major upper-triangle order
    const int p = j1 * J - j1 * (j1 - 1) / 2 + (j2 - j1);
    const double m00 = s_acc[4 * p + 0], m11 = s_acc[4 * p + 1];
    const double mr = s_acc[4 * p + 2], mi = s_acc[4 * p + 3];
    // h = a_r1^H S a_r2 with S = [[m00, m01],[conj(m01), m11]], m01 = mr + i mi
    const double2 a0 = s_a[r1][0], a1 = s_a[r1][1], b0 = s_a[r2][0], b1 = s_a[r2][1];
    // t0 = S[0,:] b = m00 b0 + m01 b1 ; t1 = conj(m01) b0 + m11 b1
    const double t0r = m00 * b0.x + mr * b1.x - mi * b1.y;
    const double t0i = m00 * b0.y + mr * b1.y + mi * b1.x;
    const double t1r = mr * b0.x + mi * b0.y + m11 * b1.x;
    const double t1i = mr * b0.y - mi * b0.x + m11 * b1.y;
    // h = conj(a0) t0 + conj(a1) t1
    double hr = a0.x * t0r + a0.y * t0i + a1.x * t1r + a1.y * t1i;
    double hi = a0.x * t0i - a0.y * t0r + a1.x * t1i - a1.y * t1r;
    hr *= invN; hi *= invN;
    if (r1 == r2) {
      hr += s_acc[4 * NP + 4 * J + map.src_of_sub[r1]] * invN;
      hi = 0.0;  // Hermitian symmetrisation (audioModel.py:733-740)
    }
    if (real_only) hi = 0.0;  // (the imaginary moments were not accumulated)
    hat_Rss[((size_t)f * R + r1) * R + r2] = make_double2(hr, hi);
    hat_Rss[((size_t)f * R + r2) * R + r1] = make_double2(hr, -hi);
  }
  for (int idx = threadIdx.x; idx < 2 * R; idx += blockDim.x) {
    const int c = idx / R, r = idx % R;
    const int j = map.src_of_sub[r];
    const double* t = s_T[j] + 4 * c;  // T_j[c][0], T_j[c][1]
    const double2 b0 = s_a[r][0], b1 = s_a[r][1];
    const double hr = t[0] * b0.x - t[1] * b0.y + t[2] * b1.x - t[3] * b1.y;
    const double hi = t[0] * b0.y + t[1] * b0.x + t[2] * b1.y + t[3] * b1.x;
    hat_Rxs[((size_t)f * 2 + c) * R + r] = make_double2(hr * invN, real_only ? 0.0 : hi * invN);
  }
}

// ---- Wiener filter (separation) ---------------------------------------------------
// Replaces compute_sigma_comp_2d / compute_inv_sigma_mix_2d / compute_Wiener_gain_2d and
// the gain application of FASST.separate_comps (audioModel.py:1088-1217, :1327-1467):
// Y_g = (sum_{j in g} v_j R_j) Sigma^-1 x, written as planes Y[g][re0, im0, re1, im1].
struct GroupMap {
  int group_of_src[MAXJ];  // output group of each spatial component, -1 = not written
};

template <typename T, typename C, int J>
__global__ void __launch_bounds__(256)
wiener_stereo_kernel(const T* __restrict__ X, const T* __restrict__ V,
                     const double* __restrict__ coef, const double* __restrict__ noise,
                     GroupMap gm, int ngroups, T* __restrict__ Y, int F, long N, long ld) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NC = ncoef(J);
  constexpr int NP = npairs(J);
  const int f = blockIdx.y;
  __shared__ C s_coef[NC];
  __shared__ T s_dcoef[NP];
  if (threadIdx.x < NC) {
    const double c = coef[(size_t)f * NC + threadIdx.x];
    s_coef[threadIdx.x] = (C)c;
    if (threadIdx.x >= 4 * J) s_dcoef[threadIdx.x - 4 * J] = (T)c;
  }
  __syncthreads();
  const C s2 = (C)noise[f];
  const long plane = (long)F * ld;
  const long row = (long)f * ld;
  const long n0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (n0 >= N) return;
  T x0r[VEC], x0i[VEC], x1r[VEC], x1i[VEC], v[J][VEC];
  load_vec<T>(X + 0 * plane + row + n0, x0r);
  load_vec<T>(X + 1 * plane + row + n0, x0i);
  load_vec<T>(X + 2 * plane + row + n0, x1r);
  load_vec<T>(X + 3 * plane + row + n0, x1i);
#pragma unroll
  for (int j = 0; j < J; ++j) load_vec<T>(V + j * plane + row + n0, v[j]);
  C y0r[VEC], y0i[VEC], y1r[VEC], y1i[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    C vj[J], i00, i11, i01r, i01i;
    T vt[J], pr[NP], det;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      vt[j] = v[j][e];
      vj[j] = (C)vt[j];
    }
    sigma_inverse<C, T, J>(vj, vt, s_coef, s_dcoef, s2, pr, det, i00, i11, i01r, i01i);
    const C a0r = (C)x0r[e], a0i = (C)x0i[e], a1r = (C)x1r[e], a1i = (C)x1i[e];
    y0r[e] = i00 * a0r + i01r * a1r - i01i * a1i;
    y0i[e] = i00 * a0i + i01r * a1i + i01i * a1r;
    y1r[e] = i01r * a0r + i01i * a0i + i11 * a1r;
    y1i[e] = i01r * a0i - i01i * a0r + i11 * a1i;
  }
  for (int g = 0; g < ngroups; ++g) {
    T o0r[VEC], o0i[VEC], o1r[VEC], o1i[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      // Sigma_g = sum_{j in g} v_j R_j
      C g00 = (C)0, g11 = (C)0, g01r = (C)0, g01i = (C)0;
#pragma unroll
      for (int j = 0; j < J; ++j)
        if (gm.group_of_src[j] == g) {
          const C vv = (C)v[j][e];
          g00 += vv * s_coef[4 * j + 0];
          g11 += vv * s_coef[4 * j + 1];
          g01r += vv * s_coef[4 * j + 2];
          g01i += vv * s_coef[4 * j + 3];
        }
      // out = Sigma_g y
      o0r[e] = (T)(g00 * y0r[e] + g01r * y1r[e] - g01i * y1i[e]);
      o0i[e] = (T)(g00 * y0i[e] + g01r * y1i[e] + g01i * y1r[e]);
      o1r[e] = (T)(g01r * y0r[e] + g01i * y0i[e] + g11 * y1r[e]);
      o1i[e] = (T)(g01r * y0i[e] - g01i * y0r[e] + g11 * y1i[e]);
      if (n0 + e >= N) o0r[e] = o0i[e] = o1r[e] = o1i[e] = (T)0;
    }
    T* out = Y + (size_t)g * 4 * plane + row + n0;
    store_vec<T>(out + 0 * plane, o0r);
    store_vec<T>(out + 1 * plane, o0i);
    store_vec<T>(out + 2 * plane, o1r);
    store_vec<T>(out + 3 * plane, o1i);
  }
}

template <typename T, typename C, int J>
static int launch_wiener(const void* X, const void* V, const double* coef, const double* noise,
                         const GroupMap& gm, int ngroups, void* Y, int F, long N, long ld,
                         cudaStream_t st) {
  constexpr int VEC = VecOf<T>::N;
  dim3 grid(ceil_div(N, 256L * VEC), F);
  wiener_stereo_kernel<T, C, J><<<grid, 256, 0, st>>>((const T*)X, (const T*)V, coef, noise, gm,
                                                     ngroups, (T*)Y, F, N, ld);
  return check_launch("wiener_stereo_kernel");
}

template <typename T, typename C>
static int dispatch_wiener(int J, const void* X, const void* V, const double* coef,
                           const double* noise, const GroupMap& gm, int ngroups, void* Y, int F,
                           long N, long ld, cudaStream_t st) {
  switch (J) {
    case 1: return launch_wiener<T, C, 1>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 2: return launch_wiener<T, C, 2>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 3: return launch_wiener<T, C, 3>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 4: return launch_wiener<T, C, 4>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 5: return launch_wiener<T, C, 5>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 6: return launch_wiener<T, C, 6>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
  }
  set_error("pf_wiener_stereo: J=%d spatial components not supported (1..%d)", J, MAXJ);
  return PF_ERR_UNSUPPORTED;
}

static size_t estep_smem_bytes(int J) {
  return (size_t)(ESTEP_DEPTH * (4 + J) + J) * ESTEP_THREADS * 16 + (size_t)8 * J * ESTEP_THREADS * 8;
}

template <typename T, int J, bool RR>
static int launch_estep(const void* X, const void* V, const double* coef, const double* noise,
                        const SubMap& map, void* hatW, double* partial, int F, long N, long ld,
                        int nsplit, cudaStream_t st) {
  dim3 grid(nsplit, F);
  const size_t smem = estep_smem_bytes(J);
  cudaError_t e = cudaFuncSetAttribute(estep_stereo_kernel<T, J, RR>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("estep_stereo_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  estep_stereo_kernel<T, J, RR><<<grid, ESTEP_THREADS, smem, st>>>(
      (const T*)X, (const T*)V, coef, noise, map, (T*)hatW, partial, F, N, ld, nsplit);
  return check_launch("estep_stereo_kernel");
}

template <typename T, bool RR>
static int dispatch_estep(int J, const void* X, const void* V, const double* coef,
                          const double* noise, const SubMap& map, void* hatW, double* partial,
                          int F, long N, long ld, int nsplit, cudaStream_t st) {
  switch (J) {
    case 1: return launch_estep<T, 1, RR>(X, V, coef, noise, map, hatW, partial, F, N, ld, nsplit, st);
    case 2: return launch_estep<T, 2, RR>(X, V, coef, noise, map, hatW, partial, F, N, ld, nsplit, st);
    case 3: return launch_estep<T, 3, RR>(X, V, coef, noise, map, hatW, partial, F, N, ld, nsplit, st);
    case 4: return launch_estep<T, 4, RR>(X, V, coef, noise, map, hatW, partial, F, N, ld, nsplit, st);
    case 5: return launch_estep<T, 5, RR>(X, V, coef, noise, map, hatW, partial, F, N, ld, nsplit, st);
    case 6: return launch_estep<T, 6, RR>(X, V, coef, noise, map, hatW, partial, F, N, ld, nsplit, st);
  }
  set_error("pf_estep_stereo: J=%d spatial components not supported (1..%d)", J, MAXJ);
  return PF_ERR_UNSUPPORTED;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_estep_plan(int J, int64_t N, int dtype, int64_t* chunk, int* nsplit,
                             int64_t* workspace_bytes, int F) {
  PF_REQUIRE(J >= 1 && J <= MAXJ, "pf_estep_plan: J=%d out of range", J);
  const long vec = dtype == PF_F64 ? 2 : 4;
  const long pass = ESTEP_THREADS * vec;
  // aim for ~64 passes per CTA so that the start and the end-of-CTA reduction (one CTA costs
  // about one pass on top of its passes) are amortised, while keeping at least ~4 CTAs per SM
  // in flight on a 148-SM part
  long passes = (N + pass - 1) / pass;
  long per_cta = 64;
  const char* e = getenv("PYFASST_ESTEP_PASSES");  // tuning: passes per CTA (a power of two)
  if (e != nullptr && atoi(e) >= 1 && atoi(e) <= 1024) per_cta = atoi(e);
  long want_ctas = 148L * 8;
  while (per_cta > 1 && (long)F * ((passes + per_cta - 1) / per_cta) < want_ctas) per_cta /= 2;
  long c = per_cta * pass;
  int ns = (int)((N + c - 1) / c);
  if (ns < 1) ns = 1;
  *chunk = c;
  *nsplit = ns;
  *workspace_bytes = (int64_t)F * ns * nacc(J) * sizeof(double) + (int64_t)F * ncoef(J) * 8;
  return PF_OK;
}

// ---- optional CUDA-event bracket around the per-bin kernel alone (bench.py's roofline) -----------
// pf_estep_timing(1) makes every following pf_estep_* call record an event pair on ITS stream right
// before and after the per-bin kernel (not the two small per-frequency kernels around it);
// pf_estep_timing_read() synchronises on them and returns the sum of the elapsed times.
namespace pf {
static bool g_estep_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_estep_events;
void estep_timing_begin(cudaStream_t st) {
  if (!g_estep_timing) return;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, st);
  g_estep_events.push_back(std::make_pair(a, b));
}
void estep_timing_end(cudaStream_t st) {
  if (!g_estep_timing) return;
  cudaEventRecord(g_estep_events.back().second, st);
}
}  // namespace pf

extern "C" int pf_estep_timing(int enable) {
  pf::g_estep_timing = enable != 0;
  return PF_OK;
}

extern "C" int pf_estep_timing_read(double* total_ms, int* launches) {
  double sum = 0.0;
  int n = 0;
  for (auto& ev : pf::g_estep_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
      sum += ms;
      ++n;
    }
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  pf::g_estep_events.clear();
  *total_ms = sum;
  *launches = n;
  return PF_OK;
}

static int estep_stereo_impl(const void* X, const void* V, const void* A,
                             const int* src_of_sub, int R, int J, const double* noise_psd,
                             int F, int64_t N, int64_t ld, void* hatW, void* hat_Rss,
                             void* hat_Rxs, double* ll_f, void* workspace,
                             int64_t workspace_bytes, int64_t N_norm, int dtype, void* stream,
                             bool real_mixing) {
  if (J > MAXJ || R > MAXR) {
    set_error("pf_estep_stereo: J=%d spatial components / R=%d sub-sources not supported "
              "(max %d / %d)", J, R, MAXJ, MAXR);
    return PF_ERR_UNSUPPORTED;
  }
  PF_REQUIRE(J >= 1 && R >= J, "pf_estep_stereo: J=%d R=%d", J, R);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_estep_stereo: bad dtype %d", dtype);
  PF_REQUIRE(ld >= N && ld % 4 == 0, "pf_estep_stereo: ld=%ld must be >= N and a multiple of 4",
             (long)ld);
  PF_REQUIRE(F > 0 && N > 0, "pf_estep_stereo: empty problem F=%d N=%ld", F, (long)N);
  SubMap map;
  int count[MAXJ] = {0};
  for (int r = 0; r < R; ++r) {
    PF_REQUIRE(src_of_sub[r] >= 0 && src_of_sub[r] < J, "pf_estep_stereo: src_of_sub[%d]=%d", r,
               src_of_sub[r]);
    map.src_of_sub[r] = src_of_sub[r];
    count[src_of_sub[r]]++;
  }
  for (int j = 0; j < J; ++j) {
    PF_REQUIRE(count[j] > 0, "pf_estep_stereo: spatial component %d has rank 0", j);
    map.invrank[j] = 1.0 / count[j];
  }
  int64_t chunk, need;
  int nsplit;
  pf_estep_plan(J, N, dtype, &chunk, &nsplit, &need, F);
  PF_REQUIRE(workspace_bytes >= need, "pf_estep_stereo: workspace %ld < %ld bytes",
             (long)workspace_bytes, (long)need);
  cudaStream_t st = as_stream(stream);
  double* partial = (double*)workspace;
  double* coef = partial + (size_t)F * nsplit * nacc(J);
  spat_coef_kernel<<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, F, coef);
  int rc = check_launch("spat_coef_kernel");
  if (rc) return rc;
  // the per-bin algebra and the moment sums run in float64 whatever the plane type
  estep_timing_begin(st);
  if (dtype == PF_F32)
    rc = real_mixing
             ? dispatch_estep<float, true>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, nsplit, st)
             : dispatch_estep<float, false>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, nsplit, st);
  else
    rc = real_mixing
             ? dispatch_estep<double, true>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, nsplit, st)
             : dispatch_estep<double, false>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, nsplit, st);
  estep_timing_end(st);
  if (rc) return rc;
  // hat_Rss / hat_Rxs are means over N_norm frames: the local N, or the length of the whole
  // mixture when the frames are sharded over several GPUs (the partial means are then summed)
  estep_finalize_kernel<<<F, 64, 0, st>>>(partial, (const double2*)A, coef, noise_psd, map, R, J,
                                          F, N_norm > 0 ? N_norm : N, nsplit, (double2*)hat_Rss,
                                          (double2*)hat_Rxs, ll_f, real_mixing ? 1 : 0);
  return check_launch("estep_finalize_kernel");
}

extern "C" int pf_estep_stereo(const void* X, const void* V, const void* A,
                               const int* src_of_sub, int R, int J, const double* noise_psd,
                               int F, int64_t N, int64_t ld, void* hatW, void* hat_Rss,
                               void* hat_Rxs, double* ll_f, void* workspace,
                               int64_t workspace_bytes, int64_t N_norm, int dtype, void* stream) {
  return estep_stereo_impl(X, V, A, src_of_sub, R, J, noise_psd, F, N, ld, hatW, hat_Rss, hat_Rxs,
                           ll_f, workspace, workspace_bytes, N_norm, dtype, stream, false);
}

extern "C" int pf_estep_stereo_inst(const void* X, const void* V, const void* A,
                                    const int* src_of_sub, int R, int J, const double* noise_psd,
                                    int F, int64_t N, int64_t ld, void* hatW, void* hat_Rss,
                                    void* hat_Rxs, double* ll_f, void* workspace,
                                    int64_t workspace_bytes, int64_t N_norm, int dtype,
                                    void* stream) {
  return estep_stereo_impl(X, V, A, src_of_sub, R, J, noise_psd, F, N, ld, hatW, hat_Rss, hat_Rxs,
                           ll_f, workspace, workspace_bytes, N_norm, dtype, stream, true);
}

extern "C" int pf_wiener_stereo(const void* X, const void* V, const void* A,
                                const int* src_of_sub, int R, int J, const double* noise_psd,
                                const int* group_of_src, int ngroups, int F, int64_t N,
                                int64_t ld, void* Y, void* workspace, int64_t workspace_bytes,
                                int dtype, void* stream) {
  PF_REQUIRE(J >= 1 && J <= MAXJ, "pf_wiener_stereo: J=%d out of range (1..%d)", J, MAXJ);
  PF_REQUIRE(R >= J && R <= MAXR, "pf_wiener_stereo: R=%d out of range (J..%d)", R, MAXR);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_wiener_stereo: bad dtype %d", dtype);
  PF_REQUIRE(ld >= N && ld % 4 == 0, "pf_wiener_stereo: ld=%ld must be >= N and a multiple of 4",
             (long)ld);
  PF_REQUIRE(F > 0 && N > 0 && ngroups >= 1 && ngroups <= J, "pf_wiener_stereo: F=%d N=%ld ngroups=%d",
             F, (long)N, ngroups);
  PF_REQUIRE(workspace_bytes >= (int64_t)F * ncoef(J) * 8, "pf_wiener_stereo: workspace too small");
  SubMap map;
  for (int r = 0; r < R; ++r) {
    PF_REQUIRE(src_of_sub[r] >= 0 && src_of_sub[r] < J, "pf_wiener_stereo: src_of_sub[%d]=%d", r,
               src_of_sub[r]);
    map.src_of_sub[r] = src_of_sub[r];
  }
  for (int j = 0; j < J; ++j) map.invrank[j] = 1.0;
  GroupMap gm;
  for (int j = 0; j < MAXJ; ++j) gm.group_of_src[j] = -1;
  for (int j = 0; j < J; ++j) {
    PF_REQUIRE(group_of_src[j] >= -1 && group_of_src[j] < ngroups,
               "pf_wiener_stereo: group_of_src[%d]=%d", j, group_of_src[j]);
    gm.group_of_src[j] = group_of_src[j];
  }
  cudaStream_t st = as_stream(stream);
  double* coef = (double*)workspace;
  spat_coef_kernel<<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, F, coef);
  int rc = check_launch("spat_coef_kernel");
  if (rc) return rc;
  if (dtype == PF_F32)
    return dispatch_wiener<float, double>(J, X, V, coef, noise_psd, gm, ngroups, Y, F, N, ld, st);
  return dispatch_wiener<double, double>(J, X, V, coef, noise_psd, gm, ngroups, Y, F, N, ld, st);
}
