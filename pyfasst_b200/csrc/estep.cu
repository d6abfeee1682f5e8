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
#include "estep.cuh"

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
// T: storage type of the planes; C: type of the per-bin algebra.  With T = float the
// algebra still runs in double: Sigma^-1 has entries ~ 1/noise, and the posterior power
// a^H (y y^H - Sigma^-1) a cancels them down by the condition number of Sigma, which in
// float32 costs eps*cond(Sigma) ~ 1e-2 at 50-60 dB bins.  The per-frequency moment sums
// are accumulated in T (their rounding is random and averages out over frames).
// OPT bit 0: packed float32 moment accumulation (FFMA2); bit 1: hardware float->double
// conversion of the loaded values.  MINB: CTAs per SM the register allocation aims for.
template <typename T, typename C, int J, int OPT, int MINB>
#ifdef PF_ESTEP_MAXNREG
__global__ void __maxnreg__(PF_ESTEP_MAXNREG)
#else
__global__ void __launch_bounds__(ESTEP_THREADS, MINB)
#endif
estep_stereo_kernel(const T* __restrict__ X, const T* __restrict__ V,
                    const double* __restrict__ coef, const double* __restrict__ noise,
                    SubMap map, T* __restrict__ hatW, double* __restrict__ partial, int F,
                    long N, long ld, long chunk, int nsplit) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NP = npairs(J);
  constexpr int NA = nacc(J);
  constexpr int NC = ncoef(J);
  constexpr float kLogPi = 1.1447298858494002f;  // log(pi): Q4, log(det*pi)

  const int f = blockIdx.y;
  const int split = blockIdx.x;
  __shared__ C s_coef[NC];
  __shared__ T s_dcoef[NP];  // the mixed discriminants in the storage type (determinant)
  __shared__ T s_fcoef[4 * J];  // R_j in the storage type (OPT bit 4)
  __shared__ double s_red[ESTEP_THREADS / 32][NA];
  if (threadIdx.x < NC) {
    const double c = coef[(size_t)f * NC + threadIdx.x];
    s_coef[threadIdx.x] = (C)c;
    if (threadIdx.x >= 4 * J) s_dcoef[threadIdx.x - 4 * J] = (T)c;
    else s_fcoef[threadIdx.x] = (T)c;
  }
  __syncthreads();
  const C s2 = (C)noise[f];
  T invrank[J];
#pragma unroll
  for (int j = 0; j < J; ++j) invrank[j] = (T)map.invrank[j];

  constexpr bool kPack = (OPT & 1) != 0 && sizeof(T) == 4;
  constexpr bool kHwCvt = (OPT & 2) != 0;
  // OPT bit 4: fewer float<->double conversions (XU pipe, 16 results/clk/SM: the busiest pipe of
  // this kernel).  Sigma is ALSO formed in the storage type (16 FFMA): the determinant needs no
  // narrowed s00/s11, and the float copy of M = y y^H - Sigma^-1 that feeds the moment sums is
  // formed from the narrowed y and the float Sigma^-1 instead of narrowing the four entries of
  // the float64 M (which still feeds tr(M R_j), where the cancellation happens).  The
  // log-likelihood integrand is summed in float over the VEC bins of a pass and widened once.
  constexpr bool kLean = (OPT & 16) != 0 && sizeof(T) == 4 && sizeof(C) == 8;
  const T s2t = (T)noise[f];
  Moments<T, J, kPack> mom;
  mom.clear();
  double acc_ll = 0.0;

  const long plane = (long)F * ld;
  const long row = (long)f * ld;
  // OPT bit 5: the splits of a row take its passes in turn (split s works on passes s, s + nsplit,
  // ...) instead of one contiguous run of frames each: the CTAs of a row, which are launched
  // together, then read and write ONE contiguous region of every plane at any moment.
  constexpr bool kInterleave = (OPT & 32) != 0;
  const long begin = kInterleave ? (long)split * ESTEP_THREADS * VEC : (long)split * chunk;
  long end = kInterleave ? N : begin + chunk;
  if (end > N) end = N;

  // Loads run ahead of the arithmetic: with only two CTAs resident per SM (the moment
  // accumulators pin ~80 registers per thread) the ~1 us HBM latency is otherwise exposed.
  //  * OPT bit 2: cp.async ring in shared memory, ESTEP_DEPTH passes ahead.  Every thread copies
  //    and later reads back ITS OWN 16 bytes per plane, so no barrier is needed -- only
  //    cp.async.wait_group on the thread's own groups (one group per pass, possibly empty).
  //  * otherwise: register double buffering, one pass ahead (float32 planes only).
  constexpr bool kRing = (OPT & 4) != 0 && sizeof(T) == 4;
  //  * OPT bit 3 (with the ring): the inputs of a pass STAY in the ring slot -- every bin reads
  //    its 4 + J scalars from shared memory when its turn comes and writes hatW back over V --
  //    so that no float4 input / output / prefetch registers are live across the per-bin
  //    algebra and three CTAs fit on an SM.  Lane l visits its four bins rotated by l / 8:
  //    the 4-byte accesses at a 16-byte stride are then bank-conflict free.
  constexpr bool kSmemIO = kRing && (OPT & 8) != 0;
  constexpr bool kPrefetch = sizeof(T) == 4 && !kRing;
  constexpr int NPL = 4 + J;
  extern __shared__ __align__(16) unsigned char s_ring_raw[];
  float4* s_ring = reinterpret_cast<float4*>(s_ring_raw);  // [ESTEP_DEPTH][NPL][ESTEP_THREADS]
  const long stride = (long)ESTEP_THREADS * VEC * (kInterleave ? nsplit : 1);
  const long first = begin + (long)threadIdx.x * VEC;
  auto ring_issue = [&](long n, int slot) {
    if (n < end) {
      float4* dst = s_ring + (size_t)slot * NPL * ESTEP_THREADS + threadIdx.x;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl)
        cp_async16(dst + pl * ESTEP_THREADS, X + pl * plane + row + n, 16);
#pragma unroll
      for (int j = 0; j < J; ++j)
        cp_async16(dst + (4 + j) * ESTEP_THREADS, V + j * plane + row + n, 16);
    }
    cp_async_commit();
  };
  // OPT bit 6: streaming cache hints (ld.global.cs / st.global.cs): every plane is touched once
  constexpr bool kStream = (OPT & 64) != 0 && sizeof(T) == 4;
  T nx[4][VEC], nv[J][VEC];
  auto ld_plane = [&](const T* p, T (&out)[VEC]) {
    if (kStream) {
      const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
      out[0] = (T)q.x; out[1] = (T)q.y;
      if (VEC == 4) { out[VEC - 2] = (T)q.z; out[VEC - 1] = (T)q.w; }
    } else {
      load_vec<T>(p, out);
    }
  };
  auto issue_loads = [&](long n) {
    ld_plane(X + 0 * plane + row + n, nx[0]);
    ld_plane(X + 1 * plane + row + n, nx[1]);
    ld_plane(X + 2 * plane + row + n, nx[2]);
    ld_plane(X + 3 * plane + row + n, nx[3]);
#pragma unroll
    for (int j = 0; j < J; ++j) ld_plane(V + j * plane + row + n, nv[j]);
  };
  if (kRing) {
#pragma unroll
    for (int d = 0; d < ESTEP_DEPTH; ++d) ring_issue(first + d * stride, d);
  }
  if (kPrefetch && first < end) issue_loads(first);
  int slot = 0;
  const int rot = (threadIdx.x >> 3) & 3;
  for (long n0 = first; n0 < end; n0 += stride) {
    T x0r[VEC], x0i[VEC], x1r[VEC], x1i[VEC], v[J][VEC], w[J][VEC];
    float* sb = reinterpret_cast<float*>(s_ring + (size_t)slot * NPL * ESTEP_THREADS + threadIdx.x);
    if (kSmemIO) {
      cp_async_wait<ESTEP_DEPTH - 1>();
      if (n0 + VEC > end) {  // frames beyond the end of the row: zero inputs (own slot, no barrier)
#pragma unroll
        for (int e = 0; e < VEC; ++e)
          if (n0 + e >= end) {
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) sb[pl * ESTEP_THREADS * 4 + e] = 0.f;
          }
      }
    } else if (kRing) {
      cp_async_wait<ESTEP_DEPTH - 1>();
      const float4* src = s_ring + (size_t)slot * NPL * ESTEP_THREADS + threadIdx.x;
      if (sizeof(T) == 4) {  // (the ring only exists for float32 planes)
        float4 q;
        q = src[0 * ESTEP_THREADS]; x0r[0] = q.x; x0r[1] = q.y; x0r[2] = q.z; x0r[3] = q.w;
        q = src[1 * ESTEP_THREADS]; x0i[0] = q.x; x0i[1] = q.y; x0i[2] = q.z; x0i[3] = q.w;
        q = src[2 * ESTEP_THREADS]; x1r[0] = q.x; x1r[1] = q.y; x1r[2] = q.z; x1r[3] = q.w;
        q = src[3 * ESTEP_THREADS]; x1i[0] = q.x; x1i[1] = q.y; x1i[2] = q.z; x1i[3] = q.w;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          q = src[(4 + j) * ESTEP_THREADS];
          v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; v[j][3] = q.w;
        }
      }
      ring_issue(n0 + ESTEP_DEPTH * stride, slot);  // refill the slot just consumed
      slot = slot + 1 == ESTEP_DEPTH ? 0 : slot + 1;
    } else {
      if (!kPrefetch) issue_loads(n0);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        x0r[e] = nx[0][e]; x0i[e] = nx[1][e]; x1r[e] = nx[2][e]; x1i[e] = nx[3][e];
#pragma unroll
        for (int j = 0; j < J; ++j) v[j][e] = nv[j][e];
      }
      if (kPrefetch && n0 + stride < end) issue_loads(n0 + stride);
    }
    // frames beyond the end of the row: zero inputs contribute nothing to the moments and give
    // hatW = 0; only the log-likelihood term is masked below (no branch around the algebra)
    if (!kSmemIO && n0 + VEC > end) {
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (n0 + e >= end) {
          x0r[e] = x0i[e] = x1r[e] = x1i[e] = (T)0;
#pragma unroll
          for (int j = 0; j < J; ++j) v[j][e] = (T)0;
        }
    }

    float pass_ll = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int es = kSmemIO ? ((e + rot) & 3) : e;  // the bin of the vector this step works on
      const bool live = n0 + es < end;
      if (kSmemIO) {
        x0r[e] = (T)sb[0 * ESTEP_THREADS * 4 + es];
        x0i[e] = (T)sb[1 * ESTEP_THREADS * 4 + es];
        x1r[e] = (T)sb[2 * ESTEP_THREADS * 4 + es];
        x1i[e] = (T)sb[3 * ESTEP_THREADS * 4 + es];
#pragma unroll
        for (int j = 0; j < J; ++j) v[j][e] = (T)sb[(4 + j) * ESTEP_THREADS * 4 + es];
      }
      // Sigma_x = sum_j v_j R_j + s2 I and its inverse (audioModel.py:613-654)
      C vj[J], i00, i11, i01r, i01i;
      T vt[J], pr[NP], det;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        vt[j] = v[j][e];
        vj[j] = (C)(kHwCvt ? widen_hw(vt[j]) : widen(vt[j]));
      }
      T f00 = s2t, f11 = s2t, f01r = (T)0, f01i = (T)0, idet_t = (T)0;
      if (kLean) {
        C s00 = s2, s11 = s2, s01r = (C)0, s01i = (C)0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          s00 += vj[j] * s_coef[4 * j + 0];
          s11 += vj[j] * s_coef[4 * j + 1];
          s01r += vj[j] * s_coef[4 * j + 2];
          s01i += vj[j] * s_coef[4 * j + 3];
          f00 += vt[j] * s_fcoef[4 * j + 0];
          f11 += vt[j] * s_fcoef[4 * j + 1];
          f01r += vt[j] * s_fcoef[4 * j + 2];
          f01i += vt[j] * s_fcoef[4 * j + 3];
        }
        det = s2t * (f00 + (f11 - s2t));
        int p = 0;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
          for (int k = j; k < J; ++k) {
            pr[p] = vt[j] * vt[k];
            det += pr[p] * s_dcoef[p];
            ++p;
          }
        det = pf_max(det, (T)1e-10);
        idet_t = fast_rcp(det);
        const C idet = (C)idet_t;
        i00 = s11 * idet;
        i11 = s00 * idet;
        i01r = -s01r * idet;
        i01i = -s01i * idet;
      } else {
        sigma_inverse<C, T, J>(vj, vt, s_coef, s_dcoef, s2, pr, det, i00, i11, i01r, i01i);
      }
      // y = Sigma^-1 x
      const C a0r = (C)(kHwCvt ? widen_hw(x0r[e]) : widen(x0r[e]));
      const C a0i = (C)(kHwCvt ? widen_hw(x0i[e]) : widen(x0i[e]));
      const C a1r = (C)(kHwCvt ? widen_hw(x1r[e]) : widen(x1r[e]));
      const C a1i = (C)(kHwCvt ? widen_hw(x1i[e]) : widen(x1i[e]));
      const C y0r = i00 * a0r + i01r * a1r - i01i * a1i;
      const C y0i = i00 * a0i + i01r * a1i + i01i * a1r;
      const C y1r = i01r * a0r + i01i * a0i + i11 * a1r;
      const C y1i = i01r * a0i - i01i * a0r + i11 * a1i;
      const T z0r = (T)y0r, z0i = (T)y0i, z1r = (T)y1r, z1i = (T)y1i;
      const T b0r = x0r[e], b0i = x0i[e], b1r = x1r[e], b1i = x1i[e];
      // log-likelihood integrand log(det*pi) + x^H Sigma^-1 x (audioModel.py:660-664)
      const T quad = b0r * z0r + b0i * z0i + b1r * z1r + b1i * z1i;
      if (sizeof(T) == 8)
        acc_ll += live ? log((double)det) + 1.1447298858494002 + (double)quad : 0.0;
      else if (kLean)
        pass_ll += live ? __logf((float)det) + kLogPi + (float)quad : 0.f;
      else
        acc_ll += (double)(live ? __logf((float)det) + kLogPi + (float)quad : 0.f);
      // M = y y^H - Sigma^-1
      const C m00 = y0r * y0r + y0i * y0i - i00;
      const C m11 = y1r * y1r + y1i * y1i - i11;
      const C m01r = y0r * y1r + y0i * y1i - i01r;
      const C m01i = y0i * y1r - y0r * y1i - i01i;
      // posterior source power (audioModel.py:727-729, :408-414): tr(M R_j) cancels the
      // ~1/noise entries of M, so it is formed in C; the rest is safe in T
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const T q = (T)(s_coef[4 * j + 0] * m00 + s_coef[4 * j + 1] * m11 +
                        (C)2 * (s_coef[4 * j + 2] * m01r + s_coef[4 * j + 3] * m01i));
        w[j][e] = pf_abs(vt[j] + vt[j] * vt[j] * (q * invrank[j]));
        if (kSmemIO) sb[(4 + j) * ESTEP_THREADS * 4 + es] = (float)w[j][e];
      }
      // S_jk += v_j v_k M ; U = x y^H ; T_j += v_j U ; sv_j += v_j   (accumulated in T)
      T t00, t11, t01r, t01i;
      if (kLean) {
        t00 = z0r * z0r + z0i * z0i - f11 * idet_t;
        t11 = z1r * z1r + z1i * z1i - f00 * idet_t;
        t01r = z0r * z1r + z0i * z1i + f01r * idet_t;
        t01i = z0i * z1r - z0r * z1i + f01i * idet_t;
      } else {
        t00 = (T)m00; t11 = (T)m11; t01r = (T)m01r; t01i = (T)m01i;
      }
      const T u[8] = {b0r * z0r + b0i * z0i, b0i * z0r - b0r * z0i,
                      b0r * z1r + b0i * z1i, b0i * z1r - b0r * z1i,
                      b1r * z0r + b1i * z0i, b1i * z0r - b1r * z0i,
                      b1r * z1r + b1i * z1i, b1i * z1r - b1r * z1i};
      mom.add(pr, vt, t00, t11, t01r, t01i, u);
    }
    if (kLean) acc_ll += (double)pass_ll;
    if (kSmemIO) {
      if (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < J; ++j)
          *reinterpret_cast<float4*>(hatW + j * plane + row + n0) =
              *reinterpret_cast<const float4*>(sb + (4 + j) * ESTEP_THREADS * 4);
      }
      ring_issue(n0 + ESTEP_DEPTH * stride, slot);  // refill the slot just consumed
      slot = slot + 1 == ESTEP_DEPTH ? 0 : slot + 1;
    } else if (kStream) {
#pragma unroll
      for (int j = 0; j < J; ++j)
        __stcs(reinterpret_cast<float4*>(hatW + j * plane + row + n0),
               make_float4((float)w[j][0], (float)w[j][1], (float)w[j][VEC - 2], (float)w[j][VEC - 1]));
    } else {
#pragma unroll
      for (int j = 0; j < J; ++j) store_vec<T>(hatW + j * plane + row + n0, w[j]);
    }
  }

  // fixed-order block reduction in double (H8: deterministic, no float atomics)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (lanes are summed in the storage type: one 32-bit shuffle per step instead of two)
#pragma unroll
  for (int i = 0; i < NA - 1; ++i) {
    const T d = warp_sum(mom.get(i));
    if (lane == 0) s_red[warp][i] = (double)d;
  }
  {
    double d = warp_sum(acc_ll);
    if (lane == 0) s_red[warp][NA - 1] = d;
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
__global__ void estep_finalize_kernel(const double* __restrict__ partial,
                                      const double2* __restrict__ A, SubMap map, int R, int J,
                                      int F, long N, int nsplit, double2* __restrict__ hat_Rss,
                                      double2* __restrict__ hat_Rxs, double* __restrict__ ll_f) {
  const int f = blockIdx.x;
  const int NA = nacc(J), NP = npairs(J);
  __shared__ double s_acc[nacc(MAXJ)];
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
      hr += s_acc[4 * NP + 8 * J + map.src_of_sub[r1]] * invN;
      hi = 0.0;  // Hermitian symmetrisation (audioModel.py:733-740)
    }
    hat_Rss[((size_t)f * R + r1) * R + r2] = make_double2(hr, hi);
    hat_Rss[((size_t)f * R + r2) * R + r1] = make_double2(hr, -hi);
  }
  for (int idx = threadIdx.x; idx < 2 * R; idx += blockDim.x) {
    const int c = idx / R, r = idx % R;
    const int j = map.src_of_sub[r];
    const double* t = s_acc + 4 * NP + 8 * j + 4 * c;  // T_j[c][0], T_j[c][1]
    const double2 b0 = s_a[r][0], b1 = s_a[r][1];
    const double hr = t[0] * b0.x - t[1] * b0.y + t[2] * b1.x - t[3] * b1.y;
    const double hi = t[0] * b0.y + t[1] * b0.x + t[2] * b1.y + t[3] * b1.x;
    hat_Rxs[((size_t)f * 2 + c) * R + r] = make_double2(hr * invN, hi * invN);
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

// Tuning variant of the float32 kernel: PYFASST_ESTEP_VARIANT = OPT bits (1: packed moment
// accumulation, 2: hardware float->double conversion, 4: cp.async ring, 8: shared-memory-resident
// I/O (15 only), 16: fewer conversions (19 only), 32: interleaved splits (35, 99), 64: streaming cache hints (99 only)).
static int estep_variant() {
  const char* e = getenv("PYFASST_ESTEP_VARIANT");
  int v = e != nullptr ? atoi(e) : ESTEP_DEFAULT_VARIANT;
  if (v < 0 || (v > 7 && v != 15 && v != 19 && v != 35 && v != 99)) v = ESTEP_DEFAULT_VARIANT;
  return v;
}

template <typename T, typename C, int J, int OPT>
static int launch_estep_opt(const void* X, const void* V, const double* coef, const double* noise,
                            const SubMap& map, void* hatW, double* partial, int F, long N,
                            long ld, long chunk, int nsplit, cudaStream_t st) {
  dim3 grid(nsplit, F);
  size_t smem = 0;
  constexpr int MINB = (OPT & 8) != 0 ? ESTEP_MINB_SMEMIO : ESTEP_MINB;
  if ((OPT & 4) != 0 && sizeof(T) == 4) {
    smem = (size_t)ESTEP_DEPTH * (4 + J) * ESTEP_THREADS * 16;
    cudaError_t e = cudaFuncSetAttribute(estep_stereo_kernel<T, C, J, OPT, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("estep_stereo_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return PF_ERR_CUDA;
    }
  }
  estep_stereo_kernel<T, C, J, OPT, MINB><<<grid, ESTEP_THREADS, smem, st>>>(
      (const T*)X, (const T*)V, coef, noise, map, (T*)hatW, partial, F, N, ld, chunk, nsplit);
  return check_launch("estep_stereo_kernel");
}

template <typename T, typename C, int J>
static int launch_estep(const void* X, const void* V, const double* coef, const double* noise,
                        const SubMap& map, void* hatW, double* partial, int F, long N,
                        long ld, long chunk, int nsplit, cudaStream_t st) {
#define PF_ESTEP_ARGS X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st
  if (sizeof(T) == 8) return launch_estep_opt<T, C, J, 32>(PF_ESTEP_ARGS);  // interleaved splits
  switch (estep_variant()) {
    case 0: return launch_estep_opt<T, C, J, 0>(PF_ESTEP_ARGS);
    case 1: return launch_estep_opt<T, C, J, 1>(PF_ESTEP_ARGS);
    case 2: return launch_estep_opt<T, C, J, 2>(PF_ESTEP_ARGS);
    case 3: return launch_estep_opt<T, C, J, 3>(PF_ESTEP_ARGS);
    case 4: return launch_estep_opt<T, C, J, 4>(PF_ESTEP_ARGS);
    case 5: return launch_estep_opt<T, C, J, 5>(PF_ESTEP_ARGS);
    case 6: return launch_estep_opt<T, C, J, 6>(PF_ESTEP_ARGS);
    case 15: return launch_estep_opt<T, C, J, 15>(PF_ESTEP_ARGS);
    case 19: return launch_estep_opt<T, C, J, 19>(PF_ESTEP_ARGS);
    case 35: return launch_estep_opt<T, C, J, 35>(PF_ESTEP_ARGS);
    case 99: return launch_estep_opt<T, C, J, 99>(PF_ESTEP_ARGS);
    default: return launch_estep_opt<T, C, J, 7>(PF_ESTEP_ARGS);
  }
#undef PF_ESTEP_ARGS
}

template <typename T, typename C>
static int dispatch_estep(int J, const void* X, const void* V, const double* coef,
                          const double* noise, const SubMap& map, void* hatW, double* partial,
                          int F, long N, long ld, long chunk, int nsplit, cudaStream_t st) {
  switch (J) {
    case 1: return launch_estep<T, C, 1>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 2: return launch_estep<T, C, 2>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 3: return launch_estep<T, C, 3>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 4: return launch_estep<T, C, 4>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 5: return launch_estep<T, C, 5>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 6: return launch_estep<T, C, 6>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
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
  const bool ws = estep_use_ws(J, (long)N, dtype);
  const long pass = ws ? estep_ws_pass() : ESTEP_THREADS * vec;
  // aim for ~64 passes per CTA (16 for the warp-specialised kernel, whose pass is 6x longer) so
  // that the start and the end-of-CTA reduction are amortised (one CTA costs about one pass on
  // top of its passes: 64 instead of 32 passes per CTA is 0.716 -> 0.702 ms on configs[1],
  // profiles/r01/estep_interleave_experiment.txt), while keeping at least ~4 CTAs per SM in
  // flight on a 148-SM part
  long passes = (N + pass - 1) / pass;
  long per_cta = ws ? estep_ws_passes_per_cta() : 64;
  if (!ws) {  // tuning: PYFASST_ESTEP_PASSES = passes per CTA (a power of two)
    const char* e = getenv("PYFASST_ESTEP_PASSES");
    if (e != nullptr && atoi(e) >= 1 && atoi(e) <= 1024) per_cta = atoi(e);
  }
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

extern "C" int pf_estep_stereo(const void* X, const void* V, const void* A,
                               const int* src_of_sub, int R, int J, const double* noise_psd,
                               int F, int64_t N, int64_t ld, void* hatW, void* hat_Rss,
                               void* hat_Rxs, double* ll_f, void* workspace,
                               int64_t workspace_bytes, int64_t N_norm, int dtype, void* stream) {
  if (J > MAXJ || R > MAXR) {
    set_error("pf_estep_stereo: J=%d spatial components / R=%d sub-sources not supported "
              "(max %d / %d)", J, R, MAXJ, MAXR);
    return PF_ERR_UNSUPPORTED;
  }
  PF_REQUIRE(J >= 1 && R >= J, "pf_estep_stereo: J=%d R=%d", J, R);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64 || dtype == PF_F32_FASTMATH,
             "pf_estep_stereo: bad dtype %d", dtype);
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
  // the per-bin algebra always runs in float64 (see estep_stereo_kernel); PF_F32_FASTMATH
  // (float algebra) exists only to measure what that costs
  if (estep_use_ws(J, (long)N, dtype))
    rc = dispatch_estep_ws(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, chunk, nsplit,
                           st);
  else if (dtype == PF_F32)
    rc = dispatch_estep<float, double>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld,
                                       chunk, nsplit, st);
  else if (dtype == PF_F32_FASTMATH)
    rc = dispatch_estep<float, float>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld,
                                      chunk, nsplit, st);
  else
    rc = dispatch_estep<double, double>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld,
                                        chunk, nsplit, st);
  if (rc) return rc;
  // hat_Rss / hat_Rxs are means over N_norm frames: the local N, or the length of the whole
  // mixture when the frames are sharded over several GPUs (the partial means are then summed)
  estep_finalize_kernel<<<F, 64, 0, st>>>(partial, (const double2*)A, map, R, J, F,
                                          N_norm > 0 ? N_norm : N, nsplit,
                                          (double2*)hat_Rss, (double2*)hat_Rxs, ll_f);
  return check_launch("estep_finalize_kernel");
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
