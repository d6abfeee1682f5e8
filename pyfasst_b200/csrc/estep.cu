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
#include "common.cuh"

namespace pf {

constexpr int ESTEP_THREADS = 128;
constexpr int MAXJ = 6;
constexpr int MAXR = 16;

__host__ __device__ constexpr int npairs(int J) { return J * (J + 1) / 2; }
// accumulators per frequency: S (4 per pair), T (8 per source), sv (J), ll (1)
__host__ __device__ constexpr int nacc(int J) { return 4 * npairs(J) + 8 * J + J + 1; }
// coefficients per frequency: R_j (4 per source), D_jk (per pair)
__host__ __device__ constexpr int ncoef(int J) { return 4 * J + npairs(J); }

struct SubMap {
  int src_of_sub[MAXR];  // spatial component of each sub-source (rank column)
  double invrank[MAXJ];
};

// ---- per-frequency coefficients from the mixing matrix ----------------------
// A: complex128 [R][2][F] (mix_matrix of retrieve_subsrc_params, audioModel.py:562-576)
template <typename T>
__global__ void spat_coef_kernel(const double2* __restrict__ A, SubMap map, int R, int J,
                                 int F, T* __restrict__ coef) {
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
  T* c = coef + (size_t)f * ncoef(J);
  for (int j = 0; j < J; ++j)
    for (int e = 0; e < 4; ++e) c[4 * j + e] = (T)Rj[j][e];
  int p = 4 * J;
  for (int j = 0; j < J; ++j)
    for (int k = j; k < J; ++k) {
      double d;
      if (j == k)
        d = Rj[j][0] * Rj[j][1] - Rj[j][2] * Rj[j][2] - Rj[j][3] * Rj[j][3];
      else
        d = Rj[j][0] * Rj[k][1] + Rj[j][1] * Rj[k][0] -
            2.0 * (Rj[j][2] * Rj[k][2] + Rj[j][3] * Rj[k][3]);
      c[p++] = (T)fmax(d, 0.0);  // mixed discriminants of PSD matrices are >= 0
    }
}

// ---- the fused per-bin kernel -------------------------------------------------
template <typename T, int J>
__global__ void __launch_bounds__(ESTEP_THREADS)
estep_stereo_kernel(const T* __restrict__ X, const T* __restrict__ V,
                    const T* __restrict__ coef, const double* __restrict__ noise,
                    SubMap map, T* __restrict__ hatW, double* __restrict__ partial, int F,
                    long N, long ld, long chunk, int nsplit) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NP = npairs(J);
  constexpr int NA = nacc(J);
  constexpr int NC = ncoef(J);
  constexpr T kEps = (T)1e-10;          // audioModel.py:61 / signalTools eps
  constexpr T kLogPi = (T)1.1447298858494002;  // log(pi): Q4, log(det*pi)

  const int f = blockIdx.y;
  const int split = blockIdx.x;
  __shared__ T s_coef[NC];
  __shared__ double s_red[ESTEP_THREADS / 32][NA];
  if (threadIdx.x < NC) s_coef[threadIdx.x] = coef[(size_t)f * NC + threadIdx.x];
  __syncthreads();
  const T s2 = (T)noise[f];
  T invrank[J];
#pragma unroll
  for (int j = 0; j < J; ++j) invrank[j] = (T)map.invrank[j];

  T acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = (T)0;

  const long plane = (long)F * ld;
  const long row = (long)f * ld;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;

  for (long n0 = begin + (long)threadIdx.x * VEC; n0 < end; n0 += (long)ESTEP_THREADS * VEC) {
    T x0r[VEC], x0i[VEC], x1r[VEC], x1i[VEC], v[J][VEC], w[J][VEC];
    load_vec<T>(X + 0 * plane + row + n0, x0r);
    load_vec<T>(X + 1 * plane + row + n0, x0i);
    load_vec<T>(X + 2 * plane + row + n0, x1r);
    load_vec<T>(X + 3 * plane + row + n0, x1i);
#pragma unroll
    for (int j = 0; j < J; ++j) load_vec<T>(V + j * plane + row + n0, v[j]);

#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (n0 + e >= end) {
#pragma unroll
        for (int j = 0; j < J; ++j) w[j][e] = (T)0;
        continue;
      }
      // Sigma_x = sum_j v_j R_j + s2 I   (audioModel.py:613-652)
      T s00 = s2, s11 = s2, s01r = (T)0, s01i = (T)0;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const T vj = v[j][e];
        s00 += vj * s_coef[4 * j + 0];
        s11 += vj * s_coef[4 * j + 1];
        s01r += vj * s_coef[4 * j + 2];
        s01i += vj * s_coef[4 * j + 3];
      }
      // det Sigma as a sum of non-negative terms (no s00*s11-|s01|^2 cancellation)
      T pr[NP];
      T det = s2 * (s00 + (s11 - s2));
      {
        int p = 0;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
          for (int k = j; k < J; ++k) {
            pr[p] = v[j][e] * v[k][e];
            det += pr[p] * s_coef[4 * J + p];
            ++p;
          }
      }
      det = pf_max(det, kEps);  // Q5 clamp (det >= 0 here, so sign(det+eps) = +1)
      const T idet = pf_rcp(det);
      const T i00 = s11 * idet, i11 = s00 * idet;
      const T i01r = -s01r * idet, i01i = -s01i * idet;
      // y = Sigma^-1 x
      const T a0r = x0r[e], a0i = x0i[e], a1r = x1r[e], a1i = x1i[e];
      const T y0r = i00 * a0r + i01r * a1r - i01i * a1i;
      const T y0i = i00 * a0i + i01r * a1i + i01i * a1r;
      const T y1r = i01r * a0r + i01i * a0i + i11 * a1r;
      const T y1i = i01r * a0i - i01i * a0r + i11 * a1i;
      // log-likelihood integrand (audioModel.py:660-664)
      const T quad = a0r * y0r + a0i * y0i + a1r * y1r + a1i * y1i;
      acc[NA - 1] += pf_log(det) + kLogPi + quad;
      // M = y y^H - Sigma^-1
      const T m00 = y0r * y0r + y0i * y0i - i00;
      const T m11 = y1r * y1r + y1i * y1i - i11;
      const T m01r = y0r * y1r + y0i * y1i - i01r;
      const T m01i = y0i * y1r - y0r * y1i - i01i;
      // posterior source power (audioModel.py:727-729, :408-414)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const T q = s_coef[4 * j + 0] * m00 + s_coef[4 * j + 1] * m11 +
                    (T)2 * (s_coef[4 * j + 2] * m01r + s_coef[4 * j + 3] * m01i);
        const T vj = v[j][e];
        w[j][e] = pf_abs(vj + vj * vj * (q * invrank[j]));
      }
      // S_jk += v_j v_k M
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        acc[4 * p + 0] += pr[p] * m00;
        acc[4 * p + 1] += pr[p] * m11;
        acc[4 * p + 2] += pr[p] * m01r;
        acc[4 * p + 3] += pr[p] * m01i;
      }
      // U = x y^H ; T_j += v_j U ; sv_j += v_j
      const T u00r = a0r * y0r + a0i * y0i, u00i = a0i * y0r - a0r * y0i;
      const T u01r = a0r * y1r + a0i * y1i, u01i = a0i * y1r - a0r * y1i;
      const T u10r = a1r * y0r + a1i * y0i, u10i = a1i * y0r - a1r * y0i;
      const T u11r = a1r * y1r + a1i * y1i, u11i = a1i * y1r - a1r * y1i;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const T vj = v[j][e];
        T* t = acc + 4 * NP + 8 * j;
        t[0] += vj * u00r; t[1] += vj * u00i;
        t[2] += vj * u01r; t[3] += vj * u01i;
        t[4] += vj * u10r; t[5] += vj * u10i;
        t[6] += vj * u11r; t[7] += vj * u11i;
        acc[4 * NP + 8 * J + j] += vj;
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) store_vec<T>(hatW + j * plane + row + n0, w[j]);
  }

  // fixed-order block reduction in double (H8: deterministic, no float atomics)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    double d = warp_sum((double)acc[i]);
    if (lane == 0) s_red[warp][i] = d;
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

template <typename T, int J>
__global__ void __launch_bounds__(256)
wiener_stereo_kernel(const T* __restrict__ X, const T* __restrict__ V, const T* __restrict__ coef,
                     const double* __restrict__ noise, GroupMap gm, int ngroups,
                     T* __restrict__ Y, int F, long N, long ld) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NC = ncoef(J);
  constexpr T kEps = (T)1e-10;
  const int f = blockIdx.y;
  __shared__ T s_coef[NC];
  if (threadIdx.x < NC) s_coef[threadIdx.x] = coef[(size_t)f * NC + threadIdx.x];
  __syncthreads();
  const T s2 = (T)noise[f];
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
  T y0r[VEC], y0i[VEC], y1r[VEC], y1i[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    T s00 = s2, s11 = s2, s01r = (T)0, s01i = (T)0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      s00 += v[j][e] * s_coef[4 * j + 0];
      s11 += v[j][e] * s_coef[4 * j + 1];
      s01r += v[j][e] * s_coef[4 * j + 2];
      s01i += v[j][e] * s_coef[4 * j + 3];
    }
    T det = s2 * (s00 + (s11 - s2));
    int p = 0;
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int k = j; k < J; ++k) det += v[j][e] * v[k][e] * s_coef[4 * J + (p++)];
    det = pf_max(det, kEps);
    const T idet = pf_rcp(det);
    const T i00 = s11 * idet, i11 = s00 * idet, i01r = -s01r * idet, i01i = -s01i * idet;
    y0r[e] = i00 * x0r[e] + i01r * x1r[e] - i01i * x1i[e];
    y0i[e] = i00 * x0i[e] + i01r * x1i[e] + i01i * x1r[e];
    y1r[e] = i01r * x0r[e] + i01i * x0i[e] + i11 * x1r[e];
    y1i[e] = i01r * x0i[e] - i01i * x0r[e] + i11 * x1i[e];
  }
  for (int g = 0; g < ngroups; ++g) {
    T o0r[VEC], o0i[VEC], o1r[VEC], o1i[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      // Sigma_g = sum_{j in g} v_j R_j
      T g00 = (T)0, g11 = (T)0, g01r = (T)0, g01i = (T)0;
#pragma unroll
      for (int j = 0; j < J; ++j)
        if (gm.group_of_src[j] == g) {
          g00 += v[j][e] * s_coef[4 * j + 0];
          g11 += v[j][e] * s_coef[4 * j + 1];
          g01r += v[j][e] * s_coef[4 * j + 2];
          g01i += v[j][e] * s_coef[4 * j + 3];
        }
      // out = Sigma_g y
      o0r[e] = g00 * y0r[e] + g01r * y1r[e] - g01i * y1i[e];
      o0i[e] = g00 * y0i[e] + g01r * y1i[e] + g01i * y1r[e];
      o1r[e] = g01r * y0r[e] + g01i * y0i[e] + g11 * y1r[e];
      o1i[e] = g01r * y0i[e] - g01i * y0r[e] + g11 * y1i[e];
      if (n0 + e >= N) o0r[e] = o0i[e] = o1r[e] = o1i[e] = (T)0;
    }
    T* out = Y + (size_t)g * 4 * plane + row + n0;
    store_vec<T>(out + 0 * plane, o0r);
    store_vec<T>(out + 1 * plane, o0i);
    store_vec<T>(out + 2 * plane, o1r);
    store_vec<T>(out + 3 * plane, o1i);
  }
}

template <typename T, int J>
static int launch_wiener(const void* X, const void* V, const void* coef, const double* noise,
                         const GroupMap& gm, int ngroups, void* Y, int F, long N, long ld,
                         cudaStream_t st) {
  constexpr int VEC = VecOf<T>::N;
  dim3 grid(ceil_div(N, 256L * VEC), F);
  wiener_stereo_kernel<T, J><<<grid, 256, 0, st>>>((const T*)X, (const T*)V, (const T*)coef, noise,
                                                  gm, ngroups, (T*)Y, F, N, ld);
  return check_launch("wiener_stereo_kernel");
}

template <typename T>
static int dispatch_wiener(int J, const void* X, const void* V, const void* coef,
                           const double* noise, const GroupMap& gm, int ngroups, void* Y, int F,
                           long N, long ld, cudaStream_t st) {
  switch (J) {
    case 1: return launch_wiener<T, 1>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 2: return launch_wiener<T, 2>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 3: return launch_wiener<T, 3>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 4: return launch_wiener<T, 4>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 5: return launch_wiener<T, 5>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
    case 6: return launch_wiener<T, 6>(X, V, coef, noise, gm, ngroups, Y, F, N, ld, st);
  }
  set_error("pf_wiener_stereo: J=%d spatial components not supported (1..%d)", J, MAXJ);
  return PF_ERR_UNSUPPORTED;
}

template <typename T, int J>
static int launch_estep(const void* X, const void* V, const void* coef, const double* noise,
                        const SubMap& map, void* hatW, double* partial, int F, long N,
                        long ld, long chunk, int nsplit, cudaStream_t st) {
  dim3 grid(nsplit, F);
  estep_stereo_kernel<T, J><<<grid, ESTEP_THREADS, 0, st>>>(
      (const T*)X, (const T*)V, (const T*)coef, noise, map, (T*)hatW, partial, F, N, ld, chunk,
      nsplit);
  return check_launch("estep_stereo_kernel");
}

template <typename T>
static int dispatch_estep(int J, const void* X, const void* V, const void* coef,
                          const double* noise, const SubMap& map, void* hatW, double* partial,
                          int F, long N, long ld, long chunk, int nsplit, cudaStream_t st) {
  switch (J) {
    case 1: return launch_estep<T, 1>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 2: return launch_estep<T, 2>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 3: return launch_estep<T, 3>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 4: return launch_estep<T, 4>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 5: return launch_estep<T, 5>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
    case 6: return launch_estep<T, 6>(X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st);
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
  // aim for ~16 passes per CTA so the end-of-CTA reduction is amortised, while
  // keeping at least ~4 CTAs per SM in flight on a 148-SM part
  long passes = (N + pass - 1) / pass;
  long per_cta = 16;
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
                               int64_t workspace_bytes, int dtype, void* stream) {
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
  void* coef = (void*)(partial + (size_t)F * nsplit * nacc(J));
  int rc;
  if (dtype == PF_F32) {
    spat_coef_kernel<float><<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, F,
                                                             (float*)coef);
    if ((rc = check_launch("spat_coef_kernel"))) return rc;
    rc = dispatch_estep<float>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, chunk,
                               nsplit, st);
  } else {
    spat_coef_kernel<double><<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, F,
                                                              (double*)coef);
    if ((rc = check_launch("spat_coef_kernel"))) return rc;
    rc = dispatch_estep<double>(J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, chunk,
                                nsplit, st);
  }
  if (rc) return rc;
  estep_finalize_kernel<<<F, 64, 0, st>>>(partial, (const double2*)A, map, R, J, F, N, nsplit,
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
  int rc;
  if (dtype == PF_F32) {
    spat_coef_kernel<float><<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, F,
                                                             (float*)workspace);
    if ((rc = check_launch("spat_coef_kernel"))) return rc;
    return dispatch_wiener<float>(J, X, V, workspace, noise_psd, gm, ngroups, Y, F, N, ld, st);
  }
  spat_coef_kernel<double><<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, F,
                                                            (double*)workspace);
  if ((rc = check_launch("spat_coef_kernel"))) return rc;
  return dispatch_wiener<double>(J, X, V, workspace, noise_psd, gm, ngroups, Y, F, N, ld, st);
}
