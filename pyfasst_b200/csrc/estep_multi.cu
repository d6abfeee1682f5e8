// K2 / K6 for I = 2..4 channels -- the fused per-bin E-step and the Wiener filter of the FASST
// GEM loop with general I x I Hermitian algebra, sm_100a.
//
// The reference implements the E-step and the separation for stereo only (audioModel.py:394,
// :605, :1127 raise for other channel counts); BASELINE.json configs[3] asks for a 4-channel
// mixture.  These kernels extend the definitions of FASST.compute_suff_stat
// (audioModel.py:580-764) and compute_Wiener_gain_2d (:1396-1467) to I channels exactly as
// oracle/fasst_oracle.py: estep_general does (batched inverse instead of the 2 x 2 closed form,
// the determinant clamp of signalTools.py:186-188 applied to the generic determinant), which is
// checked against the stereo path at I = 2.
//
// Per bin: Sigma = s2 I + sum_j v_j R_j, Cholesky Sigma = L L^H in registers (float64),
// L^-1, Sigma^-1 = L^-H L^-1, y = Sigma^-1 x, M = y y^H - Sigma^-1,
// hatW_j = | v_j + v_j^2 tr(M R_j) / rank_j |, and the per-frequency moments
//     S_jk = sum_n v_j v_k M (I^2 reals),  T_j = sum_n v_j x y^H (2 I^2 reals),  sum_n v_j,  ll.
// At I = 4, J = 4 that is 293 accumulators per frequency -- too many for registers.  Each thread
// therefore writes its bin's factors (v_j v_k, M, v_j, x y^H) to a shared-memory record, and the
// warp accumulates the outer products cooperatively: lane l owns the accumulators l, l+32, ...
// and walks the 32 records of the warp (2 shared loads + 1 FMA per accumulator and bin), in
// float64.
#include "common.cuh"

namespace pf {
void estep_timing_begin(cudaStream_t st);  // estep.cu
void estep_timing_end(cudaStream_t st);

constexpr int EM_THREADS = 128;
#ifndef EM_MINB
#define EM_MINB 2
#endif
constexpr int EM_MAXJ = 6;
constexpr int EM_MAXR = 24;
constexpr int EM_MAXI = 4;

struct MultiMap {
  int src_of_sub[EM_MAXR];
  double invrank[EM_MAXJ];
};
struct MultiGroups {
  int group_of_src[EM_MAXJ];
};

// lower-triangle index of (i, k), i > k
__host__ __device__ constexpr int tri(int i, int k) { return i * (i - 1) / 2 + k; }
__host__ __device__ constexpr int em_npairs(int J) { return J * (J + 1) / 2; }
// per-frequency accumulators: S (I^2 per pair), Z (I^2 per source), sv (J), clamp correction
// (2 I^2 per source), ll (1)
__host__ __device__ constexpr int em_nacc(int I, int J) {
  return em_npairs(J) * I * I + J * I * I + J + J * 2 * I * I + 1;
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// conj(a) * b
__device__ __forceinline__ double2 cconjmul(double2 a, double2 b) {
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}

// fused accumulations (4 DFMA each; a separate product + add costs 6 FP64 operations)
// acc += a * b
__device__ __forceinline__ void cmac(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a * conj(b)
__device__ __forceinline__ void cmsubc(double2& acc, double2 a, double2 b) {
  acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(-a.y, b.x, acc.y); acc.y = fma(a.x, b.y, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cconjmac(double2& acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}

// ---- per-frequency coefficients: R_j = sum_{r in j} a_r a_r^H as (diag[I], lower triangle) ----
// A: complex128 [R][I][F];  coef[f][j][I*I] = { Re R_ii (I), (Re, Im) R_ik for i > k }
__global__ void spat_coef_multi_kernel(const double2* __restrict__ A, MultiMap map, int R, int J,
                                       int I, int F, double* __restrict__ coef) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int NM = I * I;
  double* c = coef + (size_t)f * J * NM;
  for (int i = 0; i < J * NM; ++i) c[i] = 0.0;
  for (int r = 0; r < R; ++r) {
    double* cj = c + map.src_of_sub[r] * NM;
    for (int i = 0; i < I; ++i) {
      const double2 ai = A[((size_t)r * I + i) * F + f];
      cj[i] += ai.x * ai.x + ai.y * ai.y;
      for (int k = 0; k < i; ++k) {
        const double2 ak = A[((size_t)r * I + k) * F + f];
        cj[I + 2 * tri(i, k) + 0] += ai.x * ak.x + ai.y * ak.y;  // Re a_i conj(a_k)
        cj[I + 2 * tri(i, k) + 1] += ai.y * ak.x - ai.x * ak.y;  // Im a_i conj(a_k)
      }
    }
  }
}

// ---- shared per-bin algebra ---------------------------------------------------------------------
// In: v[J] (double), x[I] (complex), coefficients of the frequency.  Out: y = Sigma^-1 x and
// Sigma^-1 (diag + lower triangle), both already scaled by det / max(det, eps) (Q5 clamp on the
// generic determinant), the clamped determinant and x^H Sigma^-1 x.
template <int I, int J>
__device__ __forceinline__ void sigma_inverse_multi(const double (&v)[J], const double2 (&x)[I],
                                                    const double* __restrict__ coef, double s2,
                                                    double2 (&y)[I], double (&sid)[I],
                                                    double2 (&sio)[I * (I - 1) / 2 + 1],
                                                    double& detc, double& quad) {
  constexpr int NM = I * I, NT = I * (I - 1) / 2;
  double sd[I];
  double2 so[NT + 1];
#pragma unroll
  for (int i = 0; i < I; ++i) sd[i] = s2;
#pragma unroll
  for (int t = 0; t < NT; ++t) so[t] = make_double2(0.0, 0.0);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const double* c = coef + j * NM;
#pragma unroll
    for (int i = 0; i < I; ++i) sd[i] += v[j] * c[i];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      so[t].x += v[j] * c[I + 2 * t];
      so[t].y += v[j] * c[I + 2 * t + 1];
    }
  }
  // Cholesky Sigma = L L^H, in place: so -> strictly lower part of L, linv = 1 / L_jj
  double linv[I];
  double det = 1.0;
#pragma unroll
  for (int j = 0; j < I; ++j) {
    double s = sd[j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= so[tri(j, k)].x * so[tri(j, k)].x + so[tri(j, k)].y * so[tri(j, k)].y;
    s = fmax(s, 1e-300);  // Sigma is positive definite (s2 > 0); guards rounding only
    det *= s;
    linv[j] = rsqrt(s);
#pragma unroll
    for (int i = j + 1; i < I; ++i) {
      double2 a = so[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) {
        const double2 p = cmulc(so[tri(i, k)], so[tri(j, k)]);
        a.x -= p.x;
        a.y -= p.y;
      }
      so[tri(i, j)] = make_double2(a.x * linv[j], a.y * linv[j]);
    }
  }
  // Li = L^-1 (lower triangular; real diagonal linv), in lo
  double2 lo[NT + 1];
#pragma unroll
  for (int j = 0; j < I; ++j) {
#pragma unroll
    for (int i = j + 1; i < I; ++i) {
      // acc = sum_{k=j}^{i-1} L[i][k] Li[k][j]
      double2 a = make_double2(so[tri(i, j)].x * linv[j], so[tri(i, j)].y * linv[j]);
#pragma unroll
      for (int k = j + 1; k < i; ++k) {
        const double2 p = cmul(so[tri(i, k)], lo[tri(k, j)]);
        a.x += p.x;
        a.y += p.y;
      }
      lo[tri(i, j)] = make_double2(-a.x * linv[i], -a.y * linv[i]);
    }
  }
  detc = fmax(det, 1e-10);  // sign(det + eps) max(|det|, eps) with det > 0
  const double sc = det / detc;
  // Sigma^-1[i][j] = sum_{k >= i} conj(Li[k][i]) Li[k][j]   (i >= j)
#pragma unroll
  for (int i = 0; i < I; ++i) {
    double d = linv[i] * linv[i];
#pragma unroll
    for (int k = i + 1; k < I; ++k) d += lo[tri(k, i)].x * lo[tri(k, i)].x + lo[tri(k, i)].y * lo[tri(k, i)].y;
    sid[i] = d * sc;
#pragma unroll
    for (int j = 0; j < i; ++j) {
      // k = i term: conj(Li[i][i]) Li[i][j] = linv[i] * lo(i, j)
      double2 a = make_double2(linv[i] * lo[tri(i, j)].x, linv[i] * lo[tri(i, j)].y);
#pragma unroll
      for (int k = i + 1; k < I; ++k) {
        const double2 p = cconjmul(lo[tri(k, i)], lo[tri(k, j)]);
        a.x += p.x;
        a.y += p.y;
      }
      sio[tri(i, j)] = make_double2(a.x * sc, a.y * sc);
    }
  }
  // z = Li x ; quad = |z|^2 ; y = Li^H z
  double2 z[I];
  quad = 0.0;
#pragma unroll
  for (int i = 0; i < I; ++i) {
    double2 a = make_double2(linv[i] * x[i].x, linv[i] * x[i].y);
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double2 p = cmul(lo[tri(i, k)], x[k]);
      a.x += p.x;
      a.y += p.y;
    }
    z[i] = a;
    quad += a.x * a.x + a.y * a.y;
  }
  quad *= sc;
#pragma unroll
  for (int i = 0; i < I; ++i) {
    double2 a = make_double2(linv[i] * z[i].x, linv[i] * z[i].y);
#pragma unroll
    for (int k = i + 1; k < I; ++k) {
      const double2 p = cconjmul(lo[tri(k, i)], z[k]);
      a.x += p.x;
      a.y += p.y;
    }
    y[i] = make_double2(a.x * sc, a.y * sc);
  }
}

// ---- LDL^H variant of the per-bin algebra (E-step) ---------------------------------------------
// Sigma = L D L^H (L unit lower), Li = L^-1, Sigma^-1 = Li^H D^-1 Li, y = Li^H D^-1 Li x.  No
// square roots, 4 reciprocals; ~230 FP64 operations at I = 4 without the formation of Sigma.
// Out: y and Sigma^-1 (diag + lower triangle), both scaled by sc = det / max(det, eps) (Q5 clamp on
// the generic determinant), the clamped determinant, x^H Sigma^-1 x and sc.
template <int I, int J>
__device__ __forceinline__ void sigma_inverse_ldl(const double (&v)[J], const double2 (&x)[I],
                                                  const double* __restrict__ coef, double s2,
                                                  double2 (&y)[I], double (&sid)[I],
                                                  double2 (&sio)[I * (I - 1) / 2 + 1],
                                                  double& detc, double& quad, double& sc) {
  constexpr int NM = I * I, NT = I * (I - 1) / 2;
  double sd[I];
  double2 w[NT + 1];  // lower triangle of Sigma, then the unnormalised columns of L
#pragma unroll
  for (int i = 0; i < I; ++i) sd[i] = s2;
#pragma unroll
  for (int t = 0; t < NT; ++t) w[t] = make_double2(0.0, 0.0);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const double* c = coef + j * NM;
#pragma unroll
    for (int i = 0; i < I; ++i) sd[i] += v[j] * c[i];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      w[t].x += v[j] * c[I + 2 * t];
      w[t].y += v[j] * c[I + 2 * t + 1];
    }
  }
  double r[I];
  double2 L[NT + 1];
  double det = 1.0;
#pragma unroll
  for (int j = 0; j < I; ++j) {
    double d = sd[j];
#pragma unroll
    for (int k = 0; k < j; ++k) {
      d = fma(-w[tri(j, k)].x, L[tri(j, k)].x, d);
      d = fma(-w[tri(j, k)].y, L[tri(j, k)].y, d);
    }
    d = fmax(d, 1e-30);  // Sigma is positive definite (s2 > 0); guards rounding only
    det *= d;
    r[j] = fast_rcp(d);
#pragma unroll
    for (int i = j + 1; i < I; ++i) {
      double2 a = w[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) cmsubc(a, w[tri(i, k)], L[tri(j, k)]);
      w[tri(i, j)] = a;
      L[tri(i, j)] = make_double2(a.x * r[j], a.y * r[j]);
    }
  }
  // nLi = -L^-1 (strictly lower part), G = D^-1 L^-1 (strictly lower part)
  double2 Li[NT + 1], G[NT + 1];
#pragma unroll
  for (int j = 0; j < I; ++j) {
#pragma unroll
    for (int i = j + 1; i < I; ++i) {
      double2 a = make_double2(-L[tri(i, j)].x, -L[tri(i, j)].y);
#pragma unroll
      for (int k = j + 1; k < i; ++k) {  // a -= L[i][k] Li[k][j]
        a.x = fma(-L[tri(i, k)].x, Li[tri(k, j)].x, a.x); a.x = fma(L[tri(i, k)].y, Li[tri(k, j)].y, a.x);
        a.y = fma(-L[tri(i, k)].x, Li[tri(k, j)].y, a.y); a.y = fma(-L[tri(i, k)].y, Li[tri(k, j)].x, a.y);
      }
      Li[tri(i, j)] = a;
      G[tri(i, j)] = make_double2(a.x * r[i], a.y * r[i]);
    }
  }
  // Sigma^-1[i][j] = sum_{k >= i} conj(Li[k][i]) G[k][j], Li[k][k] = 1
#pragma unroll
  for (int i = 0; i < I; ++i) {
    double d = r[i];
#pragma unroll
    for (int k = i + 1; k < I; ++k) {
      d = fma(Li[tri(k, i)].x, G[tri(k, i)].x, d);
      d = fma(Li[tri(k, i)].y, G[tri(k, i)].y, d);
    }
    sid[i] = d;
#pragma unroll
    for (int j = 0; j < i; ++j) {
      double2 a = G[tri(i, j)];
#pragma unroll
      for (int k = i + 1; k < I; ++k) cconjmac(a, Li[tri(k, i)], G[tri(k, j)]);
      sio[tri(i, j)] = a;
    }
  }
  // z = Li x, t = D^-1 z, quad = z^H t, y = Li^H t
  double2 tt[I];
  quad = 0.0;
#pragma unroll
  for (int i = 0; i < I; ++i) {
    double2 a = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) cmac(a, Li[tri(i, k)], x[k]);
    tt[i] = make_double2(a.x * r[i], a.y * r[i]);
    quad = fma(a.x, tt[i].x, quad);
    quad = fma(a.y, tt[i].y, quad);
  }
#pragma unroll
  for (int i = 0; i < I; ++i) {
    double2 a = tt[i];
#pragma unroll
    for (int k = i + 1; k < I; ++k) cconjmac(a, Li[tri(k, i)], tt[k]);
    y[i] = a;
  }
  detc = det;
  sc = 1.0;
  if (det < 1e-10) {  // sign(det + eps) max(|det|, eps) with det > 0: Sigma^-1 := adj(Sigma) / eps
    detc = 1e-10;
    sc = det * 1e10;
    quad *= sc;
#pragma unroll
    for (int i = 0; i < I; ++i) {
      sid[i] *= sc;
      y[i].x *= sc;
      y[i].y *= sc;
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      sio[t].x *= sc;
      sio[t].y *= sc;
    }
  }
}

// D(8x8) += A(8x4) B(4x8) in float64 on the tensor-core path (DMMA).  Lane l holds A[l>>2][l&3],
// B[l&3][l>>2] and D[l>>2][2 (l&3) + {0, 1}].  DMMA shares the FP64 pipe with DFMA
// (profiles/r02/micro_dmma.txt: same 16 FMA per clock and scheduler), so it buys no arithmetic
// throughput -- it replaces 8 DFMA + their operand loads per lane by one instruction and two loads.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- the moments as a small GEMM over the bins --------------------------------------------------
// Per frequency:  out[w][c] = sum_n W[w][n] Mc[c][n]  with the WR = NP + J weights
// W = (v_j v_k for j <= k, then v_j) and the I^2 components of M (diag, then lower (re, im)):
// rows 0..NP-1 are S_jk, rows NP..NP+J-1 are Z_j = sum_n v_j M.  The cross moments follow from
// x y^H = Sigma M + I:   T_j = sv_j I + s2 Z_j + sum_l R_l S_lj   (as in the stereo kernel), plus,
// where the determinant clamp is active (Sigma_c^-1 = sc Sigma^-1 is not the inverse of Sigma),
// the correction  sum_n v_j (1 - sc) (x y_c^H - I),  accumulated by a second small GEMM only in
// the passes of a warp that contain a clamped bin.
// Every thread owns one bin per pass: it writes its weights and M components to a
// component-major tile of shared memory ([component][bin], component stride 36 doubles: the
// stores of a warp and the fragment loads below are both bank-conflict free), then the warp
// contracts the 32 bins with DMMA m8n8k4 (8 steps of 4 bins).
template <int I, int J>
struct Mom {
  static constexpr int NP = em_npairs(J);
  static constexpr int NM = I * I, NU = 2 * I * I;
  static constexpr int WR = NP + J;            // weight rows
  static constexpr int MT = (WR + 7) / 8;      // 8-row tiles of weights
  static constexpr int NT = (NM + 7) / 8;      // 8-column tiles of M components
  static constexpr int NT2 = (NU + 7) / 8;     // 8-column tiles of the clamp correction
  static constexpr int CS = 36;                // doubles per component (32 bins + 4: banks)
  static constexpr int NCOMP = (MT + NT) * 8 > (1 + NT2) * 8 ? (MT + NT) * 8 : (1 + NT2) * 8;
  static constexpr int NA = em_nacc(I, J);
};

template <typename T, int I, int J>
__global__ void __launch_bounds__(EM_THREADS, EM_MINB)
estep_multi_kernel(const T* __restrict__ X, const T* __restrict__ V,
                   const double* __restrict__ coef, const double* __restrict__ noise, MultiMap map,
                   T* __restrict__ hatW, double* __restrict__ partial, int F, long N, long ld,
                   long chunk, int nsplit) {
  typedef Mom<I, J> MM;
  constexpr int NM = I * I, NT = I * (I - 1) / 2, NA = MM::NA, CS = MM::CS;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char em_smem[];
  double* s_coef = reinterpret_cast<double*>(em_smem);   // [J][NM]
  double* s_coef2 = s_coef + J * NM;                     // the same, off-diagonal entries doubled
  double* s_rec = s_coef2 + J * NM;                      // [warps][NCOMP][CS]; [warps][NA] at the end
  const int f = blockIdx.y, split = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  for (int i = threadIdx.x; i < J * NM; i += EM_THREADS) {
    const double c = coef[(size_t)f * J * NM + i];
    s_coef[i] = c;
    s_coef2[i] = (i % NM) >= I ? 2.0 * c : c;
  }
  double* wrec = s_rec + (size_t)warp * MM::NCOMP * CS;
  for (int i = lane; i < MM::NCOMP * CS; i += 32) wrec[i] = 0.0;  // padding rows stay zero
  __syncthreads();
  const double s2 = noise[f];
  T invrank[J];
#pragma unroll
  for (int j = 0; j < J; ++j) invrank[j] = (T)map.invrank[j];

  double acc[MM::MT][MM::NT][2], cacc[MM::NT2][2], sv[J];
#pragma unroll
  for (int a = 0; a < MM::MT; ++a)
#pragma unroll
    for (int b = 0; b < MM::NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll
  for (int b = 0; b < MM::NT2; ++b) cacc[b][0] = cacc[b][1] = 0.0;
#pragma unroll
  for (int j = 0; j < J; ++j) sv[j] = 0.0;
  double acc_ll = 0.0;

  const long plane = (long)F * ld, row = (long)f * ld;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;
  // the inputs of the next pass are loaded before the algebra of the current one (the kernel is
  // latency bound otherwise: 3 warps per scheduler, ~1500 cycles of dependent FP64 per pass)
  T vt_n[J], xt_n[2 * I];
  auto fetch = [&](long n) {
    const bool ok = n < end;
#pragma unroll
    for (int j = 0; j < J; ++j) vt_n[j] = ok ? V[j * plane + row + n] : (T)0;
#pragma unroll
    for (int i = 0; i < 2 * I; ++i) xt_n[i] = ok ? X[i * plane + row + n] : (T)0;
  };
  fetch(begin + warp * 32 + lane);
  for (long n0 = begin + warp * 32; n0 < end; n0 += EM_THREADS) {  // warp-uniform trip count
    const long n = n0 + lane;
    const bool live = n < end;
    T vt[J];
    double v[J];
    double2 x[I];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      vt[j] = vt_n[j];
      v[j] = (double)vt[j];
    }
#pragma unroll
    for (int i = 0; i < I; ++i) x[i] = make_double2((double)xt_n[2 * i], (double)xt_n[2 * i + 1]);
    fetch(n + EM_THREADS);
    double2 y[I], sio[NT + 1];
    double sid[I], detc, quad, sc;
    sigma_inverse_ldl<I, J>(v, x, s_coef, s2, y, sid, sio, detc, quad, sc);
    // Q4: log(det * pi); float32 planes take the float logarithm as the stereo kernel does
    if (live)
      acc_ll += (sizeof(T) == 8 ? log(detc) : (double)__logf((float)detc)) + 1.1447298858494002 + quad;
    // M = y y^H - Sigma^-1 (diag + lower triangle), straight into the warp's tile
    double md[I];
    double2 mo[NT + 1];
#pragma unroll
    for (int i = 0; i < I; ++i) {
      md[i] = fma(y[i].x, y[i].x, fma(y[i].y, y[i].y, -sid[i]));
#pragma unroll
      for (int k = 0; k < i; ++k) {  // y_i conj(y_k) - Sigma^-1_ik
        mo[tri(i, k)].x = fma(y[i].x, y[k].x, fma(y[i].y, y[k].y, -sio[tri(i, k)].x));
        mo[tri(i, k)].y = fma(y[i].y, y[k].x, fma(-y[i].x, y[k].y, -sio[tri(i, k)].y));
      }
    }
    {
      double* mr = wrec + MM::MT * 8 * CS + lane;
#pragma unroll
      for (int i = 0; i < I; ++i) mr[i * CS] = md[i];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        mr[(I + 2 * t) * CS] = mo[t].x;
        mr[(I + 2 * t + 1) * CS] = mo[t].y;
      }
      double* wr = wrec + lane;
      int p = 0;
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int k = j; k < J; ++k) wr[(p++) * CS] = v[j] * v[k];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        wr[(MM::NP + j) * CS] = v[j];
        sv[j] += v[j];
      }
    }
    // posterior source powers: tr(M R_j) = sum_i M_ii R_ii + 2 sum_{i>k} Re(M_ik conj(R_ik))
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double* c = s_coef2 + j * NM;
      double q = 0.0, q2 = 0.0;  // two chains
#pragma unroll
      for (int i = 0; i < I; ++i) q = fma(md[i], c[i], q);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        q2 = fma(mo[t].x, c[I + 2 * t], q2);
        q2 = fma(mo[t].y, c[I + 2 * t + 1], q2);
      }
      const T qf = (T)(q + q2);
      if (live) hatW[j * plane + row + n] = pf_abs(vt[j] + vt[j] * vt[j] * (qf * invrank[j]));
    }
    __syncwarp();
    // contraction of the warp's 32 bins
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      double a[MM::MT], b[MM::NT];
#pragma unroll
      for (int mt = 0; mt < MM::MT; ++mt) a[mt] = wrec[(8 * mt + g) * CS + 4 * s + t4];
#pragma unroll
      for (int nt = 0; nt < MM::NT; ++nt) b[nt] = wrec[((MM::MT + nt) * 8 + g) * CS + 4 * s + t4];
#pragma unroll
      for (int mt = 0; mt < MM::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < MM::NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    const bool clamped = live && sc != 1.0;
    if (__any_sync(FULL, clamped)) {
      // slow path: rows 0..J-1 = v_j (1 - sc) (zero for the bins without clamp), columns
      // 8.. = U - I with U = x y_c^H ([a][b] (re, im)); same tile, repaired afterwards
      __syncwarp();
      const double k1 = clamped ? 1.0 - sc : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) wrec[j * CS + lane] = j < J ? k1 * v[j < J ? j : 0] : 0.0;
#pragma unroll
      for (int a = 0; a < I; ++a)
#pragma unroll
        for (int b = 0; b < I; ++b) {
          const double2 u = cmulc(x[a], y[b]);
          wrec[(8 + 2 * (a * I + b)) * CS + lane] = a == b ? u.x - 1.0 : u.x;
          wrec[(8 + 2 * (a * I + b) + 1) * CS + lane] = u.y;
        }
#pragma unroll
      for (int c = MM::NU; c < MM::NT2 * 8; ++c) wrec[(8 + c) * CS + lane] = 0.0;
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const double a = wrec[g * CS + 4 * s + t4];
#pragma unroll
        for (int nt = 0; nt < MM::NT2; ++nt)
          dmma884(cacc[nt][0], cacc[nt][1], a, wrec[((1 + nt) * 8 + g) * CS + 4 * s + t4]);
      }
      __syncwarp();
      // the padding rows / columns of the main tile that the slow path overwrote
#pragma unroll
      for (int c = MM::WR; c < MM::MT * 8; ++c) wrec[c * CS + lane] = 0.0;
#pragma unroll
      for (int c = MM::NM; c < MM::NT * 8; ++c) wrec[((MM::MT * 8) + c) * CS + lane] = 0.0;
#pragma unroll
      for (int c = (MM::MT + MM::NT) * 8; c < MM::NCOMP; ++c) wrec[c * CS + lane] = 0.0;
    }
    __syncwarp();
  }
  // block reduction in a fixed order (deterministic): the fragments go to the warp's own tile
  // (re-used as [NA] sums), then the warps are added in order
  double* red = wrec;
  __syncwarp();
  for (int i = lane; i < NA; i += 32) red[i] = 0.0;
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < MM::MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < MM::NT; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = 8 * mt + g, c = 8 * nt + 2 * t4 + i;
        if (r < MM::WR && c < NM) red[r * NM + c] = acc[mt][nt][i];
      }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const double d = warp_sum(sv[j]);
    if (lane == 0) red[MM::WR * NM + j] = d;
  }
#pragma unroll
  for (int nt = 0; nt < MM::NT2; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = 8 * nt + 2 * t4 + i;
      if (g < J && c < MM::NU) red[MM::WR * NM + J + g * MM::NU + c] = cacc[nt][i];
    }
  {
    const double d = warp_sum(acc_ll);
    if (lane == 0) red[NA - 1] = d;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NA; i += EM_THREADS) {
    double d = 0.0;
#pragma unroll
    for (int w = 0; w < EM_THREADS / 32; ++w) d += s_rec[(size_t)w * MM::NCOMP * CS + i];
    partial[((size_t)f * nsplit + split) * NA + i] = d;
  }
}

// ---- per-frequency contraction with the mixing vectors ------------------------------------------------
// hat_Rss[r1, r2] = a_r1^H S_{j1 j2} a_r2 / N + delta sv_{j1} / N ;  hat_Rxs[:, r] = T_j a_r / N with
// T_j = sv_j I + s2 Z_j + sum_l R_l S_lj + (clamp correction)
// partial: [F][nsplit][ S (I^2 per pair) | Z (I^2 per source) | sv (J) | corr (2 I^2 per source) | ll ]
__device__ __forceinline__ double2 herm_at(const double* P, int I, int a, int c) {
  if (a == c) return make_double2(P[a], 0.0);
  if (a > c) return make_double2(P[I + 2 * tri(a, c)], P[I + 2 * tri(a, c) + 1]);
  return make_double2(P[I + 2 * tri(c, a)], -P[I + 2 * tri(c, a) + 1]);
}

__global__ void estep_finalize_multi_kernel(const double* __restrict__ partial,
                                            const double2* __restrict__ A,
                                            const double* __restrict__ coef,
                                            const double* __restrict__ noise, MultiMap map, int R,
                                            int J, int I, int F, long N, int nsplit,
                                            double2* __restrict__ hat_Rss,
                                            double2* __restrict__ hat_Rxs,
                                            double* __restrict__ ll_f) {
  const int f = blockIdx.x;
  const int NM = I * I, NU = 2 * I * I, NP = em_npairs(J), NA = em_nacc(I, J);
  extern __shared__ __align__(16) unsigned char fm_smem[];
  double* s_acc = reinterpret_cast<double*>(fm_smem);                    // [NA]
  double2* s_a = reinterpret_cast<double2*>(s_acc + NA + (NA & 1));      // [R][I]
  double2* s_T = s_a + R * I;                                            // [J][I][I]
  for (int i = threadIdx.x; i < NA; i += blockDim.x) {
    double d = 0.0;
    for (int s = 0; s < nsplit; ++s) d += partial[((size_t)f * nsplit + s) * NA + i];
    s_acc[i] = d;
  }
  for (int i = threadIdx.x; i < R * I; i += blockDim.x) s_a[i] = A[(size_t)i * F + f];
  __syncthreads();
  const double invN = 1.0 / (double)N;
  if (threadIdx.x == 0) ll_f[f] = s_acc[NA - 1];
  const double* Z0 = s_acc + NP * NM;
  const double* sv = Z0 + J * NM;
  const double* C0 = sv + J;
  const double s2 = noise[f];
  for (int idx = threadIdx.x; idx < J * NM; idx += blockDim.x) {
    const int j = idx / NM, a = (idx % NM) / I, b = idx % I;
    const double2 z = herm_at(Z0 + j * NM, I, a, b);
    double tr_ = s2 * z.x + C0[j * NU + 2 * (a * I + b)] + (a == b ? sv[j] : 0.0);
    double ti = s2 * z.y + C0[j * NU + 2 * (a * I + b) + 1];
    for (int l = 0; l < J; ++l) {
      const int j1 = l < j ? l : j, j2 = l < j ? j : l;
      const double* S = s_acc + (j1 * J - j1 * (j1 - 1) / 2 + (j2 - j1)) * NM;
      const double* Rl = coef + ((size_t)f * J + l) * NM;
      for (int c = 0; c < I; ++c) {
        const double2 r = herm_at(Rl, I, a, c), s = herm_at(S, I, c, b);
        tr_ += r.x * s.x - r.y * s.y;
        ti += r.x * s.y + r.y * s.x;
      }
    }
    s_T[idx] = make_double2(tr_, ti);
  }
  for (int idx = threadIdx.x; idx < R * R; idx += blockDim.x) {
    const int r1 = idx / R, r2 = idx % R;
    if (r1 > r2) continue;
    int j1 = map.src_of_sub[r1], j2 = map.src_of_sub[r2];
    if (j1 > j2) { const int t = j1; j1 = j2; j2 = t; }
    const double* S = s_acc + (j1 * J - j1 * (j1 - 1) / 2 + (j2 - j1)) * NM;
    const double2* a = s_a + r1 * I;
    const double2* b = s_a + r2 * I;
    double hr = 0.0, hi = 0.0;
    for (int i = 0; i < I; ++i) {
      // t_i = (S b)_i with S Hermitian: S_ik = conj(S_ki) for i < k
      double tr_ = S[i] * b[i].x, ti = S[i] * b[i].y;
      for (int k = 0; k < I; ++k) {
        if (k == i) continue;
        double sr, si;
        if (i > k) { sr = S[I + 2 * tri(i, k)]; si = S[I + 2 * tri(i, k) + 1]; }
        else { sr = S[I + 2 * tri(k, i)]; si = -S[I + 2 * tri(k, i) + 1]; }
        tr_ += sr * b[k].x - si * b[k].y;
        ti += sr * b[k].y + si * b[k].x;
      }
      hr += a[i].x * tr_ + a[i].y * ti;  // conj(a_i) t_i
      hi += a[i].x * ti - a[i].y * tr_;
    }
    hr *= invN;
    hi *= invN;
    if (r1 == r2) {
      hr += sv[map.src_of_sub[r1]] * invN;
      hi = 0.0;  // Hermitian symmetrisation (audioModel.py:733-740)
    }
    hat_Rss[((size_t)f * R + r1) * R + r2] = make_double2(hr, hi);
    hat_Rss[((size_t)f * R + r2) * R + r1] = make_double2(hr, -hi);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < I * R; idx += blockDim.x) {
    const int c = idx / R, r = idx % R;
    const double2* T = s_T + (map.src_of_sub[r] * I + c) * I;  // T_j[c][b]
    const double2* b = s_a + r * I;
    double hr = 0.0, hi = 0.0;
    for (int k = 0; k < I; ++k) {
      hr += T[k].x * b[k].x - T[k].y * b[k].y;
      hi += T[k].x * b[k].y + T[k].y * b[k].x;
    }
    hat_Rxs[((size_t)f * I + c) * R + r] = make_double2(hr * invN, hi * invN);
  }
}

// ---- Wiener filter: Y_g = (sum_{j in g} v_j R_j) Sigma^-1 x, planes Y[g][2 I][F][ld] ----------------
template <typename T, int I, int J>
__global__ void __launch_bounds__(EM_THREADS)
wiener_multi_kernel(const T* __restrict__ X, const T* __restrict__ V,
                    const double* __restrict__ coef, const double* __restrict__ noise,
                    MultiGroups gm, int ngroups, T* __restrict__ Y, int F, long N, long ld) {
  constexpr int NM = I * I, NT = I * (I - 1) / 2;
  __shared__ double s_coef[EM_MAXJ * EM_MAXI * EM_MAXI];
  const int f = blockIdx.y;
  for (int i = threadIdx.x; i < J * NM; i += EM_THREADS) s_coef[i] = coef[(size_t)f * J * NM + i];
  __syncthreads();
  const long n = (long)blockIdx.x * EM_THREADS + threadIdx.x;
  if (n >= N) return;
  const long plane = (long)F * ld, row = (long)f * ld;
  double v[J];
  double2 x[I];
#pragma unroll
  for (int j = 0; j < J; ++j) v[j] = (double)V[j * plane + row + n];
#pragma unroll
  for (int i = 0; i < I; ++i)
    x[i] = make_double2((double)X[(2 * i) * plane + row + n], (double)X[(2 * i + 1) * plane + row + n]);
  double2 y[I], sio[NT + 1];
  double sid[I], detc, quad;
  sigma_inverse_multi<I, J>(v, x, s_coef, noise[f], y, sid, sio, detc, quad);
  for (int g = 0; g < ngroups; ++g) {
    double gd[I];
    double2 go[NT + 1];
#pragma unroll
    for (int i = 0; i < I; ++i) gd[i] = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) go[t] = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < J; ++j)
      if (gm.group_of_src[j] == g) {
        const double* c = s_coef + j * NM;
#pragma unroll
        for (int i = 0; i < I; ++i) gd[i] += v[j] * c[i];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          go[t].x += v[j] * c[I + 2 * t];
          go[t].y += v[j] * c[I + 2 * t + 1];
        }
      }
#pragma unroll
    for (int i = 0; i < I; ++i) {
      double2 o = make_double2(gd[i] * y[i].x, gd[i] * y[i].y);
#pragma unroll
      for (int k = 0; k < I; ++k) {
        if (k == i) continue;
        const double2 p = i > k ? cmul(go[tri(i, k)], y[k]) : cconjmul(go[tri(k, i)], y[k]);
        o.x += p.x;
        o.y += p.y;
      }
      T* out = Y + ((size_t)g * 2 * I + 2 * i) * plane + row + n;
      out[0] = (T)o.x;
      out[plane] = (T)o.y;
    }
  }
}

template <typename T, int I, int J>
static size_t em_smem_bytes() {
  static_assert(em_nacc(I, J) <= Mom<I, J>::NCOMP * Mom<I, J>::CS, "the sums re-use the tile");
  return sizeof(double) * (2 * J * I * I +
                           (size_t)(EM_THREADS / 32) * Mom<I, J>::NCOMP * Mom<I, J>::CS);
}

template <typename T, int I, int J>
static int launch_estep_multi(const void* X, const void* V, const double* coef, const double* noise,
                              const MultiMap& map, void* hatW, double* partial, int F, long N,
                              long ld, long chunk, int nsplit, cudaStream_t st) {
  const size_t smem = em_smem_bytes<T, I, J>();
  cudaError_t e = cudaFuncSetAttribute(estep_multi_kernel<T, I, J>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("estep_multi_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  dim3 grid(nsplit, F);
  estep_multi_kernel<T, I, J><<<grid, EM_THREADS, smem, st>>>(
      (const T*)X, (const T*)V, coef, noise, map, (T*)hatW, partial, F, N, ld, chunk, nsplit);
  return check_launch("estep_multi_kernel");
}

template <typename T, int I, int J>
static int launch_wiener_multi(const void* X, const void* V, const double* coef,
                               const double* noise, const MultiGroups& gm, int ngroups, void* Y,
                               int F, long N, long ld, cudaStream_t st) {
  dim3 grid(ceil_div(N, EM_THREADS), F);
  wiener_multi_kernel<T, I, J><<<grid, EM_THREADS, 0, st>>>((const T*)X, (const T*)V, coef, noise,
                                                           gm, ngroups, (T*)Y, F, N, ld);
  return check_launch("wiener_multi_kernel");
}

// dispatch over (dtype, I, J): I in 2..4, J in 1..6
#define EM_DISPATCH_J(FN, T, I_, ...)                      \
  switch (J) {                                             \
    case 1: return FN<T, I_, 1>(__VA_ARGS__);              \
    case 2: return FN<T, I_, 2>(__VA_ARGS__);              \
    case 3: return FN<T, I_, 3>(__VA_ARGS__);              \
    case 4: return FN<T, I_, 4>(__VA_ARGS__);              \
    case 5: return FN<T, I_, 5>(__VA_ARGS__);              \
    case 6: return FN<T, I_, 6>(__VA_ARGS__);              \
  }
#define EM_DISPATCH_I(FN, T, ...)                          \
  switch (I) {                                             \
    case 2: EM_DISPATCH_J(FN, T, 2, __VA_ARGS__) break;    \
    case 3: EM_DISPATCH_J(FN, T, 3, __VA_ARGS__) break;    \
    case 4: EM_DISPATCH_J(FN, T, 4, __VA_ARGS__) break;    \
  }

static int dispatch_estep_multi(int dtype, int I, int J, const void* X, const void* V,
                                const double* coef, const double* noise, const MultiMap& map,
                                void* hatW, double* partial, int F, long N, long ld, long chunk,
                                int nsplit, cudaStream_t st) {
  if (dtype == PF_F32) {
    EM_DISPATCH_I(launch_estep_multi, float, X, V, coef, noise, map, hatW, partial, F, N, ld, chunk,
                  nsplit, st)
  } else {
    EM_DISPATCH_I(launch_estep_multi, double, X, V, coef, noise, map, hatW, partial, F, N, ld,
                  chunk, nsplit, st)
  }
  set_error("pf_estep_multi: I=%d channels / J=%d spatial components not supported", I, J);
  return PF_ERR_UNSUPPORTED;
}

static int dispatch_wiener_multi(int dtype, int I, int J, const void* X, const void* V,
                                 const double* coef, const double* noise, const MultiGroups& gm,
                                 int ngroups, void* Y, int F, long N, long ld, cudaStream_t st) {
  if (dtype == PF_F32) {
    EM_DISPATCH_I(launch_wiener_multi, float, X, V, coef, noise, gm, ngroups, Y, F, N, ld, st)
  } else {
    EM_DISPATCH_I(launch_wiener_multi, double, X, V, coef, noise, gm, ngroups, Y, F, N, ld, st)
  }
  set_error("pf_wiener_multi: I=%d channels / J=%d spatial components not supported", I, J);
  return PF_ERR_UNSUPPORTED;
}

static int fill_map(const char* who, const int* src_of_sub, int R, int J, MultiMap& map) {
  int count[EM_MAXJ] = {0};
  for (int r = 0; r < R; ++r) {
    if (src_of_sub[r] < 0 || src_of_sub[r] >= J) {
      set_error("%s: src_of_sub[%d]=%d", who, r, src_of_sub[r]);
      return PF_ERR_ARG;
    }
    map.src_of_sub[r] = src_of_sub[r];
    count[src_of_sub[r]]++;
  }
  for (int j = 0; j < J; ++j) {
    if (count[j] == 0) {
      set_error("%s: spatial component %d has rank 0", who, j);
      return PF_ERR_ARG;
    }
    map.invrank[j] = 1.0 / count[j];
  }
  return PF_OK;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_estep_multi_plan(int I, int J, int F, int64_t N, int64_t* chunk, int* nsplit,
                                   int64_t* workspace_bytes) {
  PF_REQUIRE(I >= 2 && I <= EM_MAXI && J >= 1 && J <= EM_MAXJ,
             "pf_estep_multi_plan: I=%d J=%d out of range", I, J);
  // ~64 passes of 128 frames per CTA amortise the end-of-CTA reduction; keep >= 4 CTAs per SM
  const long pass = EM_THREADS;
  const long passes = (N + pass - 1) / pass;
  long per_cta = 64;
  while (per_cta > 1 && (long)F * ((passes + per_cta - 1) / per_cta) < 148L * 8) per_cta /= 2;
  const long c = per_cta * pass;
  int ns = (int)((N + c - 1) / c);
  if (ns < 1) ns = 1;
  *chunk = c;
  *nsplit = ns;
  *workspace_bytes = ((int64_t)F * ns * em_nacc(I, J) + (int64_t)F * J * I * I) * sizeof(double);
  return PF_OK;
}

extern "C" int pf_estep_multi(const void* X, const void* V, const void* A, const int* src_of_sub,
                              int R, int J, int I, const double* noise_psd, int F, int64_t N,
                              int64_t ld, void* hatW, void* hat_Rss, void* hat_Rxs, double* ll_f,
                              void* workspace, int64_t workspace_bytes, int64_t N_norm, int dtype,
                              void* stream) {
  if (I < 2 || I > EM_MAXI || J < 1 || J > EM_MAXJ || R < J || R > EM_MAXR) {
    set_error("pf_estep_multi: I=%d channels, J=%d spatial components, R=%d sub-sources not "
              "supported (I 2..%d, J 1..%d, R <= %d)", I, J, R, EM_MAXI, EM_MAXJ, EM_MAXR);
    return PF_ERR_UNSUPPORTED;
  }
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_estep_multi: bad dtype %d", dtype);
  PF_REQUIRE(ld >= N && F > 0 && N > 0, "pf_estep_multi: F=%d N=%ld ld=%ld", F, (long)N, (long)ld);
  MultiMap map;
  int rc = fill_map("pf_estep_multi", src_of_sub, R, J, map);
  if (rc) return rc;
  int64_t chunk, need;
  int nsplit;
  pf_estep_multi_plan(I, J, F, N, &chunk, &nsplit, &need);
  PF_REQUIRE(workspace_bytes >= need, "pf_estep_multi: workspace %ld < %ld bytes",
             (long)workspace_bytes, (long)need);
  cudaStream_t st = as_stream(stream);
  double* partial = (double*)workspace;
  double* coef = partial + (size_t)F * nsplit * em_nacc(I, J);
  spat_coef_multi_kernel<<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, I, F, coef);
  rc = check_launch("spat_coef_multi_kernel");
  if (rc) return rc;
  estep_timing_begin(st);
  rc = dispatch_estep_multi(dtype, I, J, X, V, coef, noise_psd, map, hatW, partial, F, N, ld, chunk,
                            nsplit, st);
  estep_timing_end(st);
  if (rc) return rc;
  const int NA = em_nacc(I, J);
  const size_t smem = sizeof(double) * (NA + (NA & 1)) + sizeof(double2) * (R * I + J * I * I);
  estep_finalize_multi_kernel<<<F, 128, smem, st>>>(partial, (const double2*)A, coef, noise_psd, map,
                                                   R, J, I, F, N_norm > 0 ? N_norm : N, nsplit,
                                                   (double2*)hat_Rss, (double2*)hat_Rxs, ll_f);
  return check_launch("estep_finalize_multi_kernel");
}

extern "C" int pf_wiener_multi(const void* X, const void* V, const void* A, const int* src_of_sub,
                               int R, int J, int I, const double* noise_psd,
                               const int* group_of_src, int ngroups, int F, int64_t N, int64_t ld,
                               void* Y, void* workspace, int64_t workspace_bytes, int dtype,
                               void* stream) {
  if (I < 2 || I > EM_MAXI || J < 1 || J > EM_MAXJ || R < J || R > EM_MAXR) {
    set_error("pf_wiener_multi: I=%d J=%d R=%d not supported", I, J, R);
    return PF_ERR_UNSUPPORTED;
  }
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_wiener_multi: bad dtype %d", dtype);
  PF_REQUIRE(ld >= N && F > 0 && N > 0 && ngroups >= 1 && ngroups <= J,
             "pf_wiener_multi: F=%d N=%ld ngroups=%d", F, (long)N, ngroups);
  PF_REQUIRE(workspace_bytes >= (int64_t)F * J * I * I * 8, "pf_wiener_multi: workspace too small");
  MultiMap map;
  int rc = fill_map("pf_wiener_multi", src_of_sub, R, J, map);
  if (rc) return rc;
  MultiGroups gm;
  for (int j = 0; j < EM_MAXJ; ++j) gm.group_of_src[j] = -1;
  for (int j = 0; j < J; ++j) {
    PF_REQUIRE(group_of_src[j] >= -1 && group_of_src[j] < ngroups,
               "pf_wiener_multi: group_of_src[%d]=%d", j, group_of_src[j]);
    gm.group_of_src[j] = group_of_src[j];
  }
  cudaStream_t st = as_stream(stream);
  double* coef = (double*)workspace;
  spat_coef_multi_kernel<<<ceil_div(F, 128), 128, 0, st>>>((const double2*)A, map, R, J, I, F, coef);
  rc = check_launch("spat_coef_multi_kernel");
  if (rc) return rc;
  return dispatch_wiener_multi(dtype, I, J, X, V, coef, noise_psd, gm, ngroups, Y, F, N, ld, st);
}
