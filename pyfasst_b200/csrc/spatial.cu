// K3 / K5 -- spatial M-step and parameter renormalisation of the FASST GEM loop.
//
// Replaces FASST.update_mix_matrix (pyfasst/audioModel.py:766-889) and
// FASST.renormalize_parameters (audioModel.py:1980-2040).  These are O(F) sized
// problems (latency, not bandwidth): everything is done in float64 on the device so
// that the GEM loop never synchronises with the host.
//
// Mixing matrix layout: complex128 A[R][I][F] (`mix_matrix` of retrieve_subsrc_params,
// audioModel.py:562-576); statistics hat_Rss[F][R][R], hat_Rxs[F][I][R] complex128.
#include "common.cuh"

namespace pf {

constexpr int SMAXR = 16;
constexpr int SMAXJ = 8;

struct IdxList {
  int n_upd, n_oth;
  int upd[SMAXR], oth[SMAXR];
};

// ---- instantaneous mixing: f-summed real statistics (audioModel.py:816-826) ------
// out[0 : I*Ku]        = sum_f Re( hat_Rxs[f][c][upd[u]] - sum_o A[oth[o]][c][f] hat_Rss[f][oth[o]][upd[u]] )
// out[I*Ku : +Ku*Ku]   = sum_f Re( hat_Rss[f][upd[u1]][upd[u2]] )
__global__ void mix_inst_stats_kernel(const double2* __restrict__ Rss,
                                      const double2* __restrict__ Rxs,
                                      const double2* __restrict__ A, IdxList L, int R, int I,
                                      int F, double* __restrict__ out) {
  const int Ku = L.n_upd;
  const int o = blockIdx.x;
  double acc = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    if (o < I * Ku) {
      const int c = o / Ku, u = o % Ku;
      double v = Rxs[((size_t)f * I + c) * R + L.upd[u]].x;
      for (int q = 0; q < L.n_oth; ++q) {
        const double2 a = A[((size_t)L.oth[q] * I + c) * F + f];
        const double2 s = Rss[((size_t)f * R + L.oth[q]) * R + L.upd[u]];
        v -= a.x * s.x - a.y * s.y;
      }
      acc += v;
    } else {
      const int p = o - I * Ku;
      acc += Rss[((size_t)f * R + L.upd[p / Ku]) * R + L.upd[p % Ku]].x;
    }
  }
  __shared__ double s_red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
    out[o] = d;
  }
}

// Solve rm_Rss^T X = rm_Rxs^T (real, LU with partial pivoting like LAPACK gesv) and
// broadcast X over all frequencies (audioModel.py:830-839).
__global__ void mix_inst_solve_kernel(const double* __restrict__ stats, double Ftot, IdxList L,
                                      int I, int F, double2* __restrict__ A, int* __restrict__ flag) {
  const int Ku = L.n_upd;
  __shared__ double M[SMAXR][SMAXR + 1];
  __shared__ double B[SMAXR][4];
  __shared__ int s_bad;
  if (threadIdx.x == 0) {
    s_bad = 0;
    // M = rm_Rss^T ; B = rm_Rxs^T   (means over f)
    for (int a = 0; a < Ku; ++a)
      for (int b = 0; b < Ku; ++b) M[a][b] = stats[I * Ku + b * Ku + a] / Ftot;
    for (int a = 0; a < Ku; ++a)
      for (int c = 0; c < I; ++c) B[a][c] = stats[c * Ku + a] / Ftot;
    for (int k = 0; k < Ku; ++k) {
      int piv = k;
      double best = fabs(M[k][k]);
      for (int r = k + 1; r < Ku; ++r)
        if (fabs(M[r][k]) > best) { best = fabs(M[r][k]); piv = r; }
      if (best == 0.0) { s_bad = 1; break; }
      if (piv != k) {
        for (int c = 0; c < Ku; ++c) { double t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
        for (int c = 0; c < I; ++c) { double t = B[k][c]; B[k][c] = B[piv][c]; B[piv][c] = t; }
      }
      for (int r = k + 1; r < Ku; ++r) {
        const double l = M[r][k] / M[k][k];
        for (int c = k + 1; c < Ku; ++c) M[r][c] -= l * M[k][c];
        for (int c = 0; c < I; ++c) B[r][c] -= l * B[k][c];
      }
    }
    if (!s_bad)
      for (int k = Ku - 1; k >= 0; --k)
        for (int c = 0; c < I; ++c) {
          double v = B[k][c];
          for (int q = k + 1; q < Ku; ++q) v -= M[k][q] * B[q][c];
          B[k][c] = v / M[k][k];
        }
    if (s_bad) atomicOr(flag, PF_FLAG_SINGULAR);
  }
  __syncthreads();
  if (s_bad) return;
  for (long i = threadIdx.x; i < (long)Ku * I * F; i += blockDim.x) {
    const int f = (int)(i % F);
    const int c = (int)((i / F) % I);
    const int u = (int)(i / ((long)F * I));
    A[((size_t)L.upd[u] * I + c) * F + f] = make_double2(B[u][c], 0.0);
  }
}

// ---- convolutive mixing: per-frequency complex solve (audioModel.py:847-857) -----
// A[:, :, f] = solve(hat_Rss[f]^T, hat_Rxs[f]^T); all sub-sources updated (Q7).
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  const double d = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}

template <int RMAX, int IMAX>
__global__ void mix_conv_solve_kernel(const double2* __restrict__ Rss,
                                      const double2* __restrict__ Rxs, int R, int I, int F,
                                      double2* __restrict__ A, int* __restrict__ flag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double2 M[RMAX][RMAX], B[RMAX][IMAX];
  for (int a = 0; a < R; ++a) {
    for (int b = 0; b < R; ++b) M[a][b] = Rss[((size_t)f * R + b) * R + a];  // transpose
    for (int c = 0; c < I; ++c) B[a][c] = Rxs[((size_t)f * I + c) * R + a];  // transpose
  }
  bool bad = false;
  for (int k = 0; k < R && !bad; ++k) {
    int piv = k;
    double best = fabs(M[k][k].x) + fabs(M[k][k].y);  // LAPACK izamax uses |re|+|im|
    for (int r = k + 1; r < R; ++r) {
      const double m = fabs(M[r][k].x) + fabs(M[r][k].y);
      if (m > best) { best = m; piv = r; }
    }
    if (best == 0.0) { bad = true; break; }
    if (piv != k) {
      for (int c = 0; c < R; ++c) { double2 t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
      for (int c = 0; c < I; ++c) { double2 t = B[k][c]; B[k][c] = B[piv][c]; B[piv][c] = t; }
    }
    for (int r = k + 1; r < R; ++r) {
      const double2 l = cdiv(M[r][k], M[k][k]);
      for (int c = k + 1; c < R; ++c) {
        const double2 t = cmul(l, M[k][c]);
        M[r][c].x -= t.x; M[r][c].y -= t.y;
      }
      for (int c = 0; c < I; ++c) {
        const double2 t = cmul(l, B[k][c]);
        B[r][c].x -= t.x; B[r][c].y -= t.y;
      }
    }
  }
  if (bad) { atomicOr(flag, PF_FLAG_SINGULAR); return; }
  for (int k = R - 1; k >= 0; --k)
    for (int c = 0; c < I; ++c) {
      double2 v = B[k][c];
      for (int q = k + 1; q < R; ++q) {
        const double2 t = cmul(M[k][q], B[q][c]);
        v.x -= t.x; v.y -= t.y;
      }
      B[k][c] = cdiv(v, M[k][k]);
    }
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < I; ++c) A[((size_t)r * I + c) * F + f] = B[r][c];
}

// The same elimination with ONE WARP per frequency (round 2): the augmented matrix [Rss^T | Rxs^T]
// lives in shared memory, the lanes share the pivot search (first maximum, as izamax), the
// multipliers and the rank-one update of a step.  Every element sees the operations of the
// one-thread kernel above in the same order, so the results are bit-identical; a 16 x 16 system
// per frequency takes ~20 us for 1025 frequencies instead of 317 us (one thread per frequency with
// its matrix in local memory was pure latency).
template <int RMAX, int IMAX, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
mix_conv_solve_warp_kernel(const double2* __restrict__ Rss, const double2* __restrict__ Rxs, int R,
                           int I, int F, double2* __restrict__ A, int* __restrict__ flag) {
  __shared__ double2 s_m[WARPS][RMAX][RMAX + IMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * WARPS + warp;
  if (f >= F) return;  // (the whole warp)
  double2 (*M)[RMAX + IMAX] = s_m[warp];
  const int C = R + I;
  for (int idx = lane; idx < R * C; idx += 32) {
    const int a = idx / C, c = idx % C;
    M[a][c] = c < R ? Rss[((size_t)f * R + c) * R + a] : Rxs[((size_t)f * I + (c - R)) * R + a];
  }
  __syncwarp();
  for (int k = 0; k < R; ++k) {
    double best = -1.0;
    int piv = k;
    for (int r = k + lane; r < R; r += 32) {
      const double m = fabs(M[r][k].x) + fabs(M[r][k].y);  // LAPACK izamax uses |re|+|im|
      if (m > best) { best = m; piv = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // first maximum
      const double mo = __shfl_xor_sync(0xffffffffu, best, o);
      const int po = __shfl_xor_sync(0xffffffffu, piv, o);
      if (mo > best || (mo == best && po < piv)) { best = mo; piv = po; }
    }
    if (best == 0.0) {
      if (lane == 0) atomicOr(flag, PF_FLAG_SINGULAR);
      return;
    }
    if (piv != k)
      for (int c = lane; c < C; c += 32) {
        const double2 t = M[k][c];
        M[k][c] = M[piv][c];
        M[piv][c] = t;
      }
    __syncwarp();
    const double2 pk = M[k][k];
    for (int r = k + 1 + lane; r < R; r += 32) M[r][k] = cdiv(M[r][k], pk);  // the multipliers
    __syncwarp();
    const int nr = R - k - 1, nc = C - k - 1;
    for (int idx = lane; idx < nr * nc; idx += 32) {
      const int r = k + 1 + idx / nc, c = k + 1 + idx % nc;
      const double2 t = cmul(M[r][k], M[k][c]);
      M[r][c].x -= t.x;
      M[r][c].y -= t.y;
    }
    __syncwarp();
  }
  for (int k = R - 1; k >= 0; --k) {
    if (lane < I) {
      double2 v = M[k][R + lane];
      for (int q = k + 1; q < R; ++q) {
        const double2 t = cmul(M[k][q], M[q][R + lane]);
        v.x -= t.x;
        v.y -= t.y;
      }
      M[k][R + lane] = cdiv(v, M[k][k]);
    }
    __syncwarp();
  }
  for (int idx = lane; idx < R * I; idx += 32) {
    const int r = idx / I, c = idx % I;
    A[((size_t)r * I + c) * F + f] = M[r][R + c];
  }
}

// ---- renormalisation (audioModel.py:1991-1996) ------------------------------------
struct SrcMap {
  int src_of_sub[SMAXR];
};

// sums[j] = sum_{r in j, c, f} |A|^2 ; counts are known on the host
__global__ void spat_energy_kernel(const double2* __restrict__ A, SrcMap map, int R, int I, int F,
                                   double* __restrict__ sums) {
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int r = 0; r < R; ++r) {
    if (map.src_of_sub[r] != j) continue;
    for (long i = threadIdx.x; i < (long)I * F; i += blockDim.x) {
      const double2 a = A[(size_t)r * I * F + i];
      acc += a.x * a.x + a.y * a.y;
    }
  }
  __shared__ double s_red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
    sums[j] = d;
  }
}

// A_r /= sqrt(energy[src(r)])  with energy = sums / counts
__global__ void spat_scale_kernel(double2* __restrict__ A, SrcMap map, int R, int I, int F,
                                  const double* __restrict__ sums, const double* __restrict__ counts) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)R * I * F) return;
  const int r = (int)(i / ((long)I * F));
  const int j = map.src_of_sub[r];
  const double s = 1.0 / sqrt(sums[j] / counts[j]);
  A[i].x *= s;
  A[i].y *= s;
}

// FB *= g ; colmax[k] = max_f FB[f][k]     (audioModel.py:2009-2010)
template <typename T>
__global__ void fb_scale_colmax_kernel(T* __restrict__ FB, int ldw, int F, int K,
                                       const double* __restrict__ sums,
                                       const double* __restrict__ counts, int j,
                                       double* __restrict__ colmax) {
  const int k = blockIdx.x;
  const double g = sums[j] / counts[j];
  double m = -1.0e300;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const double v = (double)FB[(size_t)f * ldw + k] * g;
    FB[(size_t)f * ldw + k] = (T)v;
    // the maximum is taken on the stored (rounded) value, like FB.max(axis=0)
    const double sv = (double)FB[(size_t)f * ldw + k];
    m = sv > m ? sv : m;
  }
  __shared__ double s_red[32];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) d = s_red[w] > d ? s_red[w] : d;
    colmax[k] = d;
  }
}

// FW *= vstack(w) ; w2 = FW.mean(axis=0) ; FW /= w2   (audioModel.py:2011-2018), one CTA
template <typename T>
__global__ void fw_renorm_kernel(T* __restrict__ FW, int ldfw, int Kb, int Kw,
                                 const double* __restrict__ colmax, double* __restrict__ w_out,
                                 double* __restrict__ w2_out) {
  __shared__ double s_w[64], s_w2[64];
  for (int k = threadIdx.x; k < Kb; k += blockDim.x) {
    double w = colmax[k];
    if (w == 0.0) w = 1.0;
    s_w[k] = w;
    w_out[k] = w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Kb * Kw; i += blockDim.x) {
    const int r = i / Kw, c = i % Kw;
    FW[(size_t)r * ldfw + c] = (T)((double)FW[(size_t)r * ldfw + c] * s_w[r]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Kw; c += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < Kb; ++r) s += (double)FW[(size_t)r * ldfw + c];
    double w2 = s / Kb;
    if (w2 == 0.0) w2 = 1.0;
    s_w2[c] = w2;
    w2_out[c] = w2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Kb * Kw; i += blockDim.x) {
    const int r = i / Kw, c = i % Kw;
    FW[(size_t)r * ldfw + c] = (T)((double)FW[(size_t)r * ldfw + c] / s_w2[c]);
  }
}

// The same for large FW (the K x K weights of a large dictionary): rows scaled by one CTA each,
// column means by one CTA each; the division by w2 is a scale_matrix launch.
template <typename T>
__global__ void fw_scale_rows_kernel(T* __restrict__ FW, int ldfw, int Kw,
                                     const double* __restrict__ colmax,
                                     double* __restrict__ w_out) {
  const int r = blockIdx.x;
  double w = colmax[r];
  if (w == 0.0) w = 1.0;
  if (threadIdx.x == 0) w_out[r] = w;
  for (int c = threadIdx.x; c < Kw; c += blockDim.x)
    FW[(size_t)r * ldfw + c] = (T)((double)FW[(size_t)r * ldfw + c] * w);
}
template <typename T>
__global__ void fw_col_means_kernel(const T* __restrict__ FW, int ldfw, int Kb,
                                    double* __restrict__ w2_out) {
  const int c = blockIdx.x;
  double s = 0.0;
  for (int r = threadIdx.x; r < Kb; r += blockDim.x) s += (double)FW[(size_t)r * ldfw + c];
  __shared__ double s_red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
    d /= Kb;
    w2_out[c] = d == 0.0 ? 1.0 : d;
  }
}

// M[r][c] *= (by_row ? s[r] : s[c]) or /= ; optional sum of the result (TW restart test,
// audioModel.py:2023): one atomic per CTA after a block reduction
constexpr int SCALE_EPT = 8;
template <typename T>
__global__ void scale_matrix_kernel(T* __restrict__ M, long ldm, int rows, long cols,
                                    const double* __restrict__ s, int by_row, int divide,
                                    double* __restrict__ total) {
  const int r = blockIdx.y;
  double v = 0.0;
  if (r < rows) {  // SCALE_EPT elements per thread, a block-wide stride apart (coalesced)
#pragma unroll
    for (int i = 0; i < SCALE_EPT; ++i) {
      const long c = ((long)blockIdx.x * SCALE_EPT + i) * blockDim.x + threadIdx.x;
      if (c < cols) {
        const double sc = by_row ? s[r] : s[c];
        double x = (double)M[(size_t)r * ldm + c];
        x = divide ? x / sc : x * sc;
        M[(size_t)r * ldm + c] = (T)x;
        v += x;
      }
    }
  }
  if (total != nullptr) {
    __shared__ double s_red[32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double d = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
      if (d != 0.0) atomicAdd(total, d);
    }
  }
}

}  // namespace pf

using namespace pf;

static int fill_list(IdxList* L, const int* upd, int n_upd, const int* oth, int n_oth, int R) {
  PF_REQUIRE(n_upd >= 0 && n_upd <= SMAXR && n_oth >= 0 && n_oth <= SMAXR && R <= SMAXR,
             "spatial update: too many sub-sources (max %d)", SMAXR);
  L->n_upd = n_upd;
  L->n_oth = n_oth;
  for (int i = 0; i < n_upd; ++i) L->upd[i] = upd[i];
  for (int i = 0; i < n_oth; ++i) L->oth[i] = oth[i];
  return PF_OK;
}

extern "C" int pf_mix_inst_stats(const void* hat_Rss, const void* hat_Rxs, const void* A,
                                 const int* upd, int n_upd, const int* oth, int n_oth, int R,
                                 int I, int F, double* stats, void* stream) {
  IdxList L;
  int rc = fill_list(&L, upd, n_upd, oth, n_oth, R);
  if (rc) return rc;
  PF_REQUIRE(n_upd > 0 && I >= 1 && I <= 4, "pf_mix_inst_stats: n_upd=%d I=%d", n_upd, I);
  mix_inst_stats_kernel<<<I * n_upd + n_upd * n_upd, 256, 0, as_stream(stream)>>>(
      (const double2*)hat_Rss, (const double2*)hat_Rxs, (const double2*)A, L, R, I, F, stats);
  return check_launch("mix_inst_stats_kernel");
}

extern "C" int pf_mix_inst_solve(const double* stats, double F_total, const int* upd, int n_upd,
                                 int I, int F, void* A, int* flags, void* stream) {
  IdxList L;
  int rc = fill_list(&L, upd, n_upd, nullptr, 0, n_upd);
  if (rc) return rc;
  PF_REQUIRE(n_upd > 0 && I >= 1 && I <= 4, "pf_mix_inst_solve: n_upd=%d I=%d", n_upd, I);
  mix_inst_solve_kernel<<<1, 256, 0, as_stream(stream)>>>(stats, F_total, L, I, F, (double2*)A,
                                                        flags);
  return check_launch("mix_inst_solve_kernel");
}

extern "C" int pf_mix_conv_solve(const void* hat_Rss, const void* hat_Rxs, int R, int I, int F,
                                 void* A, int* flags, void* stream) {
  PF_REQUIRE(R >= 1 && R <= SMAXR && I >= 1 && I <= 4, "pf_mix_conv_solve: R=%d I=%d", R, I);
  cudaStream_t st = as_stream(stream);
  // PYFASST_MIX_SOLVE_THREAD=1: the one-thread-per-frequency kernel (A/B)
  static const bool per_thread = [] {
    const char* e = getenv("PYFASST_MIX_SOLVE_THREAD");
    return e != nullptr && atoi(e) != 0;
  }();
  if (!per_thread) {
    mix_conv_solve_warp_kernel<SMAXR, 4, 4><<<ceil_div(F, 4), 128, 0, st>>>(
        (const double2*)hat_Rss, (const double2*)hat_Rxs, R, I, F, (double2*)A, flags);
    return check_launch("mix_conv_solve_warp_kernel");
  }
  if (R <= 8)
    mix_conv_solve_kernel<8, 4><<<ceil_div(F, 64), 64, 0, st>>>(
        (const double2*)hat_Rss, (const double2*)hat_Rxs, R, I, F, (double2*)A, flags);
  else
    mix_conv_solve_kernel<16, 4><<<ceil_div(F, 64), 64, 0, st>>>(
        (const double2*)hat_Rss, (const double2*)hat_Rxs, R, I, F, (double2*)A, flags);
  return check_launch("mix_conv_solve_kernel");
}

extern "C" int pf_spat_energy(const void* A, const int* src_of_sub, int R, int J, int I, int F,
                              double* sums, void* stream) {
  PF_REQUIRE(R >= 1 && R <= SMAXR && J >= 1 && J <= SMAXJ, "pf_spat_energy: R=%d J=%d", R, J);
  SrcMap map;
  for (int r = 0; r < R; ++r) map.src_of_sub[r] = src_of_sub[r];
  spat_energy_kernel<<<J, 256, 0, as_stream(stream)>>>((const double2*)A, map, R, I, F, sums);
  return check_launch("spat_energy_kernel");
}

extern "C" int pf_spat_scale(void* A, const int* src_of_sub, int R, int I, int F,
                             const double* sums, const double* counts, void* stream) {
  PF_REQUIRE(R >= 1 && R <= SMAXR, "pf_spat_scale: R=%d", R);
  SrcMap map;
  for (int r = 0; r < R; ++r) map.src_of_sub[r] = src_of_sub[r];
  spat_scale_kernel<<<ceil_div((long)R * I * F, 256), 256, 0, as_stream(stream)>>>(
      (double2*)A, map, R, I, F, sums, counts);
  return check_launch("spat_scale_kernel");
}

extern "C" int pf_fb_scale_colmax(void* FB, int ldw, int F, int K, const double* sums,
                                  const double* counts, int j, double* colmax, int dtype,
                                  void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_fb_scale_colmax: bad dtype %d", dtype);
  PF_REQUIRE(K >= 1 && K <= 65535, "pf_fb_scale_colmax: K=%d", K);
  if (dtype == PF_F32)
    fb_scale_colmax_kernel<float><<<K, 256, 0, as_stream(stream)>>>((float*)FB, ldw, F, K, sums,
                                                                   counts, j, colmax);
  else
    fb_scale_colmax_kernel<double><<<K, 256, 0, as_stream(stream)>>>((double*)FB, ldw, F, K, sums,
                                                                    counts, j, colmax);
  return check_launch("fb_scale_colmax_kernel");
}

extern "C" int pf_fw_renorm(void* FW, int ldfw, int Kb, int Kw, const double* colmax, double* w,
                            double* w2, int dtype, void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_fw_renorm: bad dtype %d", dtype);
  PF_REQUIRE(Kb >= 1 && Kb <= 65535 && Kw >= 1 && Kw <= 65535, "pf_fw_renorm: Kb=%d Kw=%d", Kb, Kw);
  if (Kb > 64 || Kw > 64) {  // large weights: three launches instead of one CTA
    cudaStream_t st = as_stream(stream);
    if (dtype == PF_F32) {
      fw_scale_rows_kernel<float><<<Kb, 256, 0, st>>>((float*)FW, ldfw, Kw, colmax, w);
      fw_col_means_kernel<float><<<Kw, 256, 0, st>>>((const float*)FW, ldfw, Kb, w2);
    } else {
      fw_scale_rows_kernel<double><<<Kb, 256, 0, st>>>((double*)FW, ldfw, Kw, colmax, w);
      fw_col_means_kernel<double><<<Kw, 256, 0, st>>>((const double*)FW, ldfw, Kb, w2);
    }
    int rc = check_launch("fw_col_means_kernel");
    if (rc) return rc;
    return pf_scale_matrix(FW, ldfw, Kb, Kw, w2, 0, 1, nullptr, dtype, stream);
  }
  if (dtype == PF_F32)
    fw_renorm_kernel<float><<<1, 256, 0, as_stream(stream)>>>((float*)FW, ldfw, Kb, Kw, colmax, w, w2);
  else
    fw_renorm_kernel<double><<<1, 256, 0, as_stream(stream)>>>((double*)FW, ldfw, Kb, Kw, colmax, w, w2);
  return check_launch("fw_renorm_kernel");
}

extern "C" int pf_scale_matrix(void* M, int64_t ldm, int rows, int64_t cols, const double* s,
                               int by_row, int divide, double* total, int dtype, void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_scale_matrix: bad dtype %d", dtype);
  PF_REQUIRE(rows > 0 && rows <= 65535 && cols > 0, "pf_scale_matrix: rows=%d", rows);
  dim3 grid(ceil_div(cols, 256L * SCALE_EPT), rows);
  if (dtype == PF_F32)
    scale_matrix_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((float*)M, ldm, rows, cols, s,
                                                                   by_row, divide, total);
  else
    scale_matrix_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((double*)M, ldm, rows, cols,
                                                                    s, by_row, divide, total);
  return check_launch("scale_matrix_kernel");
}
