// K4 -- spectral M-step of the FASST GEM loop: source power V = W H and the
// multiplicative IS-type updates of the NMF factors, sm_100a.
//
// Replaces the reference's FASST.comp_spat_comp_power (pyfasst/audioModel.py:430-498)
// and the NMF branch of FASST.update_spectral_components (audioModel.py:1509-1727).
//
// The reference forms F x N temporaries (hat_W / V^2 * other, other / V) and calls
// BLAS on them.  With K (NMF rank) <= 32 these contractions are bandwidth-bound
// when fused: each kernel below streams the F x N planes exactly once, forms the
// elementwise operand in registers and contracts it against the small factor held
// in shared memory.  Frames (n) are the contiguous axis of every plane.
//
//   spec_power_kernel : V[f,n] (+)= sum_k W[f,k] H[k,n]
//   fb_contract_kernel: num[f,k] = sum_n (hatW/P^2*O)[f,n] G[k,n],
//                       den[f,k] = sum_n (O/P)[f,n]        G[k,n]       (G = FW H)
//   tw_contract_kernel: num[k,n] = sum_f W[f,k] (O*hatW/P'^2)[f,n],
//                       den[k,n] = sum_f W[f,k] (O/P')[f,n],  P' = max(W H, eps) formed on
//                       the fly from the *updated* W (Gauss-Seidel order, Q2)
//   mult_update_kernel: theta *= (sum_s num_s / max(sum_s den_s, eps))^omega
// P = max(power of all spectral comps of the spatial comp, eps) (Q3),
// O = max(other-factor power, eps) which for single-factor models is the component's
// own power before the update (Q1/Q2).
#include "common.cuh"

namespace pf {

// ============================ V = W H ==========================================
constexpr int SP_THREADS = 256;  // 8 warps
constexpr int SP_FR = 8;         // rows of f per warp
constexpr int SP_KC = 32;        // k chunk staged in shared memory

template <typename T>
__global__ void __launch_bounds__(SP_THREADS)
spec_power_kernel(const T* __restrict__ W, int ldw, const T* __restrict__ H, long ldh,
                  T* __restrict__ V, long ldv, int F, int K, long N, int accumulate) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NT = 32 * VEC;            // frames per CTA
  constexpr int FT = (SP_THREADS / 32) * SP_FR;  // rows per CTA
  __shared__ __align__(16) T s_h[SP_KC][NT];
  __shared__ __align__(16) T s_w[FT][SP_KC + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long n_base = (long)blockIdx.x * NT;
  const int f_base = blockIdx.y * FT;
  T acc[SP_FR][VEC];
#pragma unroll
  for (int r = 0; r < SP_FR; ++r)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[r][e] = (T)0;

  for (int k0 = 0; k0 < K; k0 += SP_KC) {
    // stage H[k0:k0+KC, n_base:n_base+NT] and W[f_base:f_base+FT, k0:k0+KC]
    for (int i = threadIdx.x; i < SP_KC * 32; i += SP_THREADS) {
      const int k = i >> 5, c = i & 31;
      T tmp[VEC];
      const long n = n_base + (long)c * VEC;
      if (k0 + k < K && n < ldh) {
        load_vec<T>(H + (long)(k0 + k) * ldh + n, tmp);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) tmp[e] = (T)0;
      }
      store_vec<T>(&s_h[k][c * VEC], tmp);
    }
    for (int i = threadIdx.x; i < FT * SP_KC; i += SP_THREADS) {
      const int r = i / SP_KC, k = i % SP_KC;
      const int f = f_base + r;
      s_w[r][k] = (f < F && k0 + k < K) ? W[(long)f * ldw + k0 + k] : (T)0;
    }
    __syncthreads();
    const int kmax = (K - k0 < SP_KC) ? (K - k0) : SP_KC;
    for (int k = 0; k < kmax; ++k) {
      T h[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) h[e] = s_h[k][lane * VEC + e];
#pragma unroll
      for (int r = 0; r < SP_FR; ++r) {
        const T w = s_w[warp * SP_FR + r][k];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[r][e] += w * h[e];
      }
    }
    __syncthreads();
  }
  const long n = n_base + (long)lane * VEC;
  if (n >= ldv) return;
#pragma unroll
  for (int r = 0; r < SP_FR; ++r) {
    const int f = f_base + warp * SP_FR + r;
    if (f >= F) continue;
    T* out = V + (long)f * ldv + n;
    if (accumulate) {
      T old[VEC];
      load_vec<T>(out, old);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[r][e] += old[e];
    }
    // keep the padding frames (n >= N) at zero
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (n + e >= N) acc[r][e] = (T)0;
    store_vec<T>(out, acc[r]);
  }
}

// ============================ FB update: contract over frames ===================
constexpr int FB_THREADS = 256;

template <typename T, int KC, int FR>
__global__ void __launch_bounds__(FB_THREADS)
fb_contract_kernel(const T* __restrict__ hatW, const T* __restrict__ Pp, const T* __restrict__ Op,
                   long ld, const T* __restrict__ G, long ldg, int k0, int K, int F, long N,
                   long chunk, int nsplit, double* __restrict__ num, double* __restrict__ den) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NT = 32 * VEC;
  constexpr int FT = (FB_THREADS / 32) * FR;
  constexpr T kEps = (T)1e-10;
  __shared__ __align__(16) T s_g[KC][NT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f0 = blockIdx.y * FT + warp * FR;
  const int split = blockIdx.x;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;

  T an[FR][KC], ad[FR][KC];
#pragma unroll
  for (int r = 0; r < FR; ++r)
#pragma unroll
    for (int k = 0; k < KC; ++k) an[r][k] = ad[r][k] = (T)0;

  for (long nb = begin; nb < end; nb += NT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KC * 32; i += FB_THREADS) {
      const int k = i >> 5, c = i & 31;
      T tmp[VEC];
      const long n = nb + (long)c * VEC;
      if (k0 + k < K && n < ldg) {
        load_vec<T>(G + (long)(k0 + k) * ldg + n, tmp);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) tmp[e] = (T)0;
      }
      store_vec<T>(&s_g[k][c * VEC], tmp);
    }
    __syncthreads();
    const long n = nb + (long)lane * VEC;
    T e1[FR][VEC], e2[FR][VEC];
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      const int f = f0 + r;
      if (f < F && n < end) {
        T hw[VEC], p[VEC], o[VEC];
        load_vec<T>(hatW + (long)f * ld + n, hw);
        load_vec<T>(Pp + (long)f * ld + n, p);
        load_vec<T>(Op + (long)f * ld + n, o);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (n + e < end) {
            const T pe = pf_max(p[e], kEps), oe = pf_max(o[e], kEps);
            const T rp = pf_rcp(pe);
            // other / P (:1554-1558); exactly 1 when both are the same plane, as x/x is
            e2[r][e] = (oe == pe) ? (T)1 : oe * rp;
            e1[r][e] = hw[e] * rp * rp * oe;      // hat_W / P^2 * other (:1565-1571)
          } else {
            e1[r][e] = e2[r][e] = (T)0;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) e1[r][e] = e2[r][e] = (T)0;
      }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      T g[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) g[e] = s_g[k][lane * VEC + e];
#pragma unroll
      for (int r = 0; r < FR; ++r)
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          an[r][k] += e1[r][e] * g[e];
          ad[r][k] += e2[r][e] * g[e];
        }
    }
  }
  // lanes hold partial sums over their frames: reduce across the warp in double
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    const int f = f0 + r;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const double sn = warp_sum((double)an[r][k]);
      const double sd = warp_sum((double)ad[r][k]);
      if (lane == 0 && f < F && k0 + k < K) {
        num[((size_t)split * F + f) * K + k0 + k] = sn;
        den[((size_t)split * F + f) * K + k0 + k] = sd;
      }
    }
  }
}

// ============================ TW update: contract over frequencies ==============
constexpr int TW_THREADS = 128;
constexpr int TW_FT = 64;  // rows of W staged per step

template <typename T, int KC>
__global__ void __launch_bounds__(TW_THREADS)
tw_contract_kernel(const T* __restrict__ hatW, const T* __restrict__ Op, long ld,
                   const T* __restrict__ W, int ldw, const T* __restrict__ H, long ldh, int K,
                   int F, long N, int fchunk, int fsplit, double* __restrict__ num,
                   double* __restrict__ den, long ldo) {
  constexpr T kEps = (T)1e-10;
  __shared__ __align__(16) T s_w[TW_FT][KC];
  const long n = (long)blockIdx.x * TW_THREADS + threadIdx.x;
  const int split = blockIdx.y;
  const int fb = split * fchunk;
  int fe = fb + fchunk;
  if (fe > F) fe = F;
  const bool live = n < N;
  T h[KC], an[KC], ad[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    h[k] = (live && k < K) ? H[(long)k * ldh + n] : (T)0;
    an[k] = ad[k] = (T)0;
  }
  for (int ft = fb; ft < fe; ft += TW_FT) {
    __syncthreads();
    for (int i = threadIdx.x; i < TW_FT * KC; i += TW_THREADS) {
      const int r = i / KC, k = i % KC;
      const int f = ft + r;
      s_w[r][k] = (f < fe && k < K) ? W[(long)f * ldw + k] : (T)0;
    }
    __syncthreads();
    const int rows = (fe - ft < TW_FT) ? (fe - ft) : TW_FT;
    if (!live) continue;
    for (int r = 0; r < rows; ++r) {
      const long idx = (long)(ft + r) * ld + n;
      const T hw = __ldg(hatW + idx);
      const T o = pf_max(__ldg(Op + idx), kEps);
      T w[KC];
      T p = (T)0;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        w[k] = s_w[r][k];
        p += w[k] * h[k];
      }
      p = pf_max(p, kEps);                 // own power with the updated W (:1639-1645)
      const T rp = pf_rcp(p);
      const T e2 = o * rp;                 // other / P            (:1694-1701)
      const T e1 = o * (hw * rp * rp);     // other * hat_W / P^2  (:1714-1720)
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        an[k] += w[k] * e1;
        ad[k] += w[k] * e2;
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int k = 0; k < KC; ++k)
    if (k < K) {
      num[((size_t)split * K + k) * ldo + n] = (double)an[k];
      den[((size_t)split * K + k) * ldo + n] = (double)ad[k];
    }
}

// ============================ small elementwise helpers =========================
// out[i] = sum_s in[s][i]   (fixed order)
__global__ void sum_splits_kernel(const double* __restrict__ in, int nsplit, long count,
                                  double* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double d = 0.0;
  for (int s = 0; s < nsplit; ++s) d += in[(size_t)s * count + i];
  out[i] = d;
}

// theta[r][c] *= (num/max(den, eps))^omega over a rows x cols view (:1573, :1725)
template <typename T>
__global__ void mult_update_kernel(T* __restrict__ theta, long ldt, const double* __restrict__ num,
                                   const double* __restrict__ den, long ldnd, int rows,
                                   long cols, double omega) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= cols || r >= rows) return;
  const double ratio = num[(size_t)r * ldnd + c] / fmax(den[(size_t)r * ldnd + c], 1e-10);
  const double g = (omega == 1.0) ? ratio : pow(ratio, omega);
  theta[(size_t)r * ldt + c] = (T)((double)theta[(size_t)r * ldt + c] * g);
}

template <typename T, int KC, int FR>
static int launch_fb(const void* hatW, const void* P, const void* O, long ld, const void* G,
                     long ldg, int k0, int K, int F, long N, long chunk, int nsplit, double* num,
                     double* den, cudaStream_t st) {
  constexpr int FT = (FB_THREADS / 32) * FR;
  dim3 grid(nsplit, ceil_div(F, FT));
  fb_contract_kernel<T, KC, FR><<<grid, FB_THREADS, 0, st>>>(
      (const T*)hatW, (const T*)P, (const T*)O, ld, (const T*)G, ldg, k0, K, F, N, chunk, nsplit,
      num, den);
  return check_launch("fb_contract_kernel");
}

template <typename T>
static int dispatch_fb(const void* hatW, const void* P, const void* O, long ld, const void* G,
                       long ldg, int K, int F, long N, long chunk, int nsplit, double* num,
                       double* den, cudaStream_t st) {
  int rc = PF_OK;
  for (int k0 = 0; k0 < K && rc == PF_OK; k0 += 32) {
    const int kc = K - k0;
    if (kc <= 4)
      rc = launch_fb<T, 4, 8>(hatW, P, O, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, den, st);
    else if (kc <= 8)
      rc = launch_fb<T, 8, 4>(hatW, P, O, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, den, st);
    else if (kc <= 16)
      rc = launch_fb<T, 16, 2>(hatW, P, O, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, den, st);
    else
      rc = launch_fb<T, 32, 2>(hatW, P, O, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, den, st);
  }
  return rc;
}

template <typename T, int KC>
static int launch_tw(const void* hatW, const void* O, long ld, const void* W, int ldw,
                     const void* H, long ldh, int K, int F, long N, int fchunk, int fsplit,
                     double* num, double* den, long ldo, cudaStream_t st) {
  dim3 grid(ceil_div(N, TW_THREADS), fsplit);
  tw_contract_kernel<T, KC><<<grid, TW_THREADS, 0, st>>>((const T*)hatW, (const T*)O, ld,
                                                         (const T*)W, ldw, (const T*)H, ldh, K, F,
                                                         N, fchunk, fsplit, num, den, ldo);
  return check_launch("tw_contract_kernel");
}

template <typename T>
static int dispatch_tw(const void* hatW, const void* O, long ld, const void* W, int ldw,
                       const void* H, long ldh, int K, int F, long N, int fchunk, int fsplit,
                       double* num, double* den, long ldo, cudaStream_t st) {
  if (K <= 4) return launch_tw<T, 4>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num, den, ldo, st);
  if (K <= 8) return launch_tw<T, 8>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num, den, ldo, st);
  if (K <= 16) return launch_tw<T, 16>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num, den, ldo, st);
  if (K <= 32) return launch_tw<T, 32>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num, den, ldo, st);
  set_error("pf_nmf_tw_contract: K=%d > 32 NMF components not supported yet", K);
  return PF_ERR_UNSUPPORTED;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_spec_power(const void* W, int ldw, const void* H, int64_t ldh, void* V,
                             int64_t ldv, int F, int K, int64_t N, int accumulate, int dtype,
                             void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_spec_power: bad dtype %d", dtype);
  PF_REQUIRE(F > 0 && K > 0 && N > 0, "pf_spec_power: empty problem");
  PF_REQUIRE(ldv % 4 == 0 && ldh % 4 == 0 && ldv >= N, "pf_spec_power: ldv/ldh must be multiples of 4");
  cudaStream_t st = as_stream(stream);
  const int FT = (SP_THREADS / 32) * SP_FR;
  if (dtype == PF_F32) {
    dim3 grid(ceil_div(ldv, 32 * 4), ceil_div(F, FT));
    spec_power_kernel<float><<<grid, SP_THREADS, 0, st>>>((const float*)W, ldw, (const float*)H,
                                                         ldh, (float*)V, ldv, F, K, N, accumulate);
  } else {
    dim3 grid(ceil_div(ldv, 32 * 2), ceil_div(F, FT));
    spec_power_kernel<double><<<grid, SP_THREADS, 0, st>>>((const double*)W, ldw,
                                                          (const double*)H, ldh, (double*)V, ldv,
                                                          F, K, N, accumulate);
  }
  return check_launch("spec_power_kernel");
}

extern "C" int pf_nmf_fb_plan(int F, int K, int64_t N, int dtype, int64_t* chunk, int* nsplit) {
  const long vec = dtype == PF_F64 ? 2 : 4;
  const long nt = 32 * vec;
  const int kc = K >= 17 ? 32 : (K > 8 ? 16 : (K > 4 ? 8 : 4));
  const int fr = kc == 4 ? 8 : (kc == 8 ? 4 : 2);
  const int ft = (FB_THREADS / 32) * fr;
  const long fblocks = (F + ft - 1) / ft;
  long steps = (N + nt - 1) / nt;
  // enough CTAs for ~4 per SM, at least 8 steps each
  long ns = (148L * 4 + fblocks - 1) / fblocks;
  if (ns > steps / 8) ns = steps / 8;
  if (ns < 1) ns = 1;
  long per = (steps + ns - 1) / ns;
  *chunk = per * nt;
  *nsplit = (int)((N + *chunk - 1) / *chunk);
  return PF_OK;
}

extern "C" int pf_nmf_fb_contract(const void* hatW, const void* P, const void* O, int64_t ld,
                                  const void* G, int64_t ldg, int F, int K, int64_t N,
                                  double* num_partial, double* den_partial, int64_t chunk,
                                  int nsplit, int dtype, void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_nmf_fb_contract: bad dtype %d", dtype);
  PF_REQUIRE(ld % 4 == 0 && ldg % 4 == 0, "pf_nmf_fb_contract: ld/ldg must be multiples of 4");
  PF_REQUIRE(chunk > 0 && (int64_t)nsplit * chunk >= N, "pf_nmf_fb_contract: bad split plan");
  cudaStream_t st = as_stream(stream);
  if (dtype == PF_F32)
    return dispatch_fb<float>(hatW, P, O, ld, G, ldg, K, F, N, chunk, nsplit, num_partial,
                              den_partial, st);
  return dispatch_fb<double>(hatW, P, O, ld, G, ldg, K, F, N, chunk, nsplit, num_partial,
                             den_partial, st);
}

extern "C" int pf_nmf_tw_plan(int F, int K, int64_t N, int* fchunk, int* fsplit) {
  const long nblocks = (N + TW_THREADS - 1) / TW_THREADS;
  long fs = (148L * 8 + nblocks - 1) / nblocks;
  if (fs > (F + 63) / 64) fs = (F + 63) / 64;
  if (fs < 1) fs = 1;
  int fc = (int)((F + fs - 1) / fs);
  fc = ((fc + TW_FT - 1) / TW_FT) * TW_FT;
  *fchunk = fc;
  *fsplit = (F + fc - 1) / fc;
  return PF_OK;
}

extern "C" int pf_nmf_tw_contract(const void* hatW, const void* O, int64_t ld, const void* W,
                                  int ldw, const void* H, int64_t ldh, int F, int K, int64_t N,
                                  double* num_partial, double* den_partial, int64_t ldo,
                                  int fchunk, int fsplit, int dtype, void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_nmf_tw_contract: bad dtype %d", dtype);
  PF_REQUIRE(fchunk > 0 && (int64_t)fsplit * fchunk >= F, "pf_nmf_tw_contract: bad split plan");
  cudaStream_t st = as_stream(stream);
  if (dtype == PF_F32)
    return dispatch_tw<float>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num_partial,
                              den_partial, ldo, st);
  return dispatch_tw<double>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num_partial,
                             den_partial, ldo, st);
}

extern "C" int pf_sum_splits(const double* in, int nsplit, int64_t count, double* out,
                             void* stream) {
  PF_REQUIRE(nsplit >= 1 && count > 0, "pf_sum_splits: empty");
  sum_splits_kernel<<<ceil_div(count, 256), 256, 0, as_stream(stream)>>>(in, nsplit, count, out);
  return check_launch("sum_splits_kernel");
}

extern "C" int pf_mult_update(void* theta, int64_t ldt, const double* num, const double* den,
                              int64_t ldnd, int rows, int64_t cols, double omega, int dtype,
                              void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_mult_update: bad dtype %d", dtype);
  PF_REQUIRE(rows > 0 && cols > 0, "pf_mult_update: empty");
  dim3 grid(ceil_div(cols, 256), rows);
  if (dtype == PF_F32)
    mult_update_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((float*)theta, ldt, num, den,
                                                                  ldnd, rows, cols, omega);
  else
    mult_update_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((double*)theta, ldt, num, den,
                                                                   ldnd, rows, cols, omega);
  return check_launch("mult_update_kernel");
}
