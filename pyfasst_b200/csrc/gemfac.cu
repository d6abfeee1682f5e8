// K4 for general factor structures (source/filter models: several factors per spectral component,
// free FW, large dictionaries), sm_100a.
//
// The reference's update_spectral_components (pyfasst/audioModel.py:1469-1727) is, for every free
// matrix of every factor, two dense contractions of the planes
//     num_plane = hat_W / P^2 * O ,   den_plane = O / P        (P, O clamped at eps, :1513-1520)
// with the other matrices of the factor.  The fast path (nmf.cu / nmf_tc.cu) fuses the planes
// into the contractions for single-factor NMF with K <= 32; here the planes are formed once
// (one bandwidth-bound pass) and contracted by the tensor-core GEMM of gemm_tc.cu, which is what
// the K = 1093 glottal dictionary of multiChanSourceF0Filter (audioModel.py:2551-2760) needs.
#include "common.cuh"

namespace pf {

constexpr int GF_THREADS = 256;
constexpr double GF_EPS = 1e-10;  // audioModel.py:72

// out[f][0:ld] = hatW / max(P,eps)^2 * max(O,eps), out[f][ld:2ld] = max(O,eps) / max(P,eps);
// zero in the padding columns n >= N (the planes are contracted over n with padded lengths).
// With lambdaCorr > 0 (Ptot != NULL): the correlation penalty of audioModel.py:1484-1703,
//     c = lambda * Pminus / max(Ptot^2, eps),  den = O (1/P + c),  num = (hatW / P^2 + 2 c P / Ptot) O
template <typename T>
__global__ void __launch_bounds__(GF_THREADS)
gem_ratio_planes_kernel(const T* __restrict__ hatW, const T* __restrict__ P,
                        const T* __restrict__ O, T* __restrict__ out, int F, long N, long ld,
                        const T* __restrict__ Ptot, const T* __restrict__ Pminus, double lambda) {
  const long n = (long)blockIdx.x * GF_THREADS + threadIdx.x;
  const int f = blockIdx.y;
  if (n >= ld) return;
  T num = (T)0, den = (T)0;
  if (n < N) {
    const size_t i = (size_t)f * ld + n;
    const T p = pf_max(P[i], (T)GF_EPS);
    const T o = pf_max(O[i], (T)GF_EPS);
    const T ip = (T)1 / p;
    if (Ptot == nullptr) {
      den = o * ip;
      num = hatW[i] * ip * ip * o;
    } else {
      const T pt = Ptot[i];
      const T c = (T)lambda * Pminus[i] / pf_max(pt * pt, (T)GF_EPS);
      den = o * (ip + c);
      num = (hatW[i] * ip * ip + c * ((T)2 * (p / pt))) * o;
    }
  }
  out[(size_t)f * 2 * ld + n] = num;
  out[(size_t)f * 2 * ld + ld + n] = den;
}

// Ptot = max(sum_j V_j, eps), Pminus = Ptot - max(V_own, eps) (clamped at eps when `clamp`):
// the powers the correlation penalty is built from (audioModel.py:1484-1508)
template <typename T>
__global__ void __launch_bounds__(GF_THREADS)
corr_planes_kernel(const T* __restrict__ V, int J, int own, T* __restrict__ Ptot,
                   T* __restrict__ Pminus, int F, long N, long ld, int clamp) {
  const long n = (long)blockIdx.x * GF_THREADS + threadIdx.x;
  const int f = blockIdx.y;
  if (n >= ld) return;
  const size_t i = (size_t)f * ld + n;
  T pt = (T)0, pm = (T)0;
  if (n < N) {
    for (int j = 0; j < J; ++j) pt += V[(size_t)j * F * ld + i];
    pt = pf_max(pt, (T)GF_EPS);
    pm = pt - pf_max(V[(size_t)own * F * ld + i], (T)GF_EPS);
    if (clamp) pm = pf_max(pm, (T)GF_EPS);
  }
  Ptot[i] = pt;
  Pminus[i] = pm;
}

// out[r] = sum_c M[r][c] (float64, fixed order): one CTA per row
template <typename T>
__global__ void __launch_bounds__(GF_THREADS)
row_sums_kernel(const T* __restrict__ M, long ldm, long cols, double* __restrict__ out) {
  __shared__ double s_red[GF_THREADS / 32];
  const int r = blockIdx.x;
  double acc = 0.0;
  for (long c = threadIdx.x; c < cols; c += GF_THREADS) acc += (double)M[(size_t)r * ldm + c];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < GF_THREADS / 32; ++w) d += s_red[w];
    out[r] = d;
  }
}

// out (=, +=) a * b  (b == NULL: a), zero in the padding
template <typename T>
__global__ void __launch_bounds__(GF_THREADS)
mul_planes_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int F,
                  long N, long ld, int accumulate) {
  const long n = (long)blockIdx.x * GF_THREADS + threadIdx.x;
  const int f = blockIdx.y;
  if (n >= ld) return;
  const size_t i = (size_t)f * ld + n;
  T v = (T)0;
  if (n < N) {
    v = a[i];
    if (b != nullptr) v *= b[i];
    if (accumulate) v += out[i];
  }
  out[i] = v;
}

// theta[r][c] *= (num[r][c] / max(den[r][c], eps))^omega, num / den in the type of theta
template <typename T>
__global__ void __launch_bounds__(GF_THREADS)
mult_update_same_kernel(T* __restrict__ theta, long ldt, const T* __restrict__ num, long ldn,
                        const T* __restrict__ den, long ldd, int rows, long cols, double omega) {
  const long c = (long)blockIdx.x * GF_THREADS + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= cols) return;
  const double ratio = (double)num[(size_t)r * ldn + c] /
                       fmax((double)den[(size_t)r * ldd + c], GF_EPS);
  const double g = omega == 1.0 ? ratio : pow(ratio, omega);
  theta[(size_t)r * ldt + c] = (T)((double)theta[(size_t)r * ldt + c] * g);
}

// ---- sparsity re-weighting of the source activations -----------------------------------------
// multiChanSourceF0Filter.reweigh_sparsity_constraint (audioModel.py:2981-3014): per frame the
// barycentre of the activations over the dictionary index, median filtered along time
// (tools/signalTools.py:13-24), then a Gaussian mask around it whose width sigma shrinks from
// K^2 to 9 over the iterations (:2937-2977).
template <typename T>
__global__ void sparsity_barycenter_kernel(const T* __restrict__ TW, long ldt, int K, long N,
                                           double* __restrict__ mu) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double num = 0.0, den = 0.0;
  for (int k = 0; k < K - 1; ++k) {
    const double w = (double)(K - 1 - k) * (double)(K - 1 - k);
    const double v = (double)TW[(size_t)k * ldt + n];
    num += (double)k * w * v;
    den += w * fmax(v, GF_EPS);
  }
  mu[n] = num / den;
}

constexpr int GF_MAX_MEDIAN = 64;  // 2 * length
// out[n] = median(in[max(n - L, 0) : min(n + L, N - 1)]), the input value for an empty window
// or a NaN result (signalTools.py:18-23)
__global__ void median_filter_kernel(const double* __restrict__ in, long N, int L,
                                     double* __restrict__ out) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  long lo = n - L, hi = n + L;
  if (lo < 0) lo = 0;
  if (hi > N - 1) hi = N - 1;
  double w[GF_MAX_MEDIAN];
  int cnt = 0;
  bool has_nan = false;
  for (long i = lo; i < hi; ++i) {
    const double v = in[i];
    if (v != v) has_nan = true;
    int j = cnt++;
    while (j > 0 && w[j - 1] > v) {  // insertion sort
      w[j] = w[j - 1];
      --j;
    }
    w[j] = v;
  }
  double m;
  if (cnt == 0 || has_nan)
    m = in[n];
  else
    m = (cnt & 1) ? w[cnt >> 1] : 0.5 * (w[(cnt >> 1) - 1] + w[cnt >> 1]);
  out[n] = m;
}

// TW[k][n] *= mask[k][n]: mask = exp(-(k - mu_n)^2 / (2 sigma)) divided by its maximum over k
// (where that is positive); the last row gets the maximum itself, i.e. 1
template <typename T>
__global__ void sparsity_mask_kernel(T* __restrict__ TW, long ldt, int K, long N,
                                     const double* __restrict__ mu, double log_sigma0,
                                     double slope, const int* __restrict__ iter_dev) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (n >= N) return;
  const double sigma = exp(log_sigma0 + slope * (double)(iter_dev[0] - 1));
  const double m = mu[n];
  // the row nearest to mu carries the maximum (exp is monotone)
  double kc = rint(m);
  if (!(kc >= 0.0)) kc = 0.0;
  if (kc > (double)(K - 1)) kc = (double)(K - 1);
  double best = (kc - m) * (kc - m);
  if (kc >= 1.0) best = fmin(best, (kc - 1.0 - m) * (kc - 1.0 - m));
  if (kc + 1.0 <= (double)(K - 1)) best = fmin(best, (kc + 1.0 - m) * (kc + 1.0 - m));
  const double top = exp(-0.5 * best / sigma);
  double mask;
  if (k == K - 1)
    mask = top > 0.0 ? 1.0 : top;
  else {
    mask = exp(-0.5 * (((double)k - m) * ((double)k - m)) / sigma);
    if (top > 0.0) mask /= top;
  }
  TW[(size_t)k * ldt + n] = (T)((double)TW[(size_t)k * ldt + n] * mask);
}

}  // namespace pf

using namespace pf;

extern "C" int pf_sparsity_reweigh(void* TW, int64_t ldt, int K, int64_t N, int length,
                                   double log_sigma0, double slope, const int* iter_dev,
                                   double* work, int dtype, void* stream) {
  PF_REQUIRE(K > 2 && K <= 65535 && N > 0, "pf_sparsity_reweigh: K=%d N=%ld", K, (long)N);
  PF_REQUIRE(length >= 1 && 2 * length <= GF_MAX_MEDIAN,
             "pf_sparsity_reweigh: median length %d (1..%d)", length, GF_MAX_MEDIAN / 2);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_sparsity_reweigh: bad dtype %d", dtype);
  cudaStream_t st = as_stream(stream);
  double* mu = work;
  double* muf = work + N;
  const int blocks = ceil_div(N, 128);
  if (dtype == PF_F32)
    sparsity_barycenter_kernel<float><<<blocks, 128, 0, st>>>((const float*)TW, ldt, K, N, mu);
  else
    sparsity_barycenter_kernel<double><<<blocks, 128, 0, st>>>((const double*)TW, ldt, K, N, mu);
  median_filter_kernel<<<blocks, 128, 0, st>>>(mu, N, length, muf);
  dim3 grid(ceil_div(N, GF_THREADS), K);
  if (dtype == PF_F32)
    sparsity_mask_kernel<float><<<grid, GF_THREADS, 0, st>>>((float*)TW, ldt, K, N, muf,
                                                            log_sigma0, slope, iter_dev);
  else
    sparsity_mask_kernel<double><<<grid, GF_THREADS, 0, st>>>((double*)TW, ldt, K, N, muf,
                                                             log_sigma0, slope, iter_dev);
  return check_launch("sparsity_mask_kernel");
}

extern "C" int pf_gem_ratio_planes(const void* hatW, const void* P, const void* O, void* out, int F,
                                   int64_t N, int64_t ld, const void* Ptot, const void* Pminus,
                                   double lambda, int dtype, void* stream) {
  PF_REQUIRE(F > 0 && N > 0 && ld >= N, "pf_gem_ratio_planes: F=%d N=%ld ld=%ld", F, (long)N,
             (long)ld);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_gem_ratio_planes: bad dtype %d", dtype);
  PF_REQUIRE((Ptot == nullptr) == (Pminus == nullptr), "pf_gem_ratio_planes: Ptot and Pminus go together");
  dim3 grid(ceil_div(ld, GF_THREADS), F);
  if (dtype == PF_F32)
    gem_ratio_planes_kernel<float><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (const float*)hatW, (const float*)P, (const float*)O, (float*)out, F, N, ld,
        (const float*)Ptot, (const float*)Pminus, lambda);
  else
    gem_ratio_planes_kernel<double><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (const double*)hatW, (const double*)P, (const double*)O, (double*)out, F, N, ld,
        (const double*)Ptot, (const double*)Pminus, lambda);
  return check_launch("gem_ratio_planes_kernel");
}

extern "C" int pf_corr_planes(const void* V, int J, int own, void* Ptot, void* Pminus, int F,
                              int64_t N, int64_t ld, int clamp, int dtype, void* stream) {
  PF_REQUIRE(F > 0 && N > 0 && ld >= N && J > 0 && own >= 0 && own < J,
             "pf_corr_planes: F=%d N=%ld ld=%ld J=%d own=%d", F, (long)N, (long)ld, J, own);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_corr_planes: bad dtype %d", dtype);
  dim3 grid(ceil_div(ld, GF_THREADS), F);
  if (dtype == PF_F32)
    corr_planes_kernel<float><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (const float*)V, J, own, (float*)Ptot, (float*)Pminus, F, N, ld, clamp);
  else
    corr_planes_kernel<double><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (const double*)V, J, own, (double*)Ptot, (double*)Pminus, F, N, ld, clamp);
  return check_launch("corr_planes_kernel");
}

extern "C" int pf_row_sums(const void* M, int64_t ldm, int rows, int64_t cols, double* out,
                           int dtype, void* stream) {
  PF_REQUIRE(rows > 0 && cols > 0 && ldm >= cols, "pf_row_sums: rows=%d cols=%ld ldm=%ld", rows,
             (long)cols, (long)ldm);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_row_sums: bad dtype %d", dtype);
  if (dtype == PF_F32)
    row_sums_kernel<float><<<rows, GF_THREADS, 0, as_stream(stream)>>>((const float*)M, ldm, cols, out);
  else
    row_sums_kernel<double><<<rows, GF_THREADS, 0, as_stream(stream)>>>((const double*)M, ldm, cols, out);
  return check_launch("row_sums_kernel");
}

extern "C" int pf_mul_planes(const void* a, const void* b, void* out, int F, int64_t N, int64_t ld,
                             int accumulate, int dtype, void* stream) {
  PF_REQUIRE(F > 0 && N > 0 && ld >= N, "pf_mul_planes: F=%d N=%ld ld=%ld", F, (long)N, (long)ld);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_mul_planes: bad dtype %d", dtype);
  dim3 grid(ceil_div(ld, GF_THREADS), F);
  if (dtype == PF_F32)
    mul_planes_kernel<float><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (const float*)a, (const float*)b, (float*)out, F, N, ld, accumulate);
  else
    mul_planes_kernel<double><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (const double*)a, (const double*)b, (double*)out, F, N, ld, accumulate);
  return check_launch("mul_planes_kernel");
}

extern "C" int pf_mult_update_same(void* theta, int64_t ldt, const void* num, int64_t ldn,
                                   const void* den, int64_t ldd, int rows, int64_t cols,
                                   double omega, int dtype, void* stream) {
  PF_REQUIRE(rows > 0 && cols > 0, "pf_mult_update_same: rows=%d cols=%ld", rows, (long)cols);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_mult_update_same: bad dtype %d", dtype);
  dim3 grid(ceil_div(cols, GF_THREADS), rows);
  if (dtype == PF_F32)
    mult_update_same_kernel<float><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (float*)theta, ldt, (const float*)num, ldn, (const float*)den, ldd, rows, cols, omega);
  else
    mult_update_same_kernel<double><<<grid, GF_THREADS, 0, as_stream(stream)>>>(
        (double*)theta, ldt, (const double*)num, ldn, (const double*)den, ldd, rows, cols, omega);
  return check_launch("mult_update_same_kernel");
}
