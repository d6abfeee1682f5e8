// GEM-loop glue kernels: annealed noise PSD, log-likelihood bookkeeping, the small
// factor products FB.FW.  Everything the loop of FASST.estim_param_a_post_model
// (pyfasst/audioModel.py:330-382) does between the big kernels stays on the device so
// that an iteration never synchronises with the host and can be captured in a CUDA
// graph (the iteration index lives in device memory).
#include "common.cuh"

namespace pf {

// noise[f] = ((sqrt(lim0)*(I-i) + sqrt(lim1)*i)/I)^2      (audioModel.py:368-373)
__global__ void noise_anneal_kernel(const double* __restrict__ s0, const double* __restrict__ s1,
                                    const int* __restrict__ iter_dev, int n_iter, int F,
                                    double* __restrict__ noise) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const double i = (double)(*iter_dev);
  const double I = (double)n_iter;
  const double v = (s0[f] * (I - i) + s1[f] * i) / I;
  noise[f] = v * v;
}

// single CTA, fixed-order tree: deterministic
__global__ void ll_reduce_kernel(const double* __restrict__ ll_f, int F, double* __restrict__ out) {
  __shared__ double s_red[32];
  double acc = 0.0;
  for (int f = threadIdx.x; f < F; f += blockDim.x) acc += ll_f[f];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
    out[0] = d;
  }
}

__global__ void ll_store_kernel(const double* __restrict__ ll_sum, double bins,
                                double* __restrict__ logliks, int* __restrict__ iter_dev,
                                int advance) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int i = *iter_dev;
    logliks[i] = -ll_sum[0] / bins;  // audioModel.py:660-664 (mean over F*N)
    if (advance) *iter_dev = i + 1;
  }
}

template <typename T>
__global__ void small_matmul_kernel(const T* __restrict__ A, int lda, const T* __restrict__ B,
                                    int ldb, T* __restrict__ C, int ldc, int M, int K, int Nc) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * Nc) return;
  const int m = (int)(idx / Nc), n = (int)(idx % Nc);
  // accumulate in double in index order, like a dot product of float64 rows
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * lda + k] * (double)B[(size_t)k * ldb + n];
  C[(size_t)m * ldc + n] = (T)acc;
}

// sum(TW) < eps after the renormalisation: the reference re-draws that TW at random
// (audioModel.py:2023-2025).  The loop itself never synchronises with the host: the flag and the
// first iteration at which it happened are recorded, and the host replays from there.
__global__ void check_totals_kernel(double* __restrict__ totals, int count, double eps,
                                    int* __restrict__ flags, const int* __restrict__ iter_dev,
                                    int* __restrict__ first_iter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (totals[i] < eps) {
    atomicOr(flags, PF_FLAG_TW_RESTART);
    if (first_iter != nullptr) atomicMin(first_iter, iter_dev != nullptr ? *iter_dev : 0);
  }
  totals[i] = 0.0;
}

// Y[c1] = sum_c2 W[c1][c2] X[c2] per time-frequency bin (tftransforms/stft.py:181-193): X, Y planes
// [2 nc][F][ld] (re, im per channel), W complex128 [nc][nc][F] (wn = 0) or [nc][nc][F][wn]
template <typename T>
__global__ void apply_filter_kernel(const T* __restrict__ X, const double2* __restrict__ W,
                                    T* __restrict__ Y, int nc, int F, long N, long ld, long wn) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (n >= ld) return;
  const size_t plane = (size_t)F * ld, i = (size_t)f * ld + n;
  for (int c1 = 0; c1 < nc; ++c1) {
    double yr = 0.0, yi = 0.0;
    if (n < N) {
      for (int c2 = 0; c2 < nc; ++c2) {
        const size_t wi = ((size_t)c1 * nc + c2) * F + f;
        const double2 w = wn > 0 ? W[wi * wn + n] : W[wi];
        const double xr = (double)X[(size_t)(2 * c2) * plane + i];
        const double xi = (double)X[(size_t)(2 * c2 + 1) * plane + i];
        yr += w.x * xr - w.y * xi;
        yi += w.x * xi + w.y * xr;
      }
    }
    Y[(size_t)(2 * c1) * plane + i] = (T)yr;
    Y[(size_t)(2 * c1 + 1) * plane + i] = (T)yi;
  }
}

}  // namespace pf

using namespace pf;

extern "C" int pf_apply_filter(const void* X, const void* W, void* Y, int nc, int F, int64_t N,
                               int64_t ld, int64_t wn, int dtype, void* stream) {
  PF_REQUIRE(nc >= 1 && nc <= 8 && F > 0 && N > 0 && ld >= N, "pf_apply_filter: nc=%d F=%d N=%ld ld=%ld",
             nc, F, (long)N, (long)ld);
  PF_REQUIRE(wn == 0 || wn >= N, "pf_apply_filter: W has %ld frames, X has %ld", (long)wn, (long)N);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_apply_filter: bad dtype %d", dtype);
  dim3 grid(ceil_div(ld, 256), F);
  if (dtype == PF_F32)
    apply_filter_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)X, (const double2*)W, (float*)Y, nc, F, N, ld, wn);
  else
    apply_filter_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
        (const double*)X, (const double2*)W, (double*)Y, nc, F, N, ld, wn);
  return check_launch("apply_filter_kernel");
}

extern "C" int pf_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("pf_set_device(%d): %s", device, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  return PF_OK;
}

extern "C" int pf_check_totals(double* totals, int count, double eps, int* flags,
                               const int* iter_dev, int* first_iter, void* stream) {
  PF_REQUIRE(count > 0, "pf_check_totals: count=%d", count);
  check_totals_kernel<<<ceil_div(count, 64), 64, 0, as_stream(stream)>>>(totals, count, eps, flags,
                                                                        iter_dev, first_iter);
  return check_launch("check_totals_kernel");
}

extern "C" int pf_noise_anneal(const double* sqrt_lim0, const double* sqrt_lim1,
                               const int* iter_dev, int n_iter, int F, double* noise,
                               void* stream) {
  PF_REQUIRE(F > 0 && n_iter > 0, "pf_noise_anneal: F=%d n_iter=%d", F, n_iter);
  noise_anneal_kernel<<<ceil_div(F, 256), 256, 0, as_stream(stream)>>>(sqrt_lim0, sqrt_lim1,
                                                                      iter_dev, n_iter, F, noise);
  return check_launch("noise_anneal_kernel");
}

extern "C" int pf_ll_reduce(const double* ll_f, int F, double* ll_sum, void* stream) {
  PF_REQUIRE(F > 0, "pf_ll_reduce: F=%d", F);
  ll_reduce_kernel<<<1, 256, 0, as_stream(stream)>>>(ll_f, F, ll_sum);
  return check_launch("ll_reduce_kernel");
}

extern "C" int pf_ll_store(const double* ll_sum, double bins, double* logliks, int* iter_dev,
                           int advance, void* stream) {
  PF_REQUIRE(bins > 0, "pf_ll_store: bins=%g", bins);
  ll_store_kernel<<<1, 32, 0, as_stream(stream)>>>(ll_sum, bins, logliks, iter_dev, advance);
  return check_launch("ll_store_kernel");
}

extern "C" int pf_small_matmul(const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                               int M, int K, int Nc, int dtype, void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_small_matmul: bad dtype %d", dtype);
  PF_REQUIRE(M > 0 && K > 0 && Nc > 0, "pf_small_matmul: empty problem");
  const int grid = ceil_div((long)M * Nc, 256);
  if (dtype == PF_F32)
    small_matmul_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)A, lda, (const float*)B, ldb, (float*)C, ldc, M, K, Nc);
  else
    small_matmul_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
        (const double*)A, lda, (const double*)B, ldb, (double*)C, ldc, M, K, Nc);
  return check_launch("small_matmul_kernel");
}
