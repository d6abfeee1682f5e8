// K1 / K6 -- STFT front end and inverse STFT with overlap-add, sm_100a.
//
// Replaces the frame loops of pyfasst/tftransforms/stft.py (stft :3-69, istft :71-131)
// and the per-channel transform calls of FASST.comp_transf_Cx (audioModel.py:266-302) /
// FASST.separate_comps (audioModel.py:1188-1217).
//
// Forward: one CTA transforms a tile of TN consecutive frames of one channel.  Each
// frame is framed (half a window of implicit zeros in front, zero fill behind,
// stft.py:40-63), windowed, packed as nfft/2 complex samples, transformed by an in-place
// radix-2 FFT in shared memory (float64) and unpacked to the nfft/2+1 bins of the real
// transform.  The bins of the TN frames are staged in shared memory and written out as
// full 32-byte sectors of the frame-contiguous planes X[2*ch+{0,1}][F][ld].
//
// Inverse: one CTA produces TN*hop output samples of one signal.  It walks the frames
// overlapping its segment in ascending order (the reference's accumulation order,
// stft.py:112-121), loading them in sector-sized groups, inverts each with the same FFT
// and adds the synthesis-windowed samples into a shared-memory segment: overlap-add as a
// gather, deterministic, no atomics.  The segment is then divided by the overlap-added
// window product (stft.py:117-129) and optionally truncated to int16
// (audioModel.py:1227-1229).
#include <stdlib.h>

#include "common.cuh"

namespace pf {

#ifndef PF_FFT_THREADS
#define PF_FFT_THREADS 256
#endif
constexpr int FFT_THREADS = PF_FFT_THREADS;  // (-DPF_FFT_THREADS=...: CTA-shape experiments)

__device__ __forceinline__ double2 cmul_d(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Two consecutive radix-2 stages (half = 1 << lh and 2 half) on the four elements they couple:
// the same butterflies in the same order as two separate passes -- bit-identical results --
// with one shared-memory round trip and one barrier instead of two.
__device__ __forceinline__ void fft_two_stages(double2* base, const double2* __restrict__ tw,
                                               int M, int lh, int q) {
  const int half = 1 << lh;
  const int grp = q >> lh, pos = q & (half - 1);
  const int i0 = (grp << (lh + 2)) + pos, i1 = i0 + half, i2 = i1 + half, i3 = i2 + half;
  const double2 wa = __ldg(tw + pos * (M >> lh));
  const double2 wb0 = __ldg(tw + pos * (M >> (lh + 1)));
  const double2 wb1 = __ldg(tw + (pos + half) * (M >> (lh + 1)));
  const double2 a0 = base[i0], a1 = base[i1], a2 = base[i2], a3 = base[i3];
  const double2 t1 = cmul_d(wa, a1), t3 = cmul_d(wa, a3);
  const double2 p0 = make_double2(a0.x + t1.x, a0.y + t1.y);
  const double2 p1 = make_double2(a0.x - t1.x, a0.y - t1.y);
  const double2 p2 = make_double2(a2.x + t3.x, a2.y + t3.y);
  const double2 p3 = make_double2(a2.x - t3.x, a2.y - t3.y);
  const double2 u2 = cmul_d(wb0, p2), u3 = cmul_d(wb1, p3);
  base[i0] = make_double2(p0.x + u2.x, p0.y + u2.y);
  base[i2] = make_double2(p0.x - u2.x, p0.y - u2.y);
  base[i1] = make_double2(p1.x + u3.x, p1.y + u3.y);
  base[i3] = make_double2(p1.x - u3.x, p1.y - u3.y);
}
__device__ __forceinline__ void fft_one_stage(double2* base, const double2* __restrict__ tw,
                                              int M, int lh, int b) {
  const int half = 1 << lh;
  const int grp = b >> lh, pos = b & (half - 1);
  const int i0 = (grp << (lh + 1)) + pos, i1 = i0 + half;
  const double2 w = __ldg(tw + pos * (M >> lh));
  const double2 t = cmul_d(w, base[i1]);
  const double2 u = base[i0];
  base[i1] = make_double2(u.x - t.x, u.y - t.y);
  base[i0] = make_double2(u.x + t.x, u.y + t.y);
}

// in-place radix-2 decimation-in-time FFT of `nb` independent M-point sequences held bit-reversed,
// back to back, in buf; tw[k] = exp(-2 pi i k / (2M)), k < M.  All threads of the CTA take part;
// the stages run two at a time (one barrier per pair for the whole batch).
__device__ __forceinline__ void fft_inplace_batch(double2* buf, const double2* __restrict__ tw,
                                                  int M, int log2m, int nb) {
  int lh = 0;
  for (; lh + 1 < log2m; lh += 2) {
    for (int idx = threadIdx.x; idx < nb * (M >> 2); idx += FFT_THREADS) {
      const int t = idx >> (log2m - 2), q = idx & ((M >> 2) - 1);
      fft_two_stages(buf + (size_t)t * M, tw, M, lh, q);
    }
    __syncthreads();
  }
  if (lh < log2m) {  // odd number of stages: the last one alone
    for (int idx = threadIdx.x; idx < nb * (M >> 1); idx += FFT_THREADS) {
      const int t = idx >> (log2m - 1), b = idx & ((M >> 1) - 1);
      fft_one_stage(buf + (size_t)t * M, tw, M, lh, b);
    }
    __syncthreads();
  }
}
__device__ __forceinline__ void fft_inplace(double2* buf, const double2* __restrict__ tw, int M) {
  fft_inplace_batch(buf, tw, M, 31 - __clz(M), 1);
}

template <typename T>
struct Sector {  // elements of T in one 32-byte DRAM sector
  static constexpr int N = 32 / sizeof(T);
};

// ---- forward ------------------------------------------------------------------------
// sample (channel ch, time s) of the input in its host layout, as the scaled float64 the
// reference works on (audioObject.py:124-127: data / maxdata)
template <int FMT>
__device__ __forceinline__ double pcm_sample(const void* __restrict__ pcm, int nch, long L, int ch,
                                             long s, double div) {
  if (FMT == PF_PCM_F64_PLANAR) return reinterpret_cast<const double*>(pcm)[(size_t)ch * L + s] / div;
  if (FMT == PF_PCM_I16) return (double)reinterpret_cast<const int16_t*>(pcm)[(size_t)s * nch + ch] / div;
  if (FMT == PF_PCM_I32) return (double)reinterpret_cast<const int32_t*>(pcm)[(size_t)s * nch + ch] / div;
  return (double)reinterpret_cast<const float*>(pcm)[(size_t)s * nch + ch] / div;
}

template <typename T, int FMT>
__global__ void __launch_bounds__(FFT_THREADS)
stft_kernel(const void* __restrict__ pcm, int nch, double div, long L, long sample0, long Ltot,
            long frame0, const double* __restrict__ window, int wlen, int hop, int nfft, int log2m,
            const double2* __restrict__ tw, T* __restrict__ X, int F, long N, long ld, int nb) {
  constexpr int TN = Sector<T>::N;
  constexpr int ROW = 2 * TN + 1;  // padded staging row: (re[TN], im[TN]) per bin
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = nfft / 2;
  double2* buf = reinterpret_cast<double2*>(smem_raw);  // [nb][M]
  T* stage = reinterpret_cast<T*>(smem_raw + (size_t)nb * M * sizeof(double2));
  const int ch = blockIdx.y;
  const long n0 = (long)blockIdx.x * TN;

  // frame n covers samples n*hop - wlen/2 + i, i < wlen (stft.py:47-63); nb frames of the
  // tile (all TN of them when they fit in shared memory) are framed, transformed and unpacked
  // together: one barrier per FFT stage for the whole batch
  for (int t0 = 0; t0 < TN; t0 += nb) {
  for (int idx = threadIdx.x; idx < nb * M; idx += FFT_THREADS) {
    const int tb = idx >> log2m, m = idx & (M - 1);
    const long n = n0 + t0 + tb;
    const long base = (frame0 + n) * hop - wlen / 2;  // global sample index of the frame start
    double v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * m + e;
      const long s = base + i;
      const long sl = s - sample0;  // index into the window of samples we were given
      v[e] = (n < N && i < wlen && s >= 0 && s < Ltot && sl >= 0 && sl < L)
                 ? pcm_sample<FMT>(pcm, nch, L, ch, sl, div) * __ldg(window + i)
                 : 0.0;
    }
    buf[(size_t)tb * M + (__brev((unsigned)m) >> (32 - log2m))] = make_double2(v[0], v[1]);
  }
  __syncthreads();
  fft_inplace_batch(buf, tw, M, log2m, nb);
  // unpack: X[k] = Xe + W^k Xo, Xe = (Z[k]+conj Z[M-k])/2, Xo = -i (Z[k]-conj Z[M-k])/2
  for (int idx = threadIdx.x; idx < nb * F; idx += FFT_THREADS) {
    const int tb = idx / F, k = idx - tb * F, t = t0 + tb;
    const double2* z = buf + (size_t)tb * M;
    const double2 a = z[k & (M - 1)];
    const double2 bq = z[(M - k) & (M - 1)];
    const double2 b = make_double2(bq.x, -bq.y);
    const double2 xe = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
    const double2 d = make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y));
    const double2 xo = make_double2(d.y, -d.x);  // -i * d
    const double2 w = (k < M) ? __ldg(tw + k) : make_double2(-1.0, 0.0);
    const double2 r = cmul_d(w, xo);
    stage[k * ROW + t] = (T)(xe.x + r.x);
    stage[k * ROW + TN + t] = (T)(xe.y + r.y);
  }
  __syncthreads();
  }
  // write the tile: per (bin, re/im) one 32-byte sector
  const size_t plane = (size_t)F * ld;
  for (int i = threadIdx.x; i < F * 2 * TN; i += FFT_THREADS) {
    const int t = i % TN;
    const int c = (i / TN) & 1;
    const int k = i / (2 * TN);
    const long n = n0 + t;
    if (n < ld) X[(size_t)(2 * ch + c) * plane + (size_t)k * ld + n] = stage[k * ROW + c * TN + t];
  }
}

// norm[t] = sum over the frames n covering sample t, in ascending n (the reference's accumulation
// order, stft.py:112-121), of prod[t - n hop]: the overlap-added window product of istft
__global__ void overlap_norm_kernel(const double* __restrict__ prod, int wlen, int hop, long N,
                                    long total, double* __restrict__ norm) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  long n_lo = t - wlen + 1;
  n_lo = n_lo <= 0 ? 0 : (n_lo + hop - 1) / hop;
  long n_hi = t / hop;
  if (n_hi > N - 1) n_hi = N - 1;
  double acc = 0.0;
  for (long n = n_lo; n <= n_hi; ++n) acc += prod[t - n * hop];
  norm[t] = acc;
}

// psd_sum[f] = sum over planes and frames of X^2 (audioModel.py:304-319), fixed order
template <typename T>
__global__ void psd_sum_kernel(const T* __restrict__ X, int nplanes, int F, long N, long ld,
                               double* __restrict__ out) {
  const int f = blockIdx.x;
  double acc = 0.0;
  for (int p = 0; p < nplanes; ++p) {
    const T* row = X + ((size_t)p * F + f) * ld;
    for (long n = threadIdx.x; n < N; n += blockDim.x) {
      const double v = (double)row[n];
      acc += v * v;
    }
  }
  __shared__ double s_red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
    out[f] = d;
  }
}

// peak[0] = max |x| over the samples as np.abs(data).max() gives it (audioObject.py:124-126):
// for integer PCM the most negative value is skipped (np.abs wraps it onto itself, so it never
// wins the max in the reference).  Non-negative doubles order like their bit patterns, so the
// block maxima are combined with an integer atomicMax -- order independent, deterministic.
template <typename S>
__global__ void pcm_peak_kernel(const S* __restrict__ x, long count, S skip, int has_skip,
                                unsigned long long* __restrict__ peak) {
  double m = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long)gridDim.x * blockDim.x) {
    const S v = x[i];
    if (has_skip && v == skip) continue;
    const double a = fabs((double)v);
    m = a > m ? a : m;
  }
  m = warp_max(m);
  __shared__ double s_red[32];
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = s_red[w] > m ? s_red[w] : m;
    atomicMax(peak, (unsigned long long)__double_as_longlong(m));
  }
}

// ---- inverse --------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(FFT_THREADS)
istft_kernel(const T* __restrict__ Y, int F, long N, long ld, const double* __restrict__ synth,
             const double* __restrict__ norm, int wlen, int hop, int nfft, int log2m,
             const double2* __restrict__ tw, int seg_frames, double* __restrict__ out, long Lout,
             int16_t* __restrict__ pcm, int nsig, double maxdata, long drop, int pcm_round) {
  constexpr int SG = Sector<T>::N;
  constexpr int ROW = 2 * SG + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = nfft / 2;
  const int seg_len = seg_frames * hop;
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  double* seg = reinterpret_cast<double*>(smem_raw + (size_t)M * sizeof(double2));
  T* stage = reinterpret_cast<T*>(smem_raw + (size_t)M * sizeof(double2) +
                                  (size_t)seg_len * sizeof(double));
  const int sig = blockIdx.y;
  const long tau0 = (long)blockIdx.x * seg_len;  // padded time of the first sample
  const long tau1 = tau0 + seg_len;
  // frames n with n*hop < tau1 and n*hop + wlen > tau0
  long n_lo = (tau0 - wlen >= 0) ? (tau0 - wlen) / hop + 1 : 0;
  long n_hi = (tau1 - 1) / hop;
  if (n_hi > N - 1) n_hi = N - 1;
  for (int i = threadIdx.x; i < seg_len; i += FFT_THREADS) seg[i] = 0.0;
  const size_t plane = (size_t)F * ld;
  const T* Yre = Y + (size_t)(2 * sig) * plane;
  const T* Yim = Yre + plane;
  const double invM = 1.0 / (double)M;

  for (long g0 = (n_lo / SG) * SG; g0 <= n_hi; g0 += SG) {
    __syncthreads();
    // stage the group's bins: one sector per (bin, re/im)
    for (int i = threadIdx.x; i < F * 2 * SG; i += FFT_THREADS) {
      const int t = i % SG;
      const int c = (i / SG) & 1;
      const int k = i / (2 * SG);
      const long n = g0 + t;
      T v = (T)0;
      if (n < N) v = (c ? Yim : Yre)[(size_t)k * ld + n];
      stage[k * ROW + c * SG + t] = v;
    }
    __syncthreads();
    for (int t = 0; t < SG; ++t) {
      const long n = g0 + t;
      if (n < n_lo || n > n_hi) continue;  // uniform per CTA
      // pack: Z[k] = Xe + i Xo, Xe = (X[k]+conj X[M-k])/2, Xo = conj(W^k) (X[k]-conj X[M-k])/2;
      // the imaginary parts of X[0] and X[M] are ignored like numpy's irfft does
      for (int k = threadIdx.x; k < M; k += FFT_THREADS) {
        double2 a = make_double2((double)stage[k * ROW + t], (double)stage[k * ROW + SG + t]);
        double2 b = make_double2((double)stage[(M - k) * ROW + t],
                                 -(double)stage[(M - k) * ROW + SG + t]);
        if (k == 0) { a.y = 0.0; b.y = 0.0; }
        const double2 xe = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
        const double2 d = make_double2(0.5 * (a.x - b.x), 0.5 * (a.y - b.y));
        const double2 wq = __ldg(tw + k);
        const double2 xo = cmul_d(make_double2(wq.x, -wq.y), d);
        // Z = xe + i xo ; store conj(Z) so the forward FFT yields conj(IFFT) * M
        buf[__brev((unsigned)k) >> (32 - log2m)] = make_double2(xe.x - xo.y, -(xe.y + xo.x));
      }
      __syncthreads();
      fft_inplace(buf, tw, M);
      // x[2m] = Re z[m], x[2m+1] = Im z[m], z = conj(buf)/M ; overlap-add (stft.py:112-116)
      const long fr0 = n * hop;
      for (int m = threadIdx.x; m < M; m += FFT_THREADS) {
        const double2 z = buf[m];
        const double xs[2] = {z.x * invM, -z.y * invM};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * m + e;
          const long tau = fr0 + i;
          if (i < wlen && tau >= tau0 && tau < tau1) seg[tau - tau0] += __ldg(synth + i) * xs[e];
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < seg_len; i += FFT_THREADS) {
    const long tau = tau0 + i;
    // FASST drops the first half window (stft.py:123, drop = wlen/2); the SIMM back end keeps
    // it (separateLeadFunctions.py:224-232, drop = 0)
    const long tout = tau - drop;
    if (tout < 0 || tout >= Lout) continue;
    const double v = seg[i] / __ldg(norm + tau);
    out[(size_t)sig * Lout + tout] = v;
    // np.int16(y * maxdata) truncates (audioModel.py:1227); np.round then cast rounds half to
    // even (SeparateLeadStereoTF.py:1826-1827)
    if (pcm != nullptr)
      pcm[(size_t)tout * nsig + sig] =
          (int16_t)(int)(pcm_round ? rint(v * maxdata) : v * maxdata);
  }
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// twiddle table exp(-2 pi i k / nfft), k < nfft/2, cached per nfft on the device
struct TwiddleCache {
  int nfft = 0;
  int device = -1;
  double2* ptr = nullptr;
};
static TwiddleCache g_tw[4];

static const double2* twiddles(int nfft, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& c : g_tw)
    if (c.nfft == nfft && c.device == dev) return c.ptr;
  TwiddleCache* slot = nullptr;
  for (auto& c : g_tw)
    if (c.ptr == nullptr) { slot = &c; break; }
  if (slot == nullptr) {
    slot = &g_tw[0];
    cudaFree(slot->ptr);
    slot->ptr = nullptr;
  }
  const int M = nfft / 2;
  double2* host = (double2*)malloc(sizeof(double2) * M);
  for (int k = 0; k < M; ++k) {
    const double ang = -2.0 * 3.14159265358979323846 * (double)k / (double)nfft;
    host[k] = make_double2(cos(ang), sin(ang));
  }
  if (cudaMalloc(&slot->ptr, sizeof(double2) * M) != cudaSuccess) {
    free(host);
    slot->ptr = nullptr;
    return nullptr;
  }
  // synchronous copy: the table must be complete before host memory is released
  cudaMemcpy(slot->ptr, host, sizeof(double2) * M, cudaMemcpyHostToDevice);
  free(host);
  slot->nfft = nfft;
  slot->device = dev;
  (void)st;
  return slot->ptr;
}

template <typename T, int FMT>
static int launch_stft(const void* pcm, int nch, double div, long L, long sample0, long Ltot,
                       long frame0, const double* window, int wlen, int hop, int nfft, void* X,
                       long N, long ld, double* psd_sum, cudaStream_t st) {
  constexpr int TN = Sector<T>::N;
  const int M = nfft / 2, F = M + 1;
  const double2* tw = twiddles(nfft, st);
  if (tw == nullptr) {
    set_error("pf_stft: cannot allocate the twiddle table");
    return PF_ERR_CUDA;
  }
  const size_t stage_bytes = (size_t)F * (2 * TN + 1) * sizeof(T);
  // frames transformed together (one barrier per FFT stage for the batch); limited so that two
  // CTAs stay resident per SM
  int nb = TN;
  size_t budget = 110 * 1024;
  if (const char* e = getenv("PYFASST_STFT_NB")) {  // tuning override
    nb = atoi(e) > 0 ? atoi(e) : TN;
    if (nb > TN) nb = TN;
    budget = 200 * 1024;
  }
  while (nb > 1 && (size_t)nb * M * sizeof(double2) + stage_bytes > budget) nb /= 2;
  const size_t smem = (size_t)nb * M * sizeof(double2) + stage_bytes;
  cudaError_t e = cudaFuncSetAttribute(stft_kernel<T, FMT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("pf_stft: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  dim3 grid(ceil_div(N, TN), nch);
  stft_kernel<T, FMT><<<grid, FFT_THREADS, smem, st>>>(pcm, nch, div, L, sample0, Ltot, frame0,
                                                      window, wlen, hop, nfft, ilog2(M), tw,
                                                      (T*)X, F, N, ld, nb);
  int rc = check_launch("stft_kernel");
  if (rc) return rc;
  if (psd_sum != nullptr) {
    psd_sum_kernel<T><<<F, 256, 0, st>>>((const T*)X, 2 * nch, F, N, ld, psd_sum);
    rc = check_launch("psd_sum_kernel");
  }
  return rc;
}

template <typename T>
static int dispatch_stft(int fmt, const void* pcm, int nch, double div, long L, long sample0,
                         long Ltot, long frame0, const double* window, int wlen, int hop, int nfft,
                         void* X, long N, long ld, double* psd_sum, cudaStream_t st) {
#define PF_STFT_CASE(F_)                                                                       \
  case F_:                                                                                     \
    return launch_stft<T, F_>(pcm, nch, div, L, sample0, Ltot, frame0, window, wlen, hop, nfft, \
                              X, N, ld, psd_sum, st);
  switch (fmt) {
    PF_STFT_CASE(PF_PCM_F64_PLANAR)
    PF_STFT_CASE(PF_PCM_I16)
    PF_STFT_CASE(PF_PCM_I32)
    PF_STFT_CASE(PF_PCM_F32)
  }
#undef PF_STFT_CASE
  set_error("pf_stft: unknown PCM format %d", fmt);
  return PF_ERR_ARG;
}

template <typename T>
static int launch_istft(const void* Y, int nsig, int F, long N, long ld, const double* synth,
                        const double* norm, int wlen, int hop, int nfft, double* out, long Lout,
                        int16_t* pcm, double maxdata, long drop, int pcm_round, cudaStream_t st) {
  constexpr int SG = Sector<T>::N;
  const int M = nfft / 2;
  const double2* tw = twiddles(nfft, st);
  if (tw == nullptr) {
    set_error("pf_istft: cannot allocate the twiddle table");
    return PF_ERR_CUDA;
  }
  // segment of ~16 frames: 16/(16 + wlen/hop - 1) of the inverse FFTs are not redundant
  int seg_frames = 16;
  const size_t fixed = (size_t)M * sizeof(double2) + (size_t)F * (2 * SG + 1) * sizeof(T);
  while (seg_frames > 1 && fixed + (size_t)seg_frames * hop * sizeof(double) > 200 * 1024)
    seg_frames /= 2;
  const size_t smem = fixed + (size_t)seg_frames * hop * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(istft_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) {
    set_error("pf_istft: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  const long total = (N - 1) * hop + wlen;
  dim3 grid(ceil_div(total, (long)seg_frames * hop), nsig);
  istft_kernel<T><<<grid, FFT_THREADS, smem, st>>>((const T*)Y, F, N, ld, synth, norm, wlen, hop,
                                                  nfft, ilog2(M), tw, seg_frames, out, Lout, pcm,
                                                  nsig, maxdata, drop, pcm_round);
  return check_launch("istft_kernel");
}

}  // namespace pf

using namespace pf;

static int check_fft_args(const char* who, int wlen, int hop, int nfft) {
  PF_REQUIRE(nfft >= 16 && nfft <= 4096 && (nfft & (nfft - 1)) == 0,
             "%s: nfft=%d must be a power of two in [16, 4096]", who, nfft);
  PF_REQUIRE(wlen >= 2 && wlen <= nfft && wlen % 2 == 0, "%s: wlen=%d (nfft=%d)", who, wlen, nfft);
  PF_REQUIRE(hop >= 1 && hop <= wlen, "%s: hop=%d (wlen=%d)", who, hop, wlen);
  return PF_OK;
}

extern "C" int pf_stft(const void* pcm, int pcm_format, double pcm_div, int nch, int64_t L,
                       int64_t sample0, int64_t L_total, const double* window, int wlen, int hop,
                       int nfft, void* X, int64_t frame0, int64_t N, int64_t ld, double* psd_sum,
                       int dtype, void* stream) {
  int rc = check_fft_args("pf_stft", wlen, hop, nfft);
  if (rc) return rc;
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_stft: bad dtype %d", dtype);
  PF_REQUIRE(nch >= 1 && nch <= 16 && L > 0, "pf_stft: nch=%d L=%ld", nch, (long)L);
  PF_REQUIRE(pcm_div != 0.0, "pf_stft: pcm_div must not be zero");
  PF_REQUIRE(sample0 >= 0 && sample0 + L <= L_total, "pf_stft: samples [%ld, %ld) outside [0, %ld)",
             (long)sample0, (long)(sample0 + L), (long)L_total);
  const int64_t n_total = (L_total + hop - 1) / hop + 2;  // stft.py:40
  PF_REQUIRE(frame0 >= 0 && N >= 1 && frame0 + N <= n_total,
             "pf_stft: frames [%ld, %ld) outside [0, ceil(L/hop)+2 = %ld)", (long)frame0,
             (long)(frame0 + N), (long)n_total);
  PF_REQUIRE(ld >= N && ld % 4 == 0, "pf_stft: ld=%ld must be >= N and a multiple of 4", (long)ld);
  cudaStream_t st = as_stream(stream);
  if (dtype == PF_F32)
    return dispatch_stft<float>(pcm_format, pcm, nch, pcm_div, L, sample0, L_total, frame0, window,
                                wlen, hop, nfft, X, N, ld, psd_sum, st);
  return dispatch_stft<double>(pcm_format, pcm, nch, pcm_div, L, sample0, L_total, frame0, window,
                               wlen, hop, nfft, X, N, ld, psd_sum, st);
}

extern "C" int pf_pcm_peak(const void* pcm, int pcm_format, int64_t count, double* peak,
                           void* stream) {
  PF_REQUIRE(count > 0, "pf_pcm_peak: empty signal");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(peak, 0, sizeof(double), st);
  if (e != cudaSuccess) {
    set_error("pf_pcm_peak: %s", cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  unsigned long long* out = reinterpret_cast<unsigned long long*>(peak);
  const int grid = 148 * 8;
  switch (pcm_format) {
    case PF_PCM_F64_PLANAR:
      pcm_peak_kernel<double><<<grid, 256, 0, st>>>((const double*)pcm, count, 0.0, 0, out);
      break;
    case PF_PCM_I16:
      pcm_peak_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)pcm, count, INT16_MIN, 1, out);
      break;
    case PF_PCM_I32:
      pcm_peak_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t*)pcm, count, INT32_MIN, 1, out);
      break;
    case PF_PCM_F32:
      pcm_peak_kernel<float><<<grid, 256, 0, st>>>((const float*)pcm, count, 0.f, 0, out);
      break;
    default:
      set_error("pf_pcm_peak: unknown PCM format %d", pcm_format);
      return PF_ERR_ARG;
  }
  return check_launch("pcm_peak_kernel");
}

extern "C" int pf_istft(const void* Y, int nsig, int F, int64_t N, int64_t ld,
                        const double* synth, const double* norm, int wlen, int hop, int nfft,
                        double* out, int64_t Lout, int16_t* pcm, double maxdata, int64_t drop,
                        int pcm_round, int dtype, void* stream) {
  int rc = check_fft_args("pf_istft", wlen, hop, nfft);
  if (rc) return rc;
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_istft: bad dtype %d", dtype);
  PF_REQUIRE(F == nfft / 2 + 1, "pf_istft: F=%d != nfft/2+1", F);
  PF_REQUIRE(nsig >= 1 && nsig <= 65535 && N >= 1 && Lout >= 1, "pf_istft: nsig=%d N=%ld", nsig,
             (long)N);
  PF_REQUIRE(ld >= N && ld % 4 == 0, "pf_istft: ld=%ld must be >= N and a multiple of 4", (long)ld);
  PF_REQUIRE(drop >= 0 && drop <= wlen, "pf_istft: drop=%ld outside [0, wlen]", (long)drop);
  cudaStream_t st = as_stream(stream);
  if (dtype == PF_F32)
    return launch_istft<float>(Y, nsig, F, N, ld, synth, norm, wlen, hop, nfft, out, Lout, pcm,
                               maxdata, drop, pcm_round, st);
  return launch_istft<double>(Y, nsig, F, N, ld, synth, norm, wlen, hop, nfft, out, Lout, pcm,
                              maxdata, drop, pcm_round, st);
}

extern "C" int pf_overlap_norm(const double* prod, int wlen, int hop, int64_t N, double* norm,
                               void* stream) {
  PF_REQUIRE(wlen > 0 && hop > 0 && N > 0, "pf_overlap_norm: wlen=%d hop=%d N=%ld", wlen, hop,
             (long)N);
  const long total = (long)hop * (N - 1) + wlen;
  overlap_norm_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(prod, wlen, hop, N,
                                                                         total, norm);
  return check_launch("overlap_norm_kernel");
}
