// Glottal-source harmonic combs: the F0 dictionary WF0 of the SIMM source/filter model, sm_100a.
//
// Replaces the per-F0 Python loops of pyfasst/SeparateLeadStereo/separateLeadFunctions.py:
// generate_ODGD_spec (:888-949), generate_ODGD_spec_chirped (:1010-1072) and the column loops of
// generate_WF0_chirped (:237-345) / generate_WF0_TR_chirped (:696-886) ("horribly slow", :760).
//
// Column c is the power spectrum of one windowed frame of the KLGLOTT88 waveform
//   x_c(t) = Re sum_{h=1}^{P_c} A_h exp(2 pi i h (F1_c tau + (F2_c - F1_c) tau^2 / (2 T))),
//   tau = t / fs, T = Lsig / fs, A_h = F0 27/4 (E + 2 (1 + 2E)/z - 6 (1 - E)/z^2) / z,
//   z = 2 pi i h Ot, E = exp(-z), F0 = (F1 + F2) / 2          (F1 = F2: the plain comb).
// One CTA per column: (1) the P_c complex amplitudes, (2) the windowed frame -- per sample one
// sincos and a geometric recurrence over the partials (the phase is linear in h) instead of
// P_c complex exponentials, (3) a direct DFT of the frame against a shared-memory twiddle table.
// Everything in float64; ~2e9 FP64 operations for the reference's default dictionary (1092
// columns x 2048 samples x ~140 partials), a few milliseconds on a B200.
#include "common.cuh"

namespace pf {

constexpr int WF0_THREADS = 256;
constexpr int WF0_MAX_PARTIALS = 2048;
constexpr int WF0_MAX_NFFT = 8192;

__global__ void __launch_bounds__(WF0_THREADS)
wf0_comb_kernel(const double* __restrict__ f1, const double* __restrict__ f2,
                const int* __restrict__ npart, double fs, double Ot, long Lsig, long t_begin,
                const double* __restrict__ window, int wlen, int nfft, int rows,
                double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char wf0_smem[];
  double2* s_tw = reinterpret_cast<double2*>(wf0_smem);  // [nfft]  e^{+2 pi i m / nfft}
  double* s_frame = reinterpret_cast<double*>(s_tw + nfft);  // [wlen]
  double2* s_amp = reinterpret_cast<double2*>(s_frame + wlen);  // [P]
  const int c = blockIdx.x;
  const double F1 = f1[c], F2 = f2[c];
  const double F0 = (F1 + F2) / 2.0;
  const int P = npart[c];
  const double kPi = 3.141592653589793;

  for (int m = threadIdx.x; m < nfft; m += WF0_THREADS) {
    double s, co;
    sincospi(2.0 * (double)m / (double)nfft, &s, &co);
    s_tw[m] = make_double2(co, s);
  }
  // amplitudes of the partials (separateLeadFunctions.py:917-929): with z = i th,
  //   S = E + 2 (1 + 2E) / z - 6 (1 - E) / z^2,  E = cos th - i sin th,  A = F0 27/4 S / z
  for (int h = 1 + threadIdx.x; h <= P; h += WF0_THREADS) {
    const double th = ((2.0 * kPi) * (double)h) * Ot;
    double s, co;
    sincos(th, &s, &co);
    const double sre = co - 4.0 * s / th + 6.0 * (1.0 - co) / (th * th);
    const double sim = -s - 2.0 * (1.0 + 2.0 * co) / th + 6.0 * s / (th * th);
    const double g = F0 * 27.0 / 4.0 / th;
    s_amp[h - 1] = make_double2(g * sim, -g * sre);  // S / (i th) = (S_im - i S_re) / th
  }
  __syncthreads();
  // the windowed frame: samples t_begin .. t_begin + wlen - 1 of the waveform (zero outside it)
  const double T2 = 2.0 * (double)Lsig / fs;
  for (int k = threadIdx.x; k < wlen; k += WF0_THREADS) {
    const long t = t_begin + k;
    double v = 0.0;
    if (t >= 0 && t < Lsig && P > 0) {
      const double tau = (double)t / fs;
      const double cyc = F1 * tau + (F2 - F1) * (tau * tau) / T2;  // cycles of the fundamental
      double s, co;
      sincospi(2.0 * (cyc - rint(cyc)), &s, &co);
      double pr = co, pi = s;  // w^h, h = 1
      double acc = s_amp[0].x * pr - s_amp[0].y * pi;
      for (int h = 1; h < P; ++h) {
        const double nr = pr * co - pi * s;
        pi = pr * s + pi * co;
        pr = nr;
        acc += s_amp[h].x * pr - s_amp[h].y * pi;
      }
      v = acc * window[k];
    }
    s_frame[k] = v;
  }
  __syncthreads();
  // |DFT|^2 of the zero-padded frame, rows 0 .. rows-1
  for (int r = threadIdx.x; r < rows; r += WF0_THREADS) {
    double re = 0.0, im = 0.0;
    int m = 0;
    const int step = r % nfft;
    for (int k = 0; k < wlen; ++k) {
      const double2 w = s_tw[m];
      const double x = s_frame[k];
      re += x * w.x;
      im -= x * w.y;
      m += step;
      if (m >= nfft) m -= nfft;
    }
    out[(size_t)c * rows + r] = re * re + im * im;
  }
}

}  // namespace pf

using namespace pf;

extern "C" int pf_wf0_combs(const double* f1, const double* f2, const int* npart, int ncols,
                            int max_partials, double fs, double Ot, int64_t Lsig, int64_t t_begin,
                            const double* window, int wlen, int nfft, int rows, double* out,
                            void* stream) {
  PF_REQUIRE(ncols > 0 && rows > 0 && Lsig > 0, "pf_wf0_combs: empty problem (ncols=%d rows=%d)",
             ncols, rows);
  PF_REQUIRE(wlen > 0 && wlen <= nfft && nfft <= WF0_MAX_NFFT,
             "pf_wf0_combs: need 0 < wlen=%d <= nfft=%d <= %d", wlen, nfft, WF0_MAX_NFFT);
  PF_REQUIRE(max_partials >= 0 && max_partials <= WF0_MAX_PARTIALS,
             "pf_wf0_combs: %d partials (max %d): F0 too low for this sampling rate",
             max_partials, WF0_MAX_PARTIALS);
  PF_REQUIRE(fs > 0.0 && Ot > 0.0, "pf_wf0_combs: fs=%g Ot=%g", fs, Ot);
  const size_t smem = (size_t)nfft * 16 + (size_t)wlen * 8 + (size_t)(max_partials + 1) * 16;
  cudaError_t e = cudaFuncSetAttribute(wf0_comb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) {
    set_error("pf_wf0_combs: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  wf0_comb_kernel<<<ncols, WF0_THREADS, smem, as_stream(stream)>>>(
      f1, f2, npart, fs, Ot, (long)Lsig, (long)t_begin, window, wlen, nfft, rows, out);
  return check_launch("wf0_comb_kernel");
}
