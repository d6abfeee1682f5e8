// K7 / K8 -- the bandwidth-bound kernels of the SIMM / Stereo_SIMM source/filter model, sm_100a.
//
// Replaces the elementwise, reduction and small-matrix parts of
// pyfasst/SeparateLeadStereo/SIMM/SIMM.py: SIMM (:46-395, update loop :303-393) and
// Stereo_SIMM (:397-943, update loop :613-941).  The dense contractions (every np.dot with an
// F x N operand) run on the tensor cores through pf_gemm_tf32x3 / pf_gemm_tf32x3_splitk.
//
// Model:  hat_c = alpha_c^2 (SF0 * SPHI) + (WM beta_c^2) HM,  c in {R, L}; the mono model is the
// one-channel case with alpha = beta = 1.
//
// Layout (float32, frames contiguous, ldn = N rounded up to 4):
//   SF0, SPHI          [F][ldn]
//   SX, hat, SM        [F][nch * ldn]      channel c in columns [c ldn, c ldn + N)
//   work               [F][2 nch ldn]      (num | den) or (T_0 .. | I_0 ..)
//   HF0 [NF0][ldn], HPHI [K][ldn], HM [R][ldn], WM [F][ldr], ...
// Every producer writes ZERO into the padding columns n >= N of the work planes, so they can
// be contracted over n by the GEMMs; the padding of the H matrices stays zero (updates touch
// n < N only).
#include "common.cuh"

namespace pf {

constexpr float SIMM_EPS = 1e-20f;  // SIMM.py:150, :497 (not audioModel's 1e-10)
constexpr int SE_THREADS = 256;

__device__ __forceinline__ float4 ld4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void unpack(const float4 v, float (&o)[4]) {
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

// fixed-order block sum of a double (all threads receive nothing; thread 0 gets the result)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double d = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < THREADS / 32; ++w) d += s_red[w];
  return d;
}

// hat_c = max(a2_c SF0 SPHI + SM_c, eps) for 4 consecutive frames of row f, formed in registers
// by every consumer: the hat planes are never stored (SIMM.py:313; Stereo :655-664).
template <int NCH>
__device__ __forceinline__ void load_model(const float* __restrict__ SM,
                                           const float* __restrict__ SF0,
                                           const float* __restrict__ SPHI,
                                           const float* __restrict__ a2, long f, long n0, long ldn,
                                           float (&s0)[4], float (&sp)[4], float (&h)[NCH][4]) {
  unpack(ld4(SF0 + f * ldn + n0), s0);
  unpack(ld4(SPHI + f * ldn + n0), sp);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float a = a2[c];
    float m[4];
    unpack(ld4(SM + (f * NCH + c) * ldn + n0), m);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[c][e] = fmaxf(fmaf(a, s0[e] * sp[e], m[e]), SIMM_EPS);
  }
}

// ---- lead-side numerator / denominator planes ---------------------------------------------
// c_c = a2_c other / max(hat_c, eps);  num = sum_c c_c SX_c / max(hat_c, eps);  den = sum_c c_c
// with other = SPHI (HF0 update) or SF0 (HPHI / HGAMMA updates)
// (SIMM.py:304-305, :319-320, :352-353 with a2 = 1; Stereo: :622-640, :685-700, :776-790)
template <int NCH>
__global__ void __launch_bounds__(SE_THREADS)
simm_lead_terms_kernel(const float* __restrict__ SM, const float* __restrict__ SF0,
                       const float* __restrict__ SPHI, const float* __restrict__ SX,
                       const float* __restrict__ a2, int other_is_sf0, float* __restrict__ out,
                       long N, long ldn) {
  const long f = blockIdx.y;
  const long n0 = ((long)blockIdx.x * SE_THREADS + threadIdx.x) * 4;
  if (n0 >= ldn) return;
  float s0[4], sp[4], h[NCH][4], num[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
  load_model<NCH>(SM, SF0, SPHI, a2, f, n0, ldn, s0, sp, h);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float a = a2[c];
    float x[4];
    unpack(ld4(SX + (f * NCH + c) * ldn + n0), x);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ih = 1.0f / h[c][e];
      const float cc = a * (other_is_sf0 ? s0[e] : sp[e]) * ih;
      den[e] += cc;
      num[e] += cc * x[e] * ih;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (n0 + e >= N) num[e] = den[e] = 0.f;
  st4(out + f * 2 * ldn + n0, num);
  st4(out + f * 2 * ldn + ldn + n0, den);
}

// ---- accompaniment-side planes  T_c = SX_c / hat_c^2,  I_c = 1 / hat_c ------------------------
// mono (:335-337, :377-379): I = 1/max(hat,eps), T = I SX / max(hat,eps)
// stereo (:741-763, :829-843, :909-916): T = SX / max(hat^2, eps), I = 1 / max(hat, eps)
template <int NCH>
__global__ void __launch_bounds__(SE_THREADS)
simm_acc_terms_kernel(const float* __restrict__ SM, const float* __restrict__ SF0,
                      const float* __restrict__ SPHI, const float* __restrict__ SX,
                      const float* __restrict__ a2, float* __restrict__ out, int sq_clamp, long N,
                      long ldn) {
  const long f = blockIdx.y;
  const long n0 = ((long)blockIdx.x * SE_THREADS + threadIdx.x) * 4;
  if (n0 >= ldn) return;
  float s0[4], sp[4], h[NCH][4];
  load_model<NCH>(SM, SF0, SPHI, a2, f, n0, ldn, s0, sp, h);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float x[4], t[4], iv[4];
    unpack(ld4(SX + (f * NCH + c) * ldn + n0), x);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      iv[e] = 1.0f / h[c][e];
      t[e] = sq_clamp ? x[e] / fmaxf(h[c][e] * h[c][e], SIMM_EPS) : iv[e] * x[e] * iv[e];
      if (n0 + e >= N) t[e] = iv[e] = 0.f;
    }
    st4(out + (f * 2 * NCH + c) * ldn + n0, t);
    st4(out + (f * 2 * NCH + NCH + c) * ldn + n0, iv);
  }
}

// ---- model power  hat_c = max(a2_c SF0 SPHI + SM_c, eps) ------------------------------------------
// (SIMM.py:313, :329, :345, :370, :391; Stereo: :655-664 and after every update)
template <int NCH>
__global__ void __launch_bounds__(SE_THREADS)
simm_hat_kernel(const float* __restrict__ SM, const float* __restrict__ SF0,
                const float* __restrict__ SPHI, const float* __restrict__ a2,
                float* __restrict__ hat, long N, long ldn) {
  const long f = blockIdx.y;
  const long n0 = ((long)blockIdx.x * SE_THREADS + threadIdx.x) * 4;
  if (n0 >= ldn) return;
  float s0[4], sp[4], h[NCH][4];
  load_model<NCH>(SM, SF0, SPHI, a2, f, n0, ldn, s0, sp, h);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (n0 + e >= N) h[c][e] = 1.f;
    st4(hat + (f * NCH + c) * ldn + n0, h[c]);
  }
}

// ---- plane reductions (two stages, fixed order, float64) -------------------------------------------
// MODE 0: Itakura-Saito divergence sum_c sum (-log r + r - 1), r = SX_c / hat_c   (SIMM.py:35-44)
// MODE 1: the alpha update sums, per channel  sum d SX_c / max(hat_c, eps)  and  sum d,
//         d = SF0 SPHI / max(hat_c, eps)   (SIMM.py:869-884)
constexpr int SR_CTAS = 148 * 4;
template <int NCH, int MODE>
__global__ void __launch_bounds__(SE_THREADS)
simm_plane_reduce_kernel(const float* __restrict__ SX, const float* __restrict__ SM,
                         const float* __restrict__ SF0, const float* __restrict__ SPHI,
                         const float* __restrict__ a2, double* __restrict__ partial, int F, long N,
                         long ldn) {
  constexpr int NV = MODE == 0 ? 1 : 2 * NCH;
  __shared__ double s_red[SE_THREADS / 32];
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  const long per_row = ldn / 4;
  const long total = (long)F * per_row;
  for (long i = (long)blockIdx.x * SE_THREADS + threadIdx.x; i < total;
       i += (long)gridDim.x * SE_THREADS) {
    const long f = i / per_row, n0 = (i % per_row) * 4;
    float s0[4], sp[4], h[NCH][4];
    load_model<NCH>(SM, SF0, SPHI, a2, f, n0, ldn, s0, sp, h);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float x[4];
      unpack(ld4(SX + (f * NCH + c) * ldn + n0), x);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n0 + e >= N) continue;
        if (MODE == 0) {
          const float r = x[e] / h[c][e];
          a += r - 1.0f - logf(r);
        } else {
          const float ih = 1.0f / h[c][e];
          const float d = s0[e] * sp[e] * ih;
          a += d * x[e] * ih;
          b += d;
        }
      }
      if (MODE == 0) {
        acc[0] += (double)a;
      } else {
        acc[2 * c] += (double)a;
        acc[2 * c + 1] += (double)b;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double d = block_sum<SE_THREADS>(acc[i], s_red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * NV + i] = d;
  }
}

// out[slot] = sum of the IS partials
__global__ void simm_isdiv_finalize_kernel(const double* __restrict__ partial, int nparts,
                                           double* __restrict__ out) {
  __shared__ double s_red[SE_THREADS / 32];
  double a = 0.0;
  for (int i = threadIdx.x; i < nparts; i += SE_THREADS) a += partial[i];
  const double d = block_sum<SE_THREADS>(a, s_red);
  if (threadIdx.x == 0) out[0] = d;
}

// alpha update (SIMM.py:869-896): alpha_c <- max(alpha_c (num_c / den_c)^(omega/10), eps), then
// alphaR <- alphaR / max(alphaR + alphaL, .001), alphaL <- 1 - alphaR.  alpha: double[2];
// a2: float[2] (= alpha^2, what the plane kernels read).
__global__ void simm_alpha_finalize_kernel(const double* __restrict__ partial, int nparts,
                                           double omega, double* __restrict__ alpha,
                                           float* __restrict__ a2) {
  __shared__ double s_red[SE_THREADS / 32];
  __shared__ double s_sum[4];
  for (int v = 0; v < 4; ++v) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nparts; i += SE_THREADS) a += partial[(size_t)i * 4 + v];
    const double d = block_sum<SE_THREADS>(a, s_red);
    if (threadIdx.x == 0) s_sum[v] = d;
  }
  if (threadIdx.x == 0) {
    const double aR = fmax(alpha[0] * pow(s_sum[0] / s_sum[1], omega * 0.1), 1e-20);
    const double aL = fmax(alpha[1] * pow(s_sum[2] / s_sum[3], omega * 0.1), 1e-20);
    const double r = aR / fmax(aR + aL, 0.001);
    alpha[0] = r;
    alpha[1] = 1.0 - r;
    a2[0] = (float)(r * r);
    a2[1] = (float)((1.0 - r) * (1.0 - r));
  }
}

// ---- multiplicative update of an H matrix (rows x N) ---------------------------------------------
// num = sum_c w[c][r] C[r][c ldn + n],  den = sum_c w[c][r] C[r][(nch + c) ldn + n]  (w = 1 if null)
// theta <- theta (num / max(den, eps))^omega, optionally floored  (SIMM.py:308-310, :321-323,
// :338-343; Stereo :641-650, :701-708, :752-763).  C is a GEMM output [rows][2 nch ldn].
__global__ void simm_update_rows_kernel(float* __restrict__ theta, long ldt,
                                        const float* __restrict__ C, long ldc, int nch, long ldn,
                                        const float* __restrict__ w, int wld, float omega,
                                        float floor_value, int rows, long N) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= N || r >= rows) return;
  float num = 0.f, den = 0.f;
  for (int c = 0; c < nch; ++c) {
    const float wc = w ? w[(size_t)c * wld + r] : 1.0f;
    num += wc * C[(size_t)r * ldc + (size_t)c * ldn + n];
    den += wc * C[(size_t)r * ldc + (size_t)(nch + c) * ldn + n];
  }
  const float ratio = num / fmaxf(den, SIMM_EPS);
  const float g = omega == 1.0f ? ratio : powf(ratio, omega);
  float t = theta[(size_t)r * ldt + n] * g;
  if (floor_value > 0.f) t = fmaxf(t, floor_value);
  theta[(size_t)r * ldt + n] = t;
}

// ---- HPHI column normalisation (SIMM.py:324-326, :361-365; Stereo :710-716, :803-809) -----------
// HPHI[k][n] *= rowscale[k] (if given); s[n] = sum_k HPHI[k][n]; HPHI[:, n] /= s[n] where s > 0
__global__ void simm_hphi_norm_kernel(float* __restrict__ HPHI, long ldn, int K,
                                      const float* __restrict__ rowscale, long N,
                                      float* __restrict__ s_out) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    float v = HPHI[(size_t)k * ldn + n];
    if (rowscale) {
      v *= rowscale[k];
      HPHI[(size_t)k * ldn + n] = v;
    }
    s += v;
  }
  if (s > 0.f)
    for (int k = 0; k < K; ++k) HPHI[(size_t)k * ldn + n] /= s;
  s_out[n] = s;
}

// P[r][n] *= s[n]  (HF0 *= sumHPHI, and the same scaling of SF0 = WF0 HF0 instead of a new GEMM)
__global__ void __launch_bounds__(SE_THREADS)
simm_colscale_kernel(float* __restrict__ P, long ld, const float* __restrict__ s, long N) {
  const int r = blockIdx.y;
  const long n0 = ((long)blockIdx.x * SE_THREADS + threadIdx.x) * 4;
  if (n0 >= N) return;
  float v[4], sc[4];
  unpack(ld4(P + (long)r * ld + n0), v);
  unpack(ld4(s + n0), sc);
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = n0 + e < N ? v[e] * sc[e] : 0.f;
  st4(P + (long)r * ld + n0, v);
}

// P[r][n] *= s[r]
__global__ void simm_rowscale_kernel(float* __restrict__ P, long ld, const float* __restrict__ s,
                                     long N) {
  const int r = blockIdx.y;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  P[(size_t)r * ld + n] *= s[r];
}

// ---- HGAMMA update (SIMM.py:354-362; Stereo :791-802) ------------------------------------------------
// tn / td = (num | den) @ HPHI^T  [F][ldt] from the split-K GEMMs;
// HGAMMA[p][k] *= (WGAMMA^T tn / max(WGAMMA^T td, eps))^omega  -- one CTA per entry (p, k) --
// then (second kernel) the columns are normalised to sum one, s_out[k] = the column sums (they
// scale the rows of HPHI).
__global__ void __launch_bounds__(SE_THREADS)
simm_hgamma_update_kernel(float* __restrict__ HG, int ldhg, const float* __restrict__ WG, int ldwg,
                          const float* __restrict__ tn, const float* __restrict__ td, int ldt,
                          int F, int K, float omega) {
  __shared__ double s_red[SE_THREADS / 32];
  const int p = blockIdx.x / K, k = blockIdx.x % K;
  double num = 0.0, den = 0.0;
  for (int f = threadIdx.x; f < F; f += SE_THREADS) {
    const double wgv = (double)WG[(size_t)f * ldwg + p];
    num += wgv * (double)tn[(size_t)f * ldt + k];
    den += wgv * (double)td[(size_t)f * ldt + k];
  }
  num = block_sum<SE_THREADS>(num, s_red);
  den = block_sum<SE_THREADS>(den, s_red);
  if (threadIdx.x == 0) {
    const float ratio = (float)(num / fmax(den, 1e-20));
    HG[(size_t)p * ldhg + k] *= omega == 1.0f ? ratio : powf(ratio, omega);
  }
}

__global__ void simm_hgamma_norm_kernel(float* __restrict__ HG, int ldhg, int P, int K,
                                        float* __restrict__ s_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += HG[(size_t)p * ldhg + k];
  if (s > 0.f)
    for (int p = 0; p < P; ++p) HG[(size_t)p * ldhg + k] /= s;
  s_out[k] = s;
}

// ---- WM update (SIMM.py:376-388; Stereo :829-866), one CTA per accompaniment component -------------
// D: [2 nch][F][ldr], the products T_c HM^T (c < nch) and I_c HM^T.
// WM[f][r] *= (sum_c b2[c][r] D_c / sum_c b2[c][r] D_{nch+c})^omega, the denominator clamped to
// eps in the mono model only (Q12); column r normalised to sum one, s_out[r] = the sum.
__global__ void __launch_bounds__(SE_THREADS)
simm_wm_update_kernel(float* __restrict__ WM, int ldr, const float* __restrict__ D, int nch,
                      const float* __restrict__ b2, int clamp_den, float omega, int F,
                      float* __restrict__ s_out) {
  __shared__ double s_red[SE_THREADS / 32];
  __shared__ float s_sum;
  const int r = blockIdx.x;
  const size_t plane = (size_t)F * ldr;
  double acc = 0.0;
  for (int f = threadIdx.x; f < F; f += SE_THREADS) {
    float num = 0.f, den = 0.f;
    for (int c = 0; c < nch; ++c) {
      const float wc = b2 ? b2[(size_t)c * ldr + r] : 1.0f;
      num += wc * D[(size_t)c * plane + (size_t)f * ldr + r];
      den += wc * D[(size_t)(nch + c) * plane + (size_t)f * ldr + r];
    }
    if (clamp_den) den = fmaxf(den, SIMM_EPS);
    const float ratio = num / den;
    const float v = WM[(size_t)f * ldr + r] * (omega == 1.0f ? ratio : powf(ratio, omega));
    WM[(size_t)f * ldr + r] = v;
    acc += (double)v;
  }
  const double total = block_sum<SE_THREADS>(acc, s_red);
  if (threadIdx.x == 0) {
    s_sum = (float)total;
    s_out[r] = (float)total;
  }
  __syncthreads();
  const float s = s_sum;
  if (s > 0.f)
    for (int f = threadIdx.x; f < F; f += SE_THREADS) WM[(size_t)f * ldr + r] /= s;
}

// ---- beta update (SIMM.py:909-941), one CTA per accompaniment component ------------------------------
// dg_q[r] = sum_f WM[f][r] D_q[f][r] (the diagonal of WM^T (plane_q HM^T));
// betaR <- betaR (dg_TR / dg_IR)^(omega/10), betaL likewise, betaR <- betaR / max(betaR+betaL, eps),
// betaL <- 1 - betaR.  beta: double [2][ldr]; b2: float [2][ldr] = beta^2.
__global__ void __launch_bounds__(SE_THREADS)
simm_beta_update_kernel(const float* __restrict__ WM, int ldr, const float* __restrict__ D, int F,
                        double omega, double* __restrict__ beta, float* __restrict__ b2) {
  __shared__ double s_red[SE_THREADS / 32];
  __shared__ double s_dg[4];
  const int r = blockIdx.x;
  const size_t plane = (size_t)F * ldr;
  for (int q = 0; q < 4; ++q) {
    double acc = 0.0;
    for (int f = threadIdx.x; f < F; f += SE_THREADS)
      acc += (double)WM[(size_t)f * ldr + r] * (double)D[(size_t)q * plane + (size_t)f * ldr + r];
    const double d = block_sum<SE_THREADS>(acc, s_red);
    if (threadIdx.x == 0) s_dg[q] = d;
  }
  if (threadIdx.x == 0) {
    // D planes: 0 = T_R, 1 = T_L, 2 = I_R, 3 = I_L
    double bR = beta[r] * pow(s_dg[0] / s_dg[2], omega * 0.1);
    const double bL = beta[ldr + r] * pow(s_dg[1] / s_dg[3], omega * 0.1);
    bR = bR / fmax(bR + bL, 1e-20);
    beta[r] = bR;
    beta[ldr + r] = 1.0 - bR;
    b2[r] = (float)(bR * bR);
    b2[ldr + r] = (float)((1.0 - bR) * (1.0 - bR));
  }
}

// WMs[c][f][r] = WM[f][r] b2[c][r]  (the left factor of SM_c = (WM beta_c^2) HM), zero padding
__global__ void simm_wm_scaled_kernel(const float* __restrict__ WM, int ldr, int R,
                                      const float* __restrict__ b2, int nch, int F,
                                      float* __restrict__ WMs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * ldr) return;
  const int r = i % ldr;
  for (int c = 0; c < nch; ++c)
    WMs[(size_t)c * F * ldr + i] = r < R ? WM[i] * (b2 ? b2[(size_t)c * ldr + r] : 1.0f) : 0.f;
}

// ---- front / back end of the separation (SeparateLeadStereoTF.py:843-917, :1762-1871) ------------
// SX_c = |X_c|^2
__global__ void __launch_bounds__(SE_THREADS)
simm_power_kernel(const float* __restrict__ X, long ldx, float* __restrict__ SX, int nch, int F,
                  long N, long ldn) {
  const long f = blockIdx.y;
  const long n0 = ((long)blockIdx.x * SE_THREADS + threadIdx.x) * 4;
  if (n0 >= ldn) return;
  for (int c = 0; c < nch; ++c) {
    float p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      p[e] = 0.f;
      if (n0 + e < N) {
        const float re = X[((size_t)(2 * c) * F + f) * ldx + n0 + e];
        const float im = X[((size_t)(2 * c + 1) * F + f) * ldx + n0 + e];
        p[e] = re * re + im * im;
      }
    }
    st4(SX + (f * nch + c) * ldn + n0, p);
  }
}

// Y[c] = a2_c SF0 SPHI / hat_c X_c ; Y[nch + c] = SM_c / hat_c X_c
template <int NCH>
__global__ void __launch_bounds__(SE_THREADS)
simm_masks_kernel(const float* __restrict__ SM, const float* __restrict__ SF0,
                  const float* __restrict__ SPHI, const float* __restrict__ a2,
                  const float* __restrict__ X, long ldx, float* __restrict__ Y, float eps_hat,
                  int F, long N, long ldn) {
  const long f = blockIdx.y;
  const long n = (long)blockIdx.x * SE_THREADS + threadIdx.x;
  if (n >= N) return;
  const float lead = SF0[f * ldn + n] * SPHI[f * ldn + n];
  const size_t plane = (size_t)F * ldx;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float sm = SM[(f * NCH + c) * ldn + n];
    const float lv = a2[c] * lead;
    const float ih = 1.0f / fmaxf(lv + sm, eps_hat);
    const float re = X[(size_t)(2 * c) * plane + f * ldx + n];
    const float im = X[(size_t)(2 * c + 1) * plane + f * ldx + n];
    const float gl = lv * ih, gm = sm * ih;
    Y[(size_t)(2 * c) * plane + f * ldx + n] = gl * re;
    Y[(size_t)(2 * c + 1) * plane + f * ldx + n] = gl * im;
    Y[(size_t)(2 * (NCH + c)) * plane + f * ldx + n] = gm * re;
    Y[(size_t)(2 * (NCH + c) + 1) * plane + f * ldx + n] = gm * im;
  }
}

// ---- IS-NMF initialisers (pyfasst/tools/nmf.py:24-159; eps = 1e-10 there) ------------------------------
// out = (T | I):  T = SX / max(hat^2, eps),  I = 1 / max(hat, eps);  zero in the padding
__global__ void __launch_bounds__(SE_THREADS)
nmf_is_terms_kernel(const float* __restrict__ hat, const float* __restrict__ SX,
                    float* __restrict__ out, float eps, long N, long ldn) {
  const long f = blockIdx.y;
  const long n0 = ((long)blockIdx.x * SE_THREADS + threadIdx.x) * 4;
  if (n0 >= ldn) return;
  float h[4], x[4], t[4], iv[4];
  unpack(ld4(hat + f * ldn + n0), h);
  unpack(ld4(SX + f * ldn + n0), x);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    t[e] = x[e] / fmaxf(h[e] * h[e], eps);
    iv[e] = 1.0f / fmaxf(h[e], eps);
    if (n0 + e >= N) t[e] = iv[e] = 0.f;
  }
  st4(out + f * 2 * ldn + n0, t);
  st4(out + f * 2 * ldn + ldn + n0, iv);
}

// H[k][n] *= C[k][n] / max(C[k][ldn + n], eps)   (nmf.py:54-60, :151-157)
__global__ void nmf_update_rows_kernel(float* __restrict__ H, long ldh, const float* __restrict__ C,
                                       long ldc, long ldn, float eps, int rows, long N) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= N || r >= rows) return;
  H[(size_t)r * ldh + n] *= C[(size_t)r * ldc + n] / fmaxf(C[(size_t)r * ldc + ldn + n], eps);
}

// W[f][k] *= D[0][f][k] / max(D[1][f][k], eps); s = sum_f W[f][k] (0 -> 1); W[:, k] /= s;
// s_out[k] = s (it scales row k of H)   (nmf.py:40-51, :133-146).  One CTA per column.
__global__ void __launch_bounds__(SE_THREADS)
nmf_w_update_kernel(float* __restrict__ W, int ldk, const float* __restrict__ D, float eps, int F,
                    float* __restrict__ s_out) {
  __shared__ double s_red[SE_THREADS / 32];
  __shared__ float s_sum;
  const int k = blockIdx.x;
  const size_t plane = (size_t)F * ldk;
  double acc = 0.0;
  for (int f = threadIdx.x; f < F; f += SE_THREADS) {
    const float v = W[(size_t)f * ldk + k] *
                    (D[(size_t)f * ldk + k] / fmaxf(D[plane + (size_t)f * ldk + k], eps));
    W[(size_t)f * ldk + k] = v;
    acc += (double)v;
  }
  const double total = block_sum<SE_THREADS>(acc, s_red);
  if (threadIdx.x == 0) {
    s_sum = total == 0.0 ? 1.0f : (float)total;
    s_out[k] = s_sum;
  }
  __syncthreads();
  const float s = s_sum;
  for (int f = threadIdx.x; f < F; f += SE_THREADS) W[(size_t)f * ldk + k] /= s;
}

// out[f][n] = mean over channels of |X_c|^2: the one-channel power spectrum the NMF
// initialisers factorise (audioModel.py:2150-2158)
__global__ void __launch_bounds__(SE_THREADS)
mono_power_kernel(const float* __restrict__ X, long ldx, int nch, float* __restrict__ out, int F,
                  long N, long ldn) {
  const long f = blockIdx.y;
  const long n = (long)blockIdx.x * SE_THREADS + threadIdx.x;
  if (n >= ldn) return;
  float p = 0.f;
  if (n < N) {
    for (int c = 0; c < 2 * nch; ++c) {
      const float v = X[((size_t)c * F + f) * ldx + n];
      p += v * v;
    }
    p /= (float)nch;
  }
  out[f * ldn + n] = p;
}

}  // namespace pf

using namespace pf;

#define SIMM_PLANE_GRID(ldn, F) dim3(ceil_div((ldn) / 4, SE_THREADS), (F))

static int simm_check_plane(const char* who, int nch, int F, int64_t N, int64_t ldn) {
  if (!(nch == 1 || nch == 2)) {
    set_error("%s: nch=%d (1 = SIMM, 2 = Stereo_SIMM)", who, nch);
    return PF_ERR_ARG;
  }
  if (!(F > 0 && N > 0 && ldn >= N && ldn % 4 == 0)) {
    set_error("%s: F=%d N=%ld ldn=%ld (ldn >= N, ldn %% 4 == 0)", who, F, (long)N, (long)ldn);
    return PF_ERR_ARG;
  }
  return PF_OK;
}

extern "C" int pf_simm_lead_terms(const float* SM, const float* SF0, const float* SPHI,
                                  const float* SX, const float* a2, int other_is_sf0, float* out,
                                  int nch, int F, int64_t N, int64_t ldn, void* stream) {
  int rc = simm_check_plane("pf_simm_lead_terms", nch, F, N, ldn);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (nch == 1)
    simm_lead_terms_kernel<1><<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, st>>>(
        SM, SF0, SPHI, SX, a2, other_is_sf0, out, N, ldn);
  else
    simm_lead_terms_kernel<2><<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, st>>>(
        SM, SF0, SPHI, SX, a2, other_is_sf0, out, N, ldn);
  return check_launch("simm_lead_terms_kernel");
}

extern "C" int pf_simm_acc_terms(const float* SM, const float* SF0, const float* SPHI,
                                 const float* SX, const float* a2, float* out, int nch,
                                 int sq_clamp, int F, int64_t N, int64_t ldn, void* stream) {
  int rc = simm_check_plane("pf_simm_acc_terms", nch, F, N, ldn);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (nch == 1)
    simm_acc_terms_kernel<1><<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, st>>>(
        SM, SF0, SPHI, SX, a2, out, sq_clamp, N, ldn);
  else
    simm_acc_terms_kernel<2><<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, st>>>(
        SM, SF0, SPHI, SX, a2, out, sq_clamp, N, ldn);
  return check_launch("simm_acc_terms_kernel");
}

extern "C" int pf_simm_hat(const float* SM, const float* SF0, const float* SPHI, const float* a2,
                           float* hat, int nch, int F, int64_t N, int64_t ldn, void* stream) {
  int rc = simm_check_plane("pf_simm_hat", nch, F, N, ldn);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (nch == 1)
    simm_hat_kernel<1><<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, st>>>(SM, SF0, SPHI, a2, hat, N,
                                                                      ldn);
  else
    simm_hat_kernel<2><<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, st>>>(SM, SF0, SPHI, a2, hat, N,
                                                                      ldn);
  return check_launch("simm_hat_kernel");
}

extern "C" int64_t pf_simm_reduce_workspace_bytes(void) {
  return (int64_t)SR_CTAS * 4 * sizeof(double);
}

extern "C" int pf_simm_is_divergence(const float* SX, const float* SM, const float* SF0,
                                     const float* SPHI, const float* a2, int nch, int F, int64_t N,
                                     int64_t ldn, double* workspace, double* out, void* stream) {
  int rc = simm_check_plane("pf_simm_is_divergence", nch, F, N, ldn);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (nch == 1)
    simm_plane_reduce_kernel<1, 0><<<SR_CTAS, SE_THREADS, 0, st>>>(SX, SM, SF0, SPHI, a2,
                                                                  workspace, F, N, ldn);
  else
    simm_plane_reduce_kernel<2, 0><<<SR_CTAS, SE_THREADS, 0, st>>>(SX, SM, SF0, SPHI, a2,
                                                                  workspace, F, N, ldn);
  rc = check_launch("simm_plane_reduce_kernel");
  if (rc) return rc;
  simm_isdiv_finalize_kernel<<<1, SE_THREADS, 0, st>>>(workspace, SR_CTAS, out);
  return check_launch("simm_isdiv_finalize_kernel");
}

extern "C" int pf_simm_alpha_update(const float* SX, const float* SM, const float* SF0,
                                    const float* SPHI, int F, int64_t N, int64_t ldn, double omega,
                                    double* workspace, double* alpha, float* a2, void* stream) {
  int rc = simm_check_plane("pf_simm_alpha_update", 2, F, N, ldn);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  simm_plane_reduce_kernel<2, 1><<<SR_CTAS, SE_THREADS, 0, st>>>(SX, SM, SF0, SPHI, a2, workspace,
                                                                F, N, ldn);
  rc = check_launch("simm_plane_reduce_kernel");
  if (rc) return rc;
  simm_alpha_finalize_kernel<<<1, SE_THREADS, 0, st>>>(workspace, SR_CTAS, omega, alpha, a2);
  return check_launch("simm_alpha_finalize_kernel");
}

extern "C" int pf_simm_update_rows(float* theta, int64_t ldt, const float* C, int64_t ldc, int nch,
                                   int64_t ldn, const float* w, int wld, double omega,
                                   double floor_value, int rows, int64_t N, void* stream) {
  PF_REQUIRE(rows > 0 && N > 0 && nch >= 1 && ldc >= 2 * nch * ldn,
             "pf_simm_update_rows: rows=%d N=%ld nch=%d ldc=%ld ldn=%ld", rows, (long)N, nch,
             (long)ldc, (long)ldn);
  dim3 grid(ceil_div(N, 256), rows);
  simm_update_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      theta, ldt, C, ldc, nch, ldn, w, wld, (float)omega, (float)floor_value, rows, N);
  return check_launch("simm_update_rows_kernel");
}

extern "C" int pf_simm_hphi_normalise(float* HPHI, int64_t ldn, int K, const float* rowscale,
                                      int64_t N, float* s_out, void* stream) {
  PF_REQUIRE(K > 0 && N > 0 && ldn >= N, "pf_simm_hphi_normalise: K=%d N=%ld", K, (long)N);
  simm_hphi_norm_kernel<<<ceil_div(N, 256), 256, 0, as_stream(stream)>>>(HPHI, ldn, K, rowscale, N,
                                                                        s_out);
  return check_launch("simm_hphi_norm_kernel");
}

extern "C" int pf_simm_scale_columns(float* P, int64_t ld, int rows, int64_t N, const float* s,
                                     void* stream) {
  PF_REQUIRE(rows > 0 && N > 0 && ld % 4 == 0 && ld >= N, "pf_simm_scale_columns: rows=%d N=%ld",
             rows, (long)N);
  dim3 grid(ceil_div((N + 3) / 4, SE_THREADS), rows);
  simm_colscale_kernel<<<grid, SE_THREADS, 0, as_stream(stream)>>>(P, ld, s, N);
  return check_launch("simm_colscale_kernel");
}

extern "C" int pf_simm_scale_rows(float* P, int64_t ld, int rows, int64_t N, const float* s,
                                  void* stream) {
  PF_REQUIRE(rows > 0 && N > 0 && ld >= N, "pf_simm_scale_rows: rows=%d N=%ld", rows, (long)N);
  dim3 grid(ceil_div(N, 256), rows);
  simm_rowscale_kernel<<<grid, 256, 0, as_stream(stream)>>>(P, ld, s, N);
  return check_launch("simm_rowscale_kernel");
}

extern "C" int pf_simm_hgamma_update(float* HGAMMA, int ldhg, const float* WGAMMA, int ldwg,
                                     const float* tn, const float* td, int ldt, int F, int P,
                                     int K, double omega, float* s_out, void* stream) {
  PF_REQUIRE(F > 0 && P > 0 && K > 0, "pf_simm_hgamma_update: F=%d P=%d K=%d", F, P, K);
  simm_hgamma_update_kernel<<<P * K, SE_THREADS, 0, as_stream(stream)>>>(
      HGAMMA, ldhg, WGAMMA, ldwg, tn, td, ldt, F, K, (float)omega);
  int rc = check_launch("simm_hgamma_update_kernel");
  if (rc) return rc;
  simm_hgamma_norm_kernel<<<ceil_div(K, 64), 64, 0, as_stream(stream)>>>(HGAMMA, ldhg, P, K, s_out);
  return check_launch("simm_hgamma_norm_kernel");
}

extern "C" int pf_simm_wm_update(float* WM, int ldr, int R, const float* D, int nch,
                                 const float* b2, int clamp_den, double omega, int F, float* s_out,
                                 void* stream) {
  PF_REQUIRE(F > 0 && R > 0 && ldr >= R && (nch == 1 || nch == 2),
             "pf_simm_wm_update: F=%d R=%d ldr=%d nch=%d", F, R, ldr, nch);
  simm_wm_update_kernel<<<R, SE_THREADS, 0, as_stream(stream)>>>(WM, ldr, D, nch, b2, clamp_den,
                                                                (float)omega, F, s_out);
  return check_launch("simm_wm_update_kernel");
}

extern "C" int pf_simm_beta_update(const float* WM, int ldr, int R, const float* D, int F,
                                   double omega, double* beta, float* b2, void* stream) {
  PF_REQUIRE(F > 0 && R > 0 && ldr >= R, "pf_simm_beta_update: F=%d R=%d ldr=%d", F, R, ldr);
  simm_beta_update_kernel<<<R, SE_THREADS, 0, as_stream(stream)>>>(WM, ldr, D, F, omega, beta, b2);
  return check_launch("simm_beta_update_kernel");
}

extern "C" int pf_simm_power(const float* X, int64_t ldx, float* SX, int nch, int F, int64_t N,
                             int64_t ldn, void* stream) {
  int rc = simm_check_plane("pf_simm_power", nch, F, N, ldn);
  if (rc) return rc;
  PF_REQUIRE(ldx >= N, "pf_simm_power: ldx=%ld < N=%ld", (long)ldx, (long)N);
  simm_power_kernel<<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, as_stream(stream)>>>(X, ldx, SX, nch,
                                                                                  F, N, ldn);
  return check_launch("simm_power_kernel");
}

extern "C" int pf_simm_masks(const float* SM, const float* SF0, const float* SPHI, const float* a2,
                             const float* X, int64_t ldx, float* Y, double eps_hat, int nch, int F,
                             int64_t N, int64_t ldn, void* stream) {
  int rc = simm_check_plane("pf_simm_masks", nch, F, N, ldn);
  if (rc) return rc;
  PF_REQUIRE(ldx >= N, "pf_simm_masks: ldx=%ld < N=%ld", (long)ldx, (long)N);
  dim3 grid(ceil_div(N, SE_THREADS), F);
  cudaStream_t st = as_stream(stream);
  if (nch == 1)
    simm_masks_kernel<1><<<grid, SE_THREADS, 0, st>>>(SM, SF0, SPHI, a2, X, ldx, Y, (float)eps_hat,
                                                     F, N, ldn);
  else
    simm_masks_kernel<2><<<grid, SE_THREADS, 0, st>>>(SM, SF0, SPHI, a2, X, ldx, Y, (float)eps_hat,
                                                     F, N, ldn);
  return check_launch("simm_masks_kernel");
}

extern "C" int pf_nmf_is_terms(const float* hat, const float* SX, float* out, double eps, int F,
                               int64_t N, int64_t ldn, void* stream) {
  int rc = simm_check_plane("pf_nmf_is_terms", 1, F, N, ldn);
  if (rc) return rc;
  nmf_is_terms_kernel<<<SIMM_PLANE_GRID(ldn, F), SE_THREADS, 0, as_stream(stream)>>>(
      hat, SX, out, (float)eps, N, ldn);
  return check_launch("nmf_is_terms_kernel");
}

extern "C" int pf_nmf_update_rows(float* H, int64_t ldh, const float* C, int64_t ldc, int64_t ldn,
                                  double eps, int rows, int64_t N, void* stream) {
  PF_REQUIRE(rows > 0 && N > 0 && ldc >= 2 * ldn, "pf_nmf_update_rows: rows=%d N=%ld ldc=%ld", rows,
             (long)N, (long)ldc);
  dim3 grid(ceil_div(N, 256), rows);
  nmf_update_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(H, ldh, C, ldc, ldn, (float)eps, rows,
                                                            N);
  return check_launch("nmf_update_rows_kernel");
}

extern "C" int pf_nmf_w_update(float* W, int ldk, int K, const float* D, double eps, int F,
                               float* s_out, void* stream) {
  PF_REQUIRE(F > 0 && K > 0 && ldk >= K, "pf_nmf_w_update: F=%d K=%d ldk=%d", F, K, ldk);
  nmf_w_update_kernel<<<K, SE_THREADS, 0, as_stream(stream)>>>(W, ldk, D, (float)eps, F, s_out);
  return check_launch("nmf_w_update_kernel");
}

extern "C" int pf_mono_power(const float* X, int64_t ldx, int nch, float* out, int F, int64_t N,
                             int64_t ldn, void* stream) {
  PF_REQUIRE(nch >= 1 && F > 0 && N > 0 && ldn >= N && ldx >= N, "pf_mono_power: nch=%d F=%d N=%ld",
             nch, F, (long)N);
  dim3 grid(ceil_div(ldn, SE_THREADS), F);
  mono_power_kernel<<<grid, SE_THREADS, 0, as_stream(stream)>>>(X, ldx, nch, out, F, N, ldn);
  return check_launch("mono_power_kernel");
}

extern "C" int pf_simm_wm_scaled(const float* WM, int ldr, int R, const float* b2, int nch, int F,
                                 float* WMs, void* stream) {
  PF_REQUIRE(F > 0 && R > 0 && ldr >= R && (nch == 1 || nch == 2),
             "pf_simm_wm_scaled: F=%d R=%d ldr=%d nch=%d", F, R, ldr, nch);
  simm_wm_scaled_kernel<<<ceil_div((long)F * ldr, 256), 256, 0, as_stream(stream)>>>(
      WM, ldr, R, b2, nch, F, WMs);
  return check_launch("simm_wm_scaled_kernel");
}
