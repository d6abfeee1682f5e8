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
#include <stdlib.h>

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


// ---- FB update, fast path: P and O are the same plane (one spectral component per
// spatial component, single factor -- what MultiChanNMFInst_FASST / MultiChanNMFConv build).
// Then O/P = 1 exactly, so den[f,k] = sum_n G[k,n] for every f (a row sum of G, computed by
// g_rowsum_kernel) and only num[f,k] = sum_n (hatW/P)[f,n] G[k,n] has to be contracted.
// A CTA owns FB_ROWS rows and a run of frames; the hatW / P / G tiles of a step (128 frames)
// are brought to shared memory by cp.async through a 3-stage ring; a warp owns FR rows, a
// lane 4 (2) frames of the tile, and accumulates FR x KC partial sums that are reduced across
// the lanes once at the end.
constexpr int FBF_THREADS = 256;
constexpr int FBF_FR = 4;
constexpr int FBF_ROWS = (FBF_THREADS / 32) * FBF_FR;  // 32 rows per CTA
constexpr int FBF_STAGES = 3;

template <typename T, int KC>
struct FbfStage {
  static constexpr int NT = 32 * VecOf<T>::N;
  T g[KC][NT];
  T hw[FBF_ROWS][NT];
  T p[FBF_ROWS][NT];
};

template <typename T, int KC>
__global__ void __launch_bounds__(FBF_THREADS, 1)
fb_contract_same_kernel(const T* __restrict__ hatW, const T* __restrict__ Pp, long ld,
                        const T* __restrict__ G, long ldg, int k0, int K, int F, long N,
                        long chunk, int nsplit, double* __restrict__ num) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int NT = 32 * VEC;
  constexpr int VPR = NT / VEC;  // 16-byte vectors per tile row (= 32)
  constexpr T kEps = (T)1e-10;
  extern __shared__ __align__(16) unsigned char fbf_smem[];
  FbfStage<T, KC>* stages = reinterpret_cast<FbfStage<T, KC>*>(fbf_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fblk = blockIdx.y * FBF_ROWS;
  const int split = blockIdx.x;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;
  const int nsteps = (int)((end - begin + NT - 1) / NT);

  auto issue = [&](int step) {
    FbfStage<T, KC>& st = stages[step % FBF_STAGES];
    const long nb = begin + (long)step * NT;
    for (int i = threadIdx.x; i < KC * VPR; i += FBF_THREADS) {
      const int k = i / VPR, c = i % VPR;
      const long n = nb + (long)c * VEC;
      const bool ok = (k0 + k < K) && (n + VEC <= ldg);
      cp_async16(&st.g[k][c * VEC], ok ? (const void*)(G + (long)(k0 + k) * ldg + n) : (const void*)G,
                 ok ? 16 : 0);
    }
    for (int i = threadIdx.x; i < FBF_ROWS * VPR; i += FBF_THREADS) {
      const int r = i / VPR, c = i % VPR;
      const long n = nb + (long)c * VEC;
      const int f = fblk + r;
      const bool ok = (f < F) && (n + VEC <= ld);
      const long off = ok ? (long)f * ld + n : 0;
      cp_async16(&st.hw[r][c * VEC], hatW + off, ok ? 16 : 0);
      cp_async16(&st.p[r][c * VEC], Pp + off, ok ? 16 : 0);
    }
  };

  T an[FBF_FR][KC];
#pragma unroll
  for (int r = 0; r < FBF_FR; ++r)
#pragma unroll
    for (int k = 0; k < KC; ++k) an[r][k] = (T)0;

#pragma unroll
  for (int s = 0; s < FBF_STAGES - 1; ++s) {
    if (s < nsteps) issue(s);
    cp_async_commit();
  }
  for (int s = 0; s < nsteps; ++s) {
    if (s + FBF_STAGES - 1 < nsteps) issue(s + FBF_STAGES - 1);
    cp_async_commit();
    cp_async_wait<FBF_STAGES - 1>();  // stage s has landed
    __syncthreads();
    const FbfStage<T, KC>& st = stages[s % FBF_STAGES];
    T e1[FBF_FR][VEC];
#pragma unroll
    for (int r = 0; r < FBF_FR; ++r) {
      T hw[VEC], p[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        hw[e] = st.hw[warp * FBF_FR + r][lane * VEC + e];
        p[e] = st.p[warp * FBF_FR + r][lane * VEC + e];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        e1[r][e] = hw[e] * pf_rcp(pf_max(p[e], kEps));  // hatW / P^2 * O with O == P
    }
    // k in groups of 4: the group's G vectors are loaded, then 64 FMAs; the empty asm keeps
    // the compiler from hoisting all KC shared-memory loads (and their registers) up front
#pragma unroll
    for (int kg = 0; kg < KC; kg += 4) {
      T g[4][VEC];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int e = 0; e < VEC; ++e) g[kk][e] = st.g[kg + kk][lane * VEC + e];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int r = 0; r < FBF_FR; ++r)
#pragma unroll
          for (int e = 0; e < VEC; ++e) an[r][kg + kk] += e1[r][e] * g[kk][e];
      asm volatile("" ::: "memory");
    }
    __syncthreads();  // the ring slot of stage s is refilled by the next issue()
  }
  if (sizeof(T) == 4) {
    // lanes hold partial sums over their frames.  1280 shuffles per thread would cost as much
    // as ~10 steps of the main loop, so the FR*KC values are transposed through the (now idle)
    // ring instead: lane l then adds the 32 partials of values l, l+32, ... in double.
    constexpr int NV = FBF_FR * KC;
    T* red = reinterpret_cast<T*>(fbf_smem) + (size_t)warp * NV * 33;
#pragma unroll
    for (int r = 0; r < FBF_FR; ++r)
#pragma unroll
      for (int k = 0; k < KC; ++k) red[(r * KC + k) * 33 + lane] = an[r][k];
    __syncwarp();
    for (int v = lane; v < NV; v += 32) {
      double d = 0.0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) d += (double)red[v * 33 + l];
      const int r = v / KC, k = v % KC;
      const int f = fblk + warp * FBF_FR + r;
      if (f < F && k0 + k < K) num[((size_t)split * F + f) * K + k0 + k] = d;
    }
  } else {
#pragma unroll
    for (int r = 0; r < FBF_FR; ++r) {
      const int f = fblk + warp * FBF_FR + r;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const double sn = warp_sum((double)an[r][k]);
        if (lane == 0 && f < F && k0 + k < K) num[((size_t)split * F + f) * K + k0 + k] = sn;
      }
    }
  }
}

// den[split][f][k] = sum over the split's frames of G[k][n], broadcast over f
template <typename T>
__global__ void g_rowsum_kernel(const T* __restrict__ G, long ldg, int K, int F, long N,
                                long chunk, double* __restrict__ den) {
  const int split = blockIdx.x, k = blockIdx.y;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;
  double acc = 0.0;
  for (long n = begin + threadIdx.x; n < end; n += blockDim.x) acc += (double)G[(long)k * ldg + n];
  __shared__ double s_red[32];
  __shared__ double s_tot;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) d += s_red[w];
    s_tot = d;
  }
  __syncthreads();
  const double tot = s_tot;
  for (int f = threadIdx.x; f < F; f += blockDim.x) den[((size_t)split * F + f) * K + k] = tot;
}

// ============================ TW update: contract over frequencies ==============
constexpr int TW_THREADS = 128;  // one frame per thread
constexpr int TW_RT = 16;        // rows of a pipeline stage
constexpr int TW_STAGES = 3;

template <typename T, int KC>
struct TwStage {
  T hw[TW_RT][TW_THREADS];
  T o[TW_RT][TW_THREADS];
  T w[TW_RT][KC];
};

// A CTA owns 128 frames and the frequency rows [fb, fe); the hatW / O / W tiles of 16 rows
// come in through a 3-stage cp.async ring.  Thread n keeps H[:, n] and the 2 x KC partial sums
// of frame n in registers and walks the rows: P' = max(W H, eps) (own power with the updated
// W, audioModel.py:1639-1645), then num += W e1, den += W e2.
template <typename T, int KC>
__global__ void __launch_bounds__(TW_THREADS)
tw_contract_kernel(const T* __restrict__ hatW, const T* __restrict__ Op, long ld,
                   const T* __restrict__ W, int ldw, const T* __restrict__ H, long ldh, int K,
                   int F, long N, int fchunk, int fsplit, double* __restrict__ num,
                   double* __restrict__ den, long ldo) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int VPR = TW_THREADS / VEC;  // 16-byte vectors per tile row
  constexpr T kEps = (T)1e-10;
  extern __shared__ __align__(16) unsigned char tw_smem[];
  TwStage<T, KC>* stages = reinterpret_cast<TwStage<T, KC>*>(tw_smem);
  const long nb = (long)blockIdx.x * TW_THREADS;
  const long n = nb + threadIdx.x;
  const int split = blockIdx.y;
  const int fb = split * fchunk;
  int fe = fb + fchunk;
  if (fe > F) fe = F;
  const int nsteps = (fe - fb + TW_RT - 1) / TW_RT;
  const bool live = n < N;

  auto issue = [&](int step) {
    TwStage<T, KC>& st = stages[step % TW_STAGES];
    const int f0 = fb + step * TW_RT;
    for (int i = threadIdx.x; i < TW_RT * VPR; i += TW_THREADS) {
      const int r = i / VPR, c = i % VPR;
      const long nn = nb + (long)c * VEC;
      const int f = f0 + r;
      const bool ok = (f < fe) && (nn + VEC <= ld);
      const long off = ok ? (long)f * ld + nn : 0;
      cp_async16(&st.hw[r][c * VEC], hatW + off, ok ? 16 : 0);
      cp_async16(&st.o[r][c * VEC], Op + off, ok ? 16 : 0);
    }
    for (int i = threadIdx.x; i < TW_RT * KC; i += TW_THREADS) {
      const int r = i / KC, k = i % KC;
      const int f = f0 + r;
      const bool ok = (f < fe) && (k < K);
      cp_async_small<sizeof(T)>(&st.w[r][k], W + (ok ? (long)f * ldw + k : 0),
                                ok ? (int)sizeof(T) : 0);
    }
  };

  T h[KC], an[KC], ad[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    h[k] = (live && k < K) ? H[(long)k * ldh + n] : (T)0;
    an[k] = ad[k] = (T)0;
  }
#pragma unroll
  for (int s = 0; s < TW_STAGES - 1; ++s) {
    if (s < nsteps) issue(s);
    cp_async_commit();
  }
  for (int s = 0; s < nsteps; ++s) {
    if (s + TW_STAGES - 1 < nsteps) issue(s + TW_STAGES - 1);
    cp_async_commit();
    cp_async_wait<TW_STAGES - 1>();
    __syncthreads();
    const TwStage<T, KC>& st = stages[s % TW_STAGES];
    const int rows = min(TW_RT, fe - (fb + s * TW_RT));
    for (int r = 0; r < rows; ++r) {
      const T hw = st.hw[r][threadIdx.x];
      const T o = pf_max(st.o[r][threadIdx.x], kEps);
      T w[KC];
      T p = (T)0;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        w[k] = st.w[r][k];
        p += w[k] * h[k];
      }
      p = pf_max(p, kEps);
      const T rp = pf_rcp(p);
      const T e2 = o * rp;               // other / P'            (:1694-1701)
      const T e1 = o * (hw * rp * rp);   // other * hat_W / P'^2  (:1714-1720)
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        an[k] += w[k] * e1;
        ad[k] += w[k] * e2;
      }
    }
    __syncthreads();
  }
  if (n >= ldo) return;
#pragma unroll
  for (int k = 0; k < KC; ++k)
    if (k < K) {
      num[((size_t)split * K + k) * ldo + n] = live ? (double)an[k] : 0.0;
      den[((size_t)split * K + k) * ldo + n] = live ? (double)ad[k] : 0.0;
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

// Frequency sharding: the split partial sums of the TW numerators / denominators, reduced in a
// fixed order, converted to the plane type and laid out chunk-major for the reduce-scatter over
// the frames:  out[w][q][k][i] = sum_s part_q[s][k][w c + i]  (q = 0 num, 1 den; c = ld / world;
// k < K of Kmax rows).  One launch instead of two sum_splits and a strided copy.
template <typename T>
__global__ void tw_pack_chunks_kernel(const double* __restrict__ num, const double* __restrict__ den,
                                      int nsplit, long split_stride, int K, long ld, long c,
                                      int Kmax, T* __restrict__ out) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (n >= ld) return;
  double sn = 0.0, sd = 0.0;
  for (int s = 0; s < nsplit; ++s) {
    sn += num[(size_t)s * split_stride + (size_t)k * ld + n];
    sd += den[(size_t)s * split_stride + (size_t)k * ld + n];
  }
  const long w = n / c, i = n - w * c;
  T* o = out + ((size_t)w * 2 * Kmax + k) * c + i;
  o[0] = (T)sn;
  o[(size_t)Kmax * c] = (T)sd;
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

// theta *= (sum_s num[s] / max(sum_s den[s], eps))^omega with the split partial sums reduced
// on the fly in a fixed order (one launch instead of two sum_splits + mult_update)
template <typename T>
__global__ void mult_update_splits_kernel(T* __restrict__ theta, long ldt,
                                          const double* __restrict__ num,
                                          const double* __restrict__ den, int nsplit,
                                          long split_stride, long ldnd, int rows, long cols,
                                          double omega) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= cols || r >= rows) return;
  double sn = 0.0, sd = 0.0;
  for (int s = 0; s < nsplit; ++s) {
    sn += num[(size_t)s * split_stride + (size_t)r * ldnd + c];
    sd += den[(size_t)s * split_stride + (size_t)r * ldnd + c];
  }
  const double ratio = sn / fmax(sd, 1e-10);
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
  const size_t smem = sizeof(TwStage<T, KC>) * TW_STAGES;
  cudaError_t e = cudaFuncSetAttribute(tw_contract_kernel<T, KC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("tw_contract_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  tw_contract_kernel<T, KC><<<grid, TW_THREADS, smem, st>>>(
      (const T*)hatW, (const T*)O, ld, (const T*)W, ldw, (const T*)H, ldh, K, F, N, fchunk,
      fsplit, num, den, ldo);
  return check_launch("tw_contract_kernel");
}

template <typename T, int KC>
static int launch_fb_same(const void* hatW, const void* P, long ld, const void* G, long ldg,
                          int k0, int K, int F, long N, long chunk, int nsplit, double* num,
                          cudaStream_t st) {
  dim3 grid(nsplit, ceil_div(F, FBF_ROWS));
  const size_t smem = sizeof(FbfStage<T, KC>) * FBF_STAGES;
  cudaError_t e = cudaFuncSetAttribute(fb_contract_same_kernel<T, KC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("fb_contract_same_kernel: %zu bytes of shared memory: %s", smem,
              cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  fb_contract_same_kernel<T, KC><<<grid, FBF_THREADS, smem, st>>>(
      (const T*)hatW, (const T*)P, ld, (const T*)G, ldg, k0, K, F, N, chunk, nsplit, num);
  return check_launch("fb_contract_same_kernel");
}

template <typename T>
static int dispatch_fb_same(const void* hatW, const void* P, long ld, const void* G, long ldg,
                            int K, int F, long N, long chunk, int nsplit, double* num, double* den,
                            cudaStream_t st) {
  int rc = PF_OK;
  for (int k0 = 0; k0 < K && rc == PF_OK; k0 += 32) {
    const int kc = K - k0;
    if (kc <= 4)
      rc = launch_fb_same<T, 4>(hatW, P, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, st);
    else if (kc <= 8)
      rc = launch_fb_same<T, 8>(hatW, P, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, st);
    else if (kc <= 16)
      rc = launch_fb_same<T, 16>(hatW, P, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, st);
    else
      rc = launch_fb_same<T, 32>(hatW, P, ld, G, ldg, k0, K, F, N, chunk, nsplit, num, st);
  }
  if (rc) return rc;
  dim3 grid(nsplit, K);
  g_rowsum_kernel<T><<<grid, 256, 0, st>>>((const T*)G, ldg, K, F, N, chunk, den);
  return check_launch("g_rowsum_kernel");
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

// tcgen05 versions (nmf_tc.cu), float32 planes only
int pf_fb_contract_tc(const float* hatW, const float* P, long ld, const float* G, long ldg, int K,
                      int F, long N, long chunk, int nsplit, double* num, cudaStream_t st);
int pf_spec_power_tc(const float* W, int ldw, const float* H, long ldh, float* V, long ldv, int F,
                     int K, long N, cudaStream_t st);
int pf_tw_contract_fused_tc(const float* hatW, const float* O, long ld, const float* W, int ldw,
                            const float* H, long ldh, int K, int F, long N, int fchunk, int fsplit,
                            double* num, double* den, long ldo, cudaStream_t st);
constexpr bool TW_FUSED_DEFAULT = true;
int pf_tw_contract_tc(const float* hatW, const float* O, const float* Pn, long ld, const float* W,
                      int ldw, int K, int F, long N, int fchunk, int fsplit, double* num,
                      double* den, long ldo, cudaStream_t st);
static bool use_tensor_cores() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PYFASST_NO_TENSOR_CORES");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static bool tw_fused() {
  const char* e = getenv("PYFASST_TW_FUSED");
  return e != nullptr ? e[0] == '1' : TW_FUSED_DEFAULT;
}

extern "C" int pf_spec_power(const void* W, int ldw, const void* H, int64_t ldh, void* V,
                             int64_t ldv, int F, int K, int64_t N, int accumulate, int dtype,
                             void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_spec_power: bad dtype %d", dtype);
  PF_REQUIRE(F > 0 && K > 0 && N > 0, "pf_spec_power: empty problem");
  PF_REQUIRE(ldv % 4 == 0 && ldh % 4 == 0 && ldv >= N, "pf_spec_power: ldv/ldh must be multiples of 4");
  cudaStream_t st = as_stream(stream);
  const int FT = (SP_THREADS / 32) * SP_FR;
  // large float32 products go to the tensor cores (3xTF32); the CUDA-core kernel serves
  // float64, accumulation into V, K > 32 and the small factor products
  if (dtype == PF_F32 && !accumulate && K <= 32 && F >= 64 && N >= 1024 && use_tensor_cores())
    return pf_spec_power_tc((const float*)W, ldw, (const float*)H, ldh, (float*)V, ldv, F, K, N, st);
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
  long steps = (N + nt - 1) / nt;
  if (dtype == PF_F32 && use_tensor_cores()) {
    // tcgen05 kernel: CTAs of 128 rows, two resident per SM -> ~2 x 148 CTAs of equal length
    const long fblocks = (F + 127) / 128;
    long ns = (2 * 148L + fblocks - 1) / fblocks;
    if (ns > steps) ns = steps;
    if (ns < 1) ns = 1;
    const long per = (steps + ns - 1) / ns;
    *chunk = per * nt;
    *nsplit = (int)((N + *chunk - 1) / *chunk);
    return PF_OK;
  }
  // CUDA-core kernel: ~32 steps of one tile (128 / 64 frames) per CTA: many CTAs of equal,
  // moderate length balance over the 148 SMs without a tail
  long per = 32;
  const long fblocks = (F + FBF_ROWS - 1) / FBF_ROWS;
  while (per > 4 && fblocks * ((steps + per - 1) / per) < 148L * 2) per /= 2;
  *chunk = per * nt;
  *nsplit = (int)((N + *chunk - 1) / *chunk);
  (void)K;
  return PF_OK;
}

extern "C" int pf_nmf_fb_contract(const void* hatW, const void* P, const void* O, int64_t ld,
                                  const void* G, int64_t ldg, int F, int K, int64_t N,
                                  double* num_partial, double* den_partial, int64_t chunk,
                                  int nsplit, int dtype, void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_nmf_fb_contract: bad dtype %d", dtype);
  PF_REQUIRE(ld % 4 == 0 && ldg % 4 == 0, "pf_nmf_fb_contract: ld/ldg must be multiples of 4");
  PF_REQUIRE(chunk > 0 && (int64_t)nsplit * chunk >= N, "pf_nmf_fb_contract: bad split plan");
  const long nt = 32 * (dtype == PF_F64 ? 2 : 4);
  PF_REQUIRE(chunk % nt == 0, "pf_nmf_fb_contract: chunk must be a multiple of %ld frames", nt);
  cudaStream_t st = as_stream(stream);
  if (P == O) {  // one plane: O/P == 1, see fb_contract_same_kernel
    if (dtype == PF_F32 && use_tensor_cores()) {
      int rc = pf_fb_contract_tc((const float*)hatW, (const float*)P, ld, (const float*)G, ldg, K,
                                 F, N, chunk, nsplit, num_partial, st);
      if (rc) return rc;
      dim3 grid(nsplit, K);
      g_rowsum_kernel<float><<<grid, 256, 0, st>>>((const float*)G, ldg, K, F, N, chunk,
                                                   den_partial);
      return check_launch("g_rowsum_kernel");
    }
    if (dtype == PF_F32)
      return dispatch_fb_same<float>(hatW, P, ld, G, ldg, K, F, N, chunk, nsplit, num_partial,
                                     den_partial, st);
    return dispatch_fb_same<double>(hatW, P, ld, G, ldg, K, F, N, chunk, nsplit, num_partial,
                                    den_partial, st);
  }
  if (dtype == PF_F32)
    return dispatch_fb<float>(hatW, P, O, ld, G, ldg, K, F, N, chunk, nsplit, num_partial,
                              den_partial, st);
  return dispatch_fb<double>(hatW, P, O, ld, G, ldg, K, F, N, chunk, nsplit, num_partial,
                             den_partial, st);
}

extern "C" int pf_nmf_tw_plan(int F, int K, int64_t N, int dtype, int* fchunk, int* fsplit) {
  const long nblocks = (N + TW_THREADS - 1) / TW_THREADS;  // both kernels: 128 frames per CTA
  if (dtype == PF_F32 && K <= 32 && use_tensor_cores()) {
    // tcgen05 kernel: two CTAs per SM; pick the frequency split (1..4) that fills whole
    // waves of 2 x 148 CTAs best (every split costs a K x N partial-sum plane)
    long best = 1;
    double best_eff = 0.0;
    for (long fs = 1; fs <= 4; ++fs) {
      if (fs > 1 && (F + fs - 1) / fs < 4 * 16) break;
      const long ctas = nblocks * fs, slots = 2 * 148;
      const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
      if (eff > best_eff + 0.02) { best_eff = eff; best = fs; }
    }
    int fc = (int)((F + best - 1) / best);
    fc = ((fc + 15) / 16) * 16;
    *fchunk = fc;
    *fsplit = (F + fc - 1) / fc;
    return PF_OK;
  }
  // CUDA-core kernel: split the frequency axis until ~3 CTAs per SM are in flight, keeping
  // at least two pipeline stages of rows per CTA
  long fs = (148L * 3 + nblocks / 2) / nblocks;
  const long max_fs = (F + 2 * TW_RT - 1) / (2 * TW_RT);
  if (fs > max_fs) fs = max_fs;
  if (fs > 64) fs = 64;
  if (fs < 1) fs = 1;
  int fc = (int)((F + fs - 1) / fs);
  fc = ((fc + TW_RT - 1) / TW_RT) * TW_RT;
  *fchunk = fc;
  *fsplit = (F + fc - 1) / fc;
  return PF_OK;
}

extern "C" int pf_nmf_tw_contract(const void* hatW, const void* O, int64_t ld, const void* W,
                                  int ldw, const void* H, int64_t ldh, int F, int K, int64_t N,
                                  double* num_partial, double* den_partial, int64_t ldo,
                                  int fchunk, int fsplit, void* scratch_plane, int dtype,
                                  void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_nmf_tw_contract: bad dtype %d", dtype);
  PF_REQUIRE(fchunk > 0 && (int64_t)fsplit * fchunk >= F, "pf_nmf_tw_contract: bad split plan");
  cudaStream_t st = as_stream(stream);
  if (dtype == PF_F32 && K <= 32 && scratch_plane != nullptr && F >= 64 && N >= 1024 &&
      fchunk % 16 == 0 && use_tensor_cores()) {
    // tensor-core paths.  Fused (default; PYFASST_TW_FUSED=0 switches it off): P' = W' H is
    // formed in the kernel by a first MMA; otherwise P' goes through the scratch plane (two
    // kernels).  Measured on configs[1]: 2.456 against 2.518 ms per GEM iteration.
    if (tw_fused())
      return pf_tw_contract_fused_tc((const float*)hatW, (const float*)O, ld, (const float*)W, ldw,
                                     (const float*)H, ldh, K, F, N, fchunk, fsplit, num_partial,
                                     den_partial, ldo, st);
    int rc = pf_spec_power_tc((const float*)W, ldw, (const float*)H, ldh, (float*)scratch_plane,
                              ld, F, K, N, st);
    if (rc) return rc;
    return pf_tw_contract_tc((const float*)hatW, (const float*)O, (const float*)scratch_plane, ld,
                             (const float*)W, ldw, K, F, N, fchunk, fsplit, num_partial,
                             den_partial, ldo, st);
  }
  if (dtype == PF_F32)
    return dispatch_tw<float>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num_partial,
                              den_partial, ldo, st);
  return dispatch_tw<double>(hatW, O, ld, W, ldw, H, ldh, K, F, N, fchunk, fsplit, num_partial,
                             den_partial, ldo, st);
}

extern "C" int pf_tw_pack_chunks(const double* num_partial, const double* den_partial, int nsplit,
                                 int64_t split_stride, int K, int64_t ld, int world, int Kmax,
                                 void* out, int dtype, void* stream) {
  PF_REQUIRE(nsplit >= 1 && K >= 1 && K <= Kmax && world >= 1 && ld > 0 && ld % world == 0,
             "pf_tw_pack_chunks: nsplit=%d K=%d Kmax=%d ld=%ld world=%d", nsplit, K, Kmax, (long)ld,
             world);
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_tw_pack_chunks: bad dtype %d", dtype);
  dim3 grid(ceil_div(ld, 256), K);
  if (dtype == PF_F32)
    tw_pack_chunks_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        num_partial, den_partial, nsplit, split_stride, K, ld, ld / world, Kmax, (float*)out);
  else
    tw_pack_chunks_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
        num_partial, den_partial, nsplit, split_stride, K, ld, ld / world, Kmax, (double*)out);
  return check_launch("tw_pack_chunks_kernel");
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

extern "C" int pf_mult_update_splits(void* theta, int64_t ldt, const double* num_partial,
                                     const double* den_partial, int nsplit, int64_t split_stride,
                                     int64_t ldnd, int rows, int64_t cols, double omega, int dtype,
                                     void* stream) {
  PF_REQUIRE(dtype == PF_F32 || dtype == PF_F64, "pf_mult_update_splits: bad dtype %d", dtype);
  PF_REQUIRE(rows > 0 && cols > 0 && nsplit >= 1, "pf_mult_update_splits: empty");
  dim3 grid(ceil_div(cols, 128), rows);
  if (dtype == PF_F32)
    mult_update_splits_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(
        (float*)theta, ldt, num_partial, den_partial, nsplit, split_stride, ldnd, rows, cols, omega);
  else
    mult_update_splits_kernel<double><<<grid, 128, 0, as_stream(stream)>>>(
        (double*)theta, ldt, num_partial, den_partial, nsplit, split_stride, ldnd, rows, cols,
        omega);
  return check_launch("mult_update_splits_kernel");
}
