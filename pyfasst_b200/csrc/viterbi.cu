// Viterbi decoding of the melody line, sm_100a.
//
// Replaces pyfasst/SeparateLeadStereo/tracking/_tracking.pyx: viterbiTracking (:11-93), the one
// native module of the reference (Cython), called between the two SIMM estimation stages
// (SeparateLeadStereoTF.py:1150-1230).  Same recursion in float64, same tie breaking (the
// reference scans the predecessors upwards with a strict `>`: the smallest index among equal
// maxima wins; np.argmax for the last frame likewise), so the decoded path is bit-identical.
//
// One CTA walks the frames (they are inherently sequential); a destination state is served by a
// group of G lanes that split its S predecessors, then combine (value, index) with shuffles.
// cum[s] lives in shared memory (ping-pong), the antecedents go to global memory [N][S].
//
// Main path (S <= 148 * 8 states): a cooperative grid.  Every CTA owns up to 8 destination
// states and keeps ITS columns of the transition matrix in shared memory for the whole run, so
// a frame costs one pass over S x D candidates per CTA, an exchange of the S new cumulative
// scores through global memory and one grid barrier (~2.5 us per frame instead of streaming the
// S x S matrix through one SM: 52 us at S = 480).  Fallback for more states: one CTA.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pf {

constexpr int VT_THREADS = 1024;
constexpr int VC_THREADS = 256;
constexpr int VC_DMAX = 8;

// best of two candidates: larger value, then smaller index
__device__ __forceinline__ void vt_best(double& v, int& i, double v2, int i2) {
  if (v2 > v || (v2 == v && i2 < i)) {
    v = v2;
    i = i2;
  }
}

// dens: [N][S] (frame major), trans: [S][S] (trans[s_][s]: from s_ to s), ante: [N][S]
__global__ void __launch_bounds__(VT_THREADS)
viterbi_forward_kernel(const double* __restrict__ dens, const double* __restrict__ prior,
                       const double* __restrict__ trans, int S, long N, int G,
                       int* __restrict__ ante, double* __restrict__ last_cum) {
  extern __shared__ double vt_cum[];  // [2][S]
  const int g = threadIdx.x % G;             // lane of the group
  const int slot = threadIdx.x / G;          // group index within the CTA
  const int groups = VT_THREADS / G;
  for (int s = threadIdx.x; s < S; s += VT_THREADS) {
    vt_cum[s] = prior[s] + dens[s];
    ante[s] = -1;
  }
  __syncthreads();
  for (long n = 1; n < N; ++n) {
    const double* prev = vt_cum + ((n - 1) & 1) * S;
    double* cur = vt_cum + (n & 1) * S;
    for (int s0 = 0; s0 < S; s0 += groups) {   // uniform trip count: shuffles stay convergent
      const int s = s0 + slot;
      double best = -INFINITY;
      int arg = 0x7fffffff;
      if (s < S) {
        // predecessor g first, then g + G, ...: increasing order with a strict `>`
        best = prev[g < S ? g : 0] + trans[(size_t)(g < S ? g : 0) * S + s];
        arg = g < S ? g : 0x7fffffff;
        if (g >= S) best = -INFINITY;
        for (int p = g + G; p < S; p += G) {
          const double v = prev[p] + trans[(size_t)p * S + s];
          if (v > best) {
            best = v;
            arg = p;
          }
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, arg, o);
        vt_best(best, arg, v2, i2);
      }
      if (s < S && g == 0) {
        cur[s] = best + dens[(size_t)n * S + s];
        ante[(size_t)n * S + s] = arg;
      }
    }
    __syncthreads();
  }
  const double* fin = vt_cum + ((N - 1) & 1) * S;
  for (int s = threadIdx.x; s < S; s += VT_THREADS) last_cum[s] = fin[s];
}

// cooperative version: CTA b owns destinations [b D, b D + D)
__global__ void __launch_bounds__(VC_THREADS)
viterbi_coop_kernel(const double* __restrict__ dens, const double* __restrict__ prior,
                    const double* __restrict__ trans, int S, long N, int D,
                    double* __restrict__ cum_g, int* __restrict__ ante,
                    double* __restrict__ last_cum) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double vc_smem[];
  double* ts = vc_smem;                 // [S][VC_DMAX] this CTA's columns of trans
  double* cum_s = ts + (size_t)S * VC_DMAX;  // [S]
  __shared__ double s_val[VC_THREADS / 32][VC_DMAX];
  __shared__ int s_arg[VC_THREADS / 32][VC_DMAX];
  const int d0 = blockIdx.x * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < S * VC_DMAX; i += VC_THREADS) {
    const int p = i / VC_DMAX, d = i % VC_DMAX;
    ts[i] = (d < D && d0 + d < S) ? trans[(size_t)p * S + d0 + d] : -INFINITY;
  }
  if (threadIdx.x < D && d0 + threadIdx.x < S) {
    const int s = d0 + threadIdx.x;
    cum_g[s] = prior[s] + dens[s];
    ante[s] = -1;
  }
  grid.sync();
  for (long n = 1; n < N; ++n) {
    const double* prev = cum_g + ((n - 1) & 1) * (size_t)S;
    for (int p = threadIdx.x; p < S; p += VC_THREADS) cum_s[p] = __ldcg(prev + p);
    __syncthreads();
    double best[VC_DMAX];
    int arg[VC_DMAX];
#pragma unroll
    for (int d = 0; d < VC_DMAX; ++d) {
      best[d] = -INFINITY;
      arg[d] = 0x7fffffff;
    }
    for (int p = threadIdx.x; p < S; p += VC_THREADS) {  // increasing p, strict `>`
      const double c = cum_s[p];
      const double* t = ts + (size_t)p * VC_DMAX;
#pragma unroll
      for (int d = 0; d < VC_DMAX; ++d) {
        const double v = c + t[d];
        if (v > best[d] || arg[d] == 0x7fffffff) {
          best[d] = v;
          arg[d] = p;
        }
      }
    }
#pragma unroll
    for (int d = 0; d < VC_DMAX; ++d) {
      for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, best[d], o);
        const int i2 = __shfl_xor_sync(0xffffffffu, arg[d], o);
        vt_best(best[d], arg[d], v2, i2);
      }
      if (lane == 0) {
        s_val[warp][d] = best[d];
        s_arg[warp][d] = arg[d];
      }
    }
    __syncthreads();
    if (threadIdx.x < D && d0 + threadIdx.x < S) {
      const int d = threadIdx.x, s = d0 + d;
      double b = s_val[0][d];
      int a = s_arg[0][d];
      for (int w = 1; w < VC_THREADS / 32; ++w) vt_best(b, a, s_val[w][d], s_arg[w][d]);
      cum_g[(n & 1) * (size_t)S + s] = b + dens[(size_t)n * S + s];
      ante[(size_t)n * S + s] = a;
    }
    grid.sync();
  }
  if (blockIdx.x == 0) {
    const double* fin = cum_g + ((N - 1) & 1) * (size_t)S;
    for (int s = threadIdx.x; s < S; s += VC_THREADS) last_cum[s] = __ldcg(fin + s);
  }
}

__global__ void viterbi_backtrack_kernel(const int* __restrict__ ante,
                                         const double* __restrict__ last_cum, int S, long N,
                                         long long* __restrict__ path) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int best = 0;
  for (int s = 1; s < S; ++s)
    if (last_cum[s] > last_cum[best]) best = s;  // np.argmax: the first maximum
  path[N - 1] = best;
  for (long n = N - 2; n >= 0; --n) {
    best = ante[(size_t)(n + 1) * S + best];
    path[n] = best;
  }
}

}  // namespace pf

using namespace pf;

extern "C" int64_t pf_viterbi_workspace_bytes(int S, int64_t N) {
  return (int64_t)N * S * sizeof(int) + (int64_t)3 * S * sizeof(double);
}

extern "C" int pf_viterbi(const double* log_density_ns, const double* log_prior,
                          const double* log_trans, int S, int64_t N, void* workspace,
                          int64_t workspace_bytes, long long* path, void* stream) {
  PF_REQUIRE(S >= 1 && S <= 6000 && N >= 1, "pf_viterbi: S=%d N=%ld (S <= 6000)", S, (long)N);
  PF_REQUIRE(workspace_bytes >= pf_viterbi_workspace_bytes(S, N),
             "pf_viterbi: workspace %ld < %ld bytes", (long)workspace_bytes,
             (long)pf_viterbi_workspace_bytes(S, N));
  cudaStream_t st = as_stream(stream);
  double* last_cum = (double*)workspace;
  double* cum_g = last_cum + S;  // [2][S] ping-pong of the cooperative kernel
  int* ante = (int*)(cum_g + 2 * (size_t)S);
  // cooperative grid: D destination states per CTA, at most one CTA per SM
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  int D = sms > 0 ? (S + sms - 1) / sms : VC_DMAX + 1;
  if (coop && D <= VC_DMAX && N > 1 && getenv("PYFASST_VITERBI_SINGLE_CTA") == nullptr) {
    const int ctas = (S + D - 1) / D;
    const size_t smem = ((size_t)S * VC_DMAX + S) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(viterbi_coop_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      long Nl = N;
      void* args[] = {(void*)&log_density_ns, (void*)&log_prior, (void*)&log_trans, (void*)&S,
                      (void*)&Nl, (void*)&D, (void*)&cum_g, (void*)&ante, (void*)&last_cum};
      e = cudaLaunchCooperativeKernel((void*)viterbi_coop_kernel, dim3(ctas), dim3(VC_THREADS),
                                      args, smem, st);
    }
    if (e != cudaSuccess) {
      set_error("viterbi_coop_kernel: %s", cudaGetErrorString(e));
      return PF_ERR_CUDA;
    }
    int rc = check_launch("viterbi_coop_kernel");
    if (rc) return rc;
    viterbi_backtrack_kernel<<<1, 32, 0, st>>>(ante, last_cum, S, N, path);
    return check_launch("viterbi_backtrack_kernel");
  }
  int G = 1;  // lanes per destination state: as many as keep all states in one sweep
  while (G < 32 && (long)S * (G * 2) <= VT_THREADS) G *= 2;
  const size_t smem = 2 * (size_t)S * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(viterbi_forward_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("viterbi_forward_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  viterbi_forward_kernel<<<1, VT_THREADS, smem, st>>>(log_density_ns, log_prior, log_trans, S, N, G,
                                                     ante, last_cum);
  int rc = check_launch("viterbi_forward_kernel");
  if (rc) return rc;
  viterbi_backtrack_kernel<<<1, 32, 0, st>>>(ante, last_cum, S, N, path);
  return check_launch("viterbi_backtrack_kernel");
}
