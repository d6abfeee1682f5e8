// K2, warp-specialised variant of the fused stereo E-step (float32 planes, J <= 4), sm_100a.
//
// Same per-bin algebra and outputs as estep_stereo_kernel (estep.cu; reference:
// FASST.compute_suff_stat, pyfasst/audioModel.py:580-764).  The fused kernel holds 77 moment
// accumulators next to the float64 per-bin algebra: 255 registers, 8 warps per SM, and half of
// the issue slots idle behind fixed-latency FP64 dependencies
// (profiles/r01/ncu_estep_stereo_kernel.txt).  Here the two halves run in different warps: a
// TEAM is ALG algebra warps (Sigma, y, M, hatW -- no accumulators) and one MOMENT warp (the
// packed float32 accumulators only).  An algebra lane leaves a 64-byte record per bin
// (v_j | M | x | y) in shared memory; the moment warp consumes the records of the previous pass
// while the algebra warps work on the next one (two stages, one named barrier per pass and
// team).  Both roles fit in 128-168 registers => 12-16 warps per SM.
#include "estep.cuh"

namespace pf {

constexpr int WS_SLOTS = 4;  // float4 per record
#ifndef MUA
#define MUA 3
#endif
#ifndef MUE
#define MUE 2
#endif
constexpr int MOM_UNROLL_A = MUA, MOM_UNROLL_E = MUE;

// ALG algebra warps + 1 moment warp per team, TEAMS teams per CTA, MINB CTAs per SM,
// VEC frames per algebra lane and pass (4: float4 accesses, 2: float2 -- half the staging
// registers).
template <int VEC_, int ALG_, int TEAMS_, int MINB_>
struct WsCfg {
  static constexpr int VEC = VEC_, ALG = ALG_, TEAMS = TEAMS_, MINB = MINB_;
  static constexpr int TEAM_THREADS = (ALG + 1) * 32;
  static constexpr int THREADS = TEAMS * TEAM_THREADS;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int LANES = 32 * VEC;             // frames per algebra warp and pass
  static constexpr int PASS = TEAMS * ALG * LANES;    // frames per CTA and pass
  static constexpr int STAGE = VEC * WS_SLOTS * 32;   // float4 per (algebra warp, stage)
  static constexpr size_t SMEM = (size_t)TEAMS * ALG * 2 * STAGE * sizeof(float4);
};
typedef WsCfg<2, 3, 2, 2> WsCfg0;  // 16 warps / SM at 128 registers: 12 algebra + 4 moment
typedef WsCfg<2, 2, 5, 1> WsCfg1;  // 15 warps / SM at 136 registers: 10 algebra + 5 moment
typedef WsCfg<4, 2, 2, 2> WsCfg2;  // 12 warps / SM at 168 registers:  8 algebra + 4 moment
typedef WsCfg<2, 2, 2, 2> WsCfg3;  // as 2 with float2 accesses
constexpr int WS_DEFAULT_CFG = 3;

template <int VEC>
__device__ __forceinline__ void load_frames(const float* p, float (&out)[VEC]) {
  if (VEC == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    out[0] = v.x; out[1] = v.y; out[VEC - 2] = v.z; out[VEC - 1] = v.w;
  } else {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    out[0] = v.x; out[1] = v.y;
  }
}
template <int VEC>
__device__ __forceinline__ void store_frames(float* p, const float (&in)[VEC]) {
  if (VEC == 4)
    *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[VEC - 2], in[VEC - 1]);
  else
    *reinterpret_cast<float2*>(p) = make_float2(in[0], in[1]);
}

template <int COUNT>
__device__ __forceinline__ void team_barrier(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(COUNT) : "memory");
}

// The warp of team t that accumulates the moments.  Warps are dealt to the four schedulers of an
// SM round-robin (warp id mod 4); the moment warps -- pure FFMA2 streams -- are spread over all
// four instead of landing on one.
template <typename Cfg>
__device__ __forceinline__ int moment_warp_of_team(int t, int cta_parity) {
  if (Cfg::WARPS == 15) {  // 5 teams of 3
    // warps 0, 5, 10, 3, 14: schedulers 0, 1, 2, 3, 2
    return t == 0 ? 0 : t == 1 ? 5 : t == 2 ? 10 : t == 3 ? 3 : 14;
  }
  if (Cfg::WARPS == 8) {  // 2 teams of 4: schedulers {0, 1} or {2, 3} by CTA parity
    return t * 4 + ((t + 2 * cta_parity) & 3);
  }
  // 2 teams of 3 (6 warps): warps {0, 5} or {2, 3}: schedulers {0, 1} / {2, 3}
  return cta_parity ? (t == 0 ? 2 : 3) : (t == 0 ? 0 : 5);
}

template <int J, typename Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
estep_stereo_ws_kernel(const float* __restrict__ X, const float* __restrict__ V,
                       const double* __restrict__ coef, const double* __restrict__ noise,
                       SubMap map, float* __restrict__ hatW, double* __restrict__ partial, int F,
                       long N, long ld, long chunk, int nsplit) {
  static_assert(J <= 4, "one float4 record slot holds the source powers");
  constexpr int NP = npairs(J);
  constexpr int NA = nacc(J);
  constexpr int NC = ncoef(J);
  constexpr float kLogPi = 1.1447298858494002f;
  constexpr int VEC = Cfg::VEC, ALG = Cfg::ALG, TEAMS = Cfg::TEAMS;
  constexpr int PASS = Cfg::PASS, STAGE = Cfg::STAGE, LANES = Cfg::LANES;

  const int f = blockIdx.y;
  const int split = blockIdx.x;
  extern __shared__ __align__(16) unsigned char s_ws_raw[];
  float4* s_rec = reinterpret_cast<float4*>(s_ws_raw);
  __shared__ double s_coef[NC];      // R_j (4 per source), then the mixed discriminants
  __shared__ double s_q[4 * J];      // R_j with the off-diagonal entries doubled: tr(M R_j)
  __shared__ float s_dcoef[NP];
  __shared__ double s_red[TEAMS][NA];
  __shared__ double s_ll[Cfg::WARPS];
  if (threadIdx.x < NC) {
    const double c = coef[(size_t)f * NC + threadIdx.x];
    s_coef[threadIdx.x] = c;
    if (threadIdx.x >= 4 * J)
      s_dcoef[threadIdx.x - 4 * J] = (float)c;
    else
      s_q[threadIdx.x] = (threadIdx.x & 2) ? 2.0 * c : c;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // roles: is this warp the moment warp of a team; otherwise its rank among the algebra warps
  int team = -1, alg = 0;
  {
    const int par = blockIdx.x & 1;
    int below = 0;  // moment warps with a smaller id
#pragma unroll
    for (int t = 0; t < TEAMS; ++t) {
      const int m = moment_warp_of_team<Cfg>(t, par);
      if (m == warp) team = t;
      below += m < warp ? 1 : 0;
    }
    if (team < 0) {
      const int idx = warp - below;  // 0 .. TEAMS*ALG-1
      team = idx / ALG;
      alg = idx - team * ALG;
      team = -team - 1;  // (negative: algebra warp of team -team-1)
    }
  }
  const bool is_moment = team >= 0;
  if (!is_moment) team = -team - 1;

  const long plane = (long)F * ld;
  const long row = (long)f * ld;
  const long begin = (long)split * chunk;
  long end = begin + chunk;
  if (end > N) end = N;
  const int npass = (int)((end - begin + PASS - 1) / PASS);
  float4* const team_rec = s_rec + (size_t)team * ALG * 2 * STAGE;

  if (is_moment) {
    Moments<float, J, true> mom;
    mom.clear();
    for (int i = 1; i <= npass; ++i) {
      team_barrier<Cfg::TEAM_THREADS>(team);  // the records of pass i-1 are complete
      const float4* rs = team_rec + (size_t)((i - 1) & 1) * STAGE + lane;
      const long nt0 = begin + (long)(i - 1) * PASS + (long)team * ALG * LANES;
#pragma unroll MOM_UNROLL_A
      for (int a = 0; a < ALG; ++a) {
        if (nt0 + a * LANES >= end) break;  // (warp-uniform) that warp had no frames in this pass
        const float4* r = rs + (size_t)a * 2 * STAGE;
#pragma unroll MOM_UNROLL_E
        for (int e = 0; e < VEC; ++e) {
          const float4 qv = r[(e * WS_SLOTS + 0) * 32];
          const float4 qt = r[(e * WS_SLOTS + 1) * 32];
          const float4 qx = r[(e * WS_SLOTS + 2) * 32];
          const float4 qz = r[(e * WS_SLOTS + 3) * 32];
          const float vv[4] = {qv.x, qv.y, qv.z, qv.w};
          float vt[J], pr[NP];
#pragma unroll
          for (int j = 0; j < J; ++j) vt[j] = vv[j];
          int p = 0;
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = j; k < J; ++k) pr[p++] = vt[j] * vt[k];
          const float u[8] = {qx.x * qz.x + qx.y * qz.y, qx.y * qz.x - qx.x * qz.y,
                              qx.x * qz.z + qx.y * qz.w, qx.y * qz.z - qx.x * qz.w,
                              qx.z * qz.x + qx.w * qz.y, qx.w * qz.x - qx.z * qz.y,
                              qx.z * qz.z + qx.w * qz.w, qx.w * qz.z - qx.z * qz.w};
          mom.add(pr, vt, qt.x, qt.y, qt.z, qt.w, u);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NA - 1; ++i) {
      const float d = warp_sum(mom.get(i));
      if (lane == 0) s_red[team][i] = (double)d;
    }
    if (lane == 0) s_ll[warp] = 0.0;
  } else {
    const double s2 = noise[f];
    float invrank[J];
#pragma unroll
    for (int j = 0; j < J; ++j) invrank[j] = (float)map.invrank[j];
    double acc_ll = 0.0;
    const long lane0 = begin + (long)(team * ALG + alg) * LANES + (long)lane * VEC;
    float nx[4][VEC], nv[J][VEC];
    auto issue_loads = [&](long n) {
      load_frames<VEC>(X + 0 * plane + row + n, nx[0]);
      load_frames<VEC>(X + 1 * plane + row + n, nx[1]);
      load_frames<VEC>(X + 2 * plane + row + n, nx[2]);
      load_frames<VEC>(X + 3 * plane + row + n, nx[3]);
#pragma unroll
      for (int j = 0; j < J; ++j) load_frames<VEC>(V + j * plane + row + n, nv[j]);
    };
    if (lane0 < end) issue_loads(lane0);
    for (int i = 0; i < npass; ++i) {
      const long n0 = lane0 + (long)i * PASS;
      if (n0 - (long)lane * VEC < end) {  // warp-uniform: the warp has frames in this pass
        float x0r[VEC], x0i[VEC], x1r[VEC], x1i[VEC], v[J][VEC], w[J][VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          x0r[e] = nx[0][e]; x0i[e] = nx[1][e]; x1r[e] = nx[2][e]; x1i[e] = nx[3][e];
#pragma unroll
          for (int j = 0; j < J; ++j) v[j][e] = nv[j][e];
        }
        // frames beyond the end of the row: zero inputs contribute nothing to the moments and
        // give hatW = 0; only the log-likelihood term is masked below
        if (n0 + VEC > end) {
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (n0 + e >= end) {
              x0r[e] = x0i[e] = x1r[e] = x1i[e] = 0.f;
#pragma unroll
              for (int j = 0; j < J; ++j) v[j][e] = 0.f;
            }
        }
        if (n0 + PASS < end) issue_loads(n0 + PASS);
        float4* r = team_rec + (size_t)(alg * 2 + (i & 1)) * STAGE + lane;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const bool live = n0 + e < end;
          double vj[J], i00, i11, i01r, i01i;
          float vt[J], pr[NP], det;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            vt[j] = v[j][e];
            vj[j] = (double)vt[j];
          }
          sigma_inverse<double, float, J>(vj, vt, s_coef, s_dcoef, s2, pr, det, i00, i11, i01r,
                                          i01i);
          const double a0r = (double)x0r[e], a0i = (double)x0i[e];
          const double a1r = (double)x1r[e], a1i = (double)x1i[e];
          const double y0r = i00 * a0r + i01r * a1r - i01i * a1i;
          const double y0i = i00 * a0i + i01r * a1i + i01i * a1r;
          const double y1r = i01r * a0r + i01i * a0i + i11 * a1r;
          const double y1i = i01r * a0i - i01i * a0r + i11 * a1i;
          const float z0r = (float)y0r, z0i = (float)y0i, z1r = (float)y1r, z1i = (float)y1i;
          const float quad = x0r[e] * z0r + x0i[e] * z0i + x1r[e] * z1r + x1i[e] * z1i;
          acc_ll += (double)(live ? __logf(det) + kLogPi + quad : 0.f);
          const double m00 = y0r * y0r + y0i * y0i - i00;
          const double m11 = y1r * y1r + y1i * y1i - i11;
          const double m01r = y0r * y1r + y0i * y1i - i01r;
          const double m01i = y0i * y1r - y0r * y1i - i01i;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const float q = (float)(s_q[4 * j + 0] * m00 + s_q[4 * j + 1] * m11 +
                                    (s_q[4 * j + 2] * m01r + s_q[4 * j + 3] * m01i));
            w[j][e] = fabsf(vt[j] + vt[j] * vt[j] * (q * invrank[j]));
          }
          r[(e * WS_SLOTS + 0) * 32] = make_float4(vt[0], J > 1 ? vt[J > 1 ? 1 : 0] : 0.f,
                                                   J > 2 ? vt[J > 2 ? 2 : 0] : 0.f,
                                                   J > 3 ? vt[J > 3 ? 3 : 0] : 0.f);
          r[(e * WS_SLOTS + 1) * 32] = make_float4((float)m00, (float)m11, (float)m01r, (float)m01i);
          r[(e * WS_SLOTS + 2) * 32] = make_float4(x0r[e], x0i[e], x1r[e], x1i[e]);
          r[(e * WS_SLOTS + 3) * 32] = make_float4(z0r, z0i, z1r, z1i);
        }
        if (n0 < end) {
#pragma unroll
          for (int j = 0; j < J; ++j) store_frames<VEC>(hatW + j * plane + row + n0, w[j]);
        }
      }
      team_barrier<Cfg::TEAM_THREADS>(team);
    }
    const double d = warp_sum(acc_ll);
    if (lane == 0) s_ll[warp] = d;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NA; i += Cfg::THREADS) {
    double d = 0.0;
    if (i < NA - 1) {
#pragma unroll
      for (int t = 0; t < TEAMS; ++t) d += s_red[t][i];
    } else {
#pragma unroll
      for (int w2 = 0; w2 < Cfg::WARPS; ++w2) d += s_ll[w2];
    }
    partial[((size_t)f * nsplit + split) * NA + i] = d;
  }
}

// ---- host side ------------------------------------------------------------------
// PYFASST_ESTEP_WSCFG = 0..3 selects the team layout (see the WsCfg typedefs).
static int estep_ws_cfg() {
  const char* e = getenv("PYFASST_ESTEP_WSCFG");
  const int v = e != nullptr ? atoi(e) : WS_DEFAULT_CFG;
  return v < 0 || v > 3 ? WS_DEFAULT_CFG : v;
}

long estep_ws_pass() {
  switch (estep_ws_cfg()) {
    case 0: return WsCfg0::PASS;
    case 1: return WsCfg1::PASS;
    case 2: return WsCfg2::PASS;
    default: return WsCfg3::PASS;
  }
}

// ~12k frames per CTA: the pipeline fill and the end-of-CTA reduction are amortised
long estep_ws_passes_per_cta() {
  long p = 1;
  while (2 * p * estep_ws_pass() <= 12288) p *= 2;
  return p;
}

// Which float32 kernel runs: the fused kernel, unless PYFASST_ESTEP_KERNEL=ws asks for this one
// (J <= 4).  Measured on configs[1] (B200, profiles/r01/estep_ws_experiment.txt): the fused
// kernel 0.765 ms; this one 0.85 ms at best (layout 0 without spills: 16 warps per SM, issue
// slots 57 % used instead of 49 %, but 315 instead of 245 instructions per bin -- the record
// traffic, the second computation of the pair products and the loop overhead of two roles cost
// more than the extra warps hide), 1.0-1.4 ms for the layouts whose algebra role spills.  Kept as
// an opt-in experiment, exercised by tests/test_kernels_gpu.py.
bool estep_use_ws(int J, long N, int dtype) {
  (void)N;
  if (dtype != PF_F32 || J > 4) return false;
  const char* e = getenv("PYFASST_ESTEP_KERNEL");
  return e != nullptr && e[0] == 'w';
}

template <int J, typename Cfg>
static int launch_estep_ws(const void* X, const void* V, const double* coef, const double* noise,
                           const SubMap& map, void* hatW, double* partial, int F, long N, long ld,
                           long chunk, int nsplit, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(estep_stereo_ws_kernel<J, Cfg>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) {
    set_error("estep_stereo_ws_kernel: %zu bytes of shared memory: %s", Cfg::SMEM,
              cudaGetErrorString(e));
    return PF_ERR_CUDA;
  }
  if (chunk % Cfg::PASS != 0) {
    set_error("estep_stereo_ws_kernel: chunk %ld is not a multiple of the pass %d", chunk,
              Cfg::PASS);
    return PF_ERR_ARG;
  }
  dim3 grid(nsplit, F);
  estep_stereo_ws_kernel<J, Cfg><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(
      (const float*)X, (const float*)V, coef, noise, map, (float*)hatW, partial, F, N, ld, chunk,
      nsplit);
  return check_launch("estep_stereo_ws_kernel");
}

template <int J>
static int launch_estep_ws_cfg(const void* X, const void* V, const double* coef,
                               const double* noise, const SubMap& map, void* hatW,
                               double* partial, int F, long N, long ld, long chunk, int nsplit,
                               cudaStream_t st) {
#define PF_WS_ARGS X, V, coef, noise, map, hatW, partial, F, N, ld, chunk, nsplit, st
  switch (estep_ws_cfg()) {
    case 0: return launch_estep_ws<J, WsCfg0>(PF_WS_ARGS);
    case 1: return launch_estep_ws<J, WsCfg1>(PF_WS_ARGS);
    case 2: return launch_estep_ws<J, WsCfg2>(PF_WS_ARGS);
    default: return launch_estep_ws<J, WsCfg3>(PF_WS_ARGS);
  }
}

int dispatch_estep_ws(int J, const void* X, const void* V, const double* coef,
                      const double* noise, const SubMap& map, void* hatW, double* partial, int F,
                      long N, long ld, long chunk, int nsplit, cudaStream_t st) {
  switch (J) {
    case 1: return launch_estep_ws_cfg<1>(PF_WS_ARGS);
    case 2: return launch_estep_ws_cfg<2>(PF_WS_ARGS);
    case 3: return launch_estep_ws_cfg<3>(PF_WS_ARGS);
    case 4: return launch_estep_ws_cfg<4>(PF_WS_ARGS);
  }
#undef PF_WS_ARGS
  set_error("estep_stereo_ws_kernel: J=%d not supported (1..4)", J);
  return PF_ERR_UNSUPPORTED;
}

}  // namespace pf
