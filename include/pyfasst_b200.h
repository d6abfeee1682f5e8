/* pyfasst_b200 -- C ABI of the B200 (sm_100a) kernels behind pyfasst's FASST hot path.
 *
 * The reference (s-ben/pyfasst) is pure Python/NumPy and has no FFI layer; its
 * operator boundary is the set of FASST methods listed below.  Each entry point
 * here replaces the numerical body of one of them (reference file:line given
 * per function, relative to /root/reference/pyfasst/).  The Python package
 * `pyfasst_b200` keeps the reference's class / method names and binds these
 * symbols with ctypes (pyfasst_b200/_lib.py); INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types cross the ABI.
 *  - every pointer is a DEVICE pointer unless the parameter is documented "host".
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all work is enqueued on it, nothing synchronises with the host.
 *  - return value: PF_OK (0) or a negative PF_ERR_*; pf_last_error() gives the
 *    message (thread local).  No exception crosses the ABI.
 *  - `dtype` selects the storage/arithmetic type of the F x N planes and of the
 *    NMF factors: PF_F32 (fast path) or PF_F64 (the reference's precision).
 *    Mixing matrices, sufficient statistics, noise PSD and all reductions are
 *    always float64 / complex128 (interleaved re,im doubles).
 *  - planes are row-major [rows][ld] with frames contiguous; ld % 4 == 0 and
 *    plane base pointers 16-byte aligned; padding frames (n >= N) are never read
 *    as data and are kept at zero by the producers.
 */
#ifndef PYFASST_B200_H_
#define PYFASST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 10

#define PF_OK 0
#define PF_ERR_ARG (-1)         /* invalid argument (ValueError on the Python side) */
#define PF_ERR_CUDA (-2)        /* CUDA runtime / launch error */
#define PF_ERR_UNSUPPORTED (-3) /* NotImplementedError on the Python side */

#define PF_F32 0
#define PF_F64 1

/* bits of the device-side status word written by the solvers */
#define PF_FLAG_SINGULAR 1 /* np.linalg.LinAlgError('Singular Matrix'), audioModel.py:858-861 */
#define PF_FLAG_TW_RESTART 2 /* sum(TW) < eps: the reference re-draws TW, audioModel.py:2023-2025 */

/* ---- library ---------------------------------------------------------------- */
const char* pf_last_error(void);
int pf_abi_version(void);
/* kernels launched by this library since load (bench.py "gpu_launches") */
unsigned long long pf_launch_count(void);
/* Make `device` current for this library's CUDA runtime (call once per process/thread
 * with the device the tensors live on). */
int pf_set_device(int device);

/* Strided 2-D copy between host and device (direction inferred from the pointers): `height` rows of
 * `width_bytes` bytes, pitches in bytes; complete on return.  Moves a rank's column range of a
 * user-visible parameter matrix (audioModel.py:1573, :1725 update those arrays in place). */
int pf_copy_2d(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes,
               int64_t height, void* stream);

/* ---- K1: STFT front end  (tftransforms/stft.py:3-69, audioModel.py:250-328) --- */
/* Framing + window + real FFT of `nch` channels in one launch.
 * pcm     : the samples in the layout the host has them in (pcm_format):
 *           PF_PCM_F64_PLANAR double [nch][L]; PF_PCM_I16 / PF_PCM_I32 / PF_PCM_F32
 *           interleaved [L][nch] as read from a WAV file.  Every sample is divided by
 *           pcm_div in float64 first (the reference's data / (1.1*max), audioObject.py:124-127)
 * window  : double [wlen] (host-built np.hanning etc.), nfft >= wlen, nfft = 2^k
 * X       : dtype planes [2*nch][F][ld] = (re, im) per channel, F = nfft/2+1,
 *           N = ceil(L/hop)+2 frames (stft.py:40)
 * psd_sum : double [F], sum over channels and frames of |X|^2 (for the annealing
 *           limits, audioModel.py:304-323); may be NULL
 * The FFT itself always runs in float64. */
#define PF_PCM_F64_PLANAR 0
#define PF_PCM_I16 1
#define PF_PCM_I32 2
#define PF_PCM_F32 3
/* Frame sharding: `pcm` may be a window [sample0, sample0 + L) of a longer signal of
 * L_total samples, and X then holds the frames [frame0, frame0 + N) of that signal
 * (whole signal: sample0 = 0, L = L_total, frame0 = 0, N = ceil(L/hop)+2). */
int pf_stft(const void* pcm, int pcm_format, double pcm_div, int nch, int64_t L, int64_t sample0,
            int64_t L_total, const double* window, int wlen, int hop, int nfft, void* X,
            int64_t frame0, int64_t N, int64_t ld, double* psd_sum, int dtype, void* stream);

/* peak[0] = max |x| over `count` samples in the host layout `pcm_format`, as
 * np.abs(data).max() gives it (audioObject.py:124-126: the scaling factor is 1.1 times this;
 * the most negative integer, which np.abs wraps onto itself, is skipped). */
int pf_pcm_peak(const void* pcm, int pcm_format, int64_t count, double* peak, void* stream);

/* ---- K6: inverse STFT with overlap-add  (tftransforms/stft.py:71-131) --------- */
/* Y       : dtype planes [2*nsig][F][ld]
 * synth   : double [wlen] synthesis window; norm : double [(N-1)*hop+wlen]
 *           = overlap-added synth*analysis (stft.py:117-129), host-built
 * out     : double [nsig][Lout] ; sample t is the overlap-added sample t + drop: FASST drops
 *           the first half window (drop = wlen/2, stft.py:123), the SIMM back end keeps it
 *           (drop = 0, SeparateLeadStereo/separateLeadFunctions.py:163-233, whose patched
 *           normalisation sequence :218-221 is simply passed in `norm`)
 * pcm     : optional int16 [Lout][nsig] interleaved, = (int16)trunc(out*maxdata)
 *           (audioModel.py:1227-1229), or round-half-even when pcm_round
 *           (SeparateLeadStereoTF.py:1826-1827) ; NULL to skip */
/* norm [hop (N-1) + wlen] = overlap-added prod = synth * analysis window, frames accumulated in
 * ascending order like the reference (stft.py:112-121); zeros -> 1 and the edge patch of the
 * SIMM inverse (separateLeadFunctions.py:218-221) are applied by the caller */
int pf_overlap_norm(const double* prod, int wlen, int hop, int64_t N, double* norm, void* stream);
int pf_istft(const void* Y, int nsig, int F, int64_t N, int64_t ld, const double* synth,
             const double* norm, int wlen, int hop, int nfft, double* out, int64_t Lout,
             int16_t* pcm, double maxdata, int64_t drop, int pcm_round, int dtype, void* stream);

/* Y[c1] = sum_c2 W[c1][c2] X[c2] per bin: the filtering step of filter_stft
 * (tftransforms/stft.py:181-193).  X, Y planes [2 nc][F][ld]; W complex128 [nc][nc][F] (wn = 0)
 * or [nc][nc][F][wn] with wn >= N frames */
int pf_apply_filter(const void* X, const void* W, void* Y, int nc, int F, int64_t N, int64_t ld,
                    int64_t wn, int dtype, void* stream);

/* ---- K6: Wiener filter  (audioModel.py:1088-1236, :1327-1467) ----------------- */
/* For every bin: Sigma = sum_j v_j R_j + s2 I, Y_g = (sum_{j in g} v_j R_j) Sigma^-1 x for
 * `ngroups` output signals.  Y planes [ngroups][4][F][ld] (re0, im0, re1, im1).
 * A: complex128 [R][2][F]; src_of_sub: host int[R]; group_of_src: host int[J], the output
 * group of each spatial component (-1 = none).  workspace: >= F*(4J+J(J+1)/2)*8 bytes.
 * noise_psd is the PSD of the LAST GEM iteration (audioModel.py:1385). */
int pf_wiener_stereo(const void* X, const void* V, const void* A, const int* src_of_sub, int R,
                     int J, const double* noise_psd, const int* group_of_src, int ngroups, int F,
                     int64_t N, int64_t ld, void* Y, void* workspace, int64_t workspace_bytes,
                     int dtype, void* stream);

/* ---- K2: fused E-step  (audioModel.py:580-764, tools/signalTools.py:132-196) --- */
/* Measurement aid (bench.py's roofline): pf_estep_timing(1) makes every following E-step call record
 * a CUDA event pair on its stream right before and after the per-bin kernel alone;
 * pf_estep_timing_read() waits for them and returns their summed elapsed time and their number. */
int pf_estep_timing(int enable);
int pf_estep_timing_read(double* total_ms, int* launches);
/* Workspace planning (host only, no device call): bytes needed by pf_estep_stereo for this
 * shape.  *nsplit CTAs share one frequency row and *chunk is the number of frames each of them
 * covers (chunk * nsplit >= N); the CTAs of a row take its passes in turn, so the frames of one
 * CTA are not contiguous. */
int pf_estep_plan(int J, int64_t N, int dtype, int64_t* chunk, int* nsplit,
                  int64_t* workspace_bytes, int F);
/* X       : dtype planes [4][F][ld] (re0, im0, re1, im1): Cx = x x^H is rank one
 *           (audioModel.py:293-302) so the kernel reads x instead of Cx
 * V       : dtype planes [J][F][ld], power of each spatial component
 *           (comp_spat_comp_power, audioModel.py:430-498)
 * A       : complex128 [R][2][F] mixing vectors per sub-source (audioModel.py:562-576)
 * src_of_sub : host int[R], spatial component of each sub-source
 * noise_psd  : double [F]
 * hatW    : dtype planes [J][F][ld]  = rank-mean of |Re diag| (audioModel.py:727-729,:408-414)
 * hat_Rss : complex128 [F][R][R], hat_Rxs : complex128 [F][2][R]
 * ll_f    : double [F], sum over frames of log(det Sigma * pi) + x^H Sigma^-1 x
 *           (loglik = -sum(ll_f) / (F N), audioModel.py:660-664)
 * N_norm  : number of frames the means hat_Rss / hat_Rxs are taken over (0 = N); when the
 *           frames of a mixture are sharded over GPUs, pass the total and sum the results */
int pf_estep_stereo(const void* X, const void* V, const void* A, const int* src_of_sub, int R,
                    int J, const double* noise_psd, int F, int64_t N, int64_t ld, void* hatW,
                    void* hat_Rss, void* hat_Rxs, double* ll_f, void* workspace,
                    int64_t workspace_bytes, int64_t N_norm, int dtype, void* stream);
/* The same for REAL mixing vectors (Im A = 0: instantaneous mixing, audioModel.py:2349-2393) when
 * the statistics only feed the instantaneous mixing update, which takes their real parts
 * (np.real(np.mean(...)), audioModel.py:818-820): Sigma is real symmetric, Im M01 is neither
 * needed for hatW nor accumulated (112 instead of 143 FP64 operations per bin).  hatW, ll_f and the
 * real parts of hat_Rss / hat_Rxs are those of pf_estep_stereo; the imaginary parts are zero. */
int pf_estep_stereo_inst(const void* X, const void* V, const void* A, const int* src_of_sub, int R,
                    int J, const double* noise_psd, int F, int64_t N, int64_t ld, void* hatW,
                    void* hat_Rss, void* hat_Rxs, double* ll_f, void* workspace,
                    int64_t workspace_bytes, int64_t N_norm, int dtype, void* stream);

/* ---- K2 / K6 for I = 2..4 channels (extension: the reference is stereo only, audioModel.py:394,
 * :605, :1127; definitions as oracle/fasst_oracle.py: estep_general -- batched I x I Hermitian
 * inverse, determinant clamp on the generic determinant).  X: planes [2 I][F][ld]
 * (re, im per channel), A: complex128 [R][I][F], hat_Rxs: complex128 [F][I][R]; everything
 * else as pf_estep_stereo / pf_wiener_stereo (Y: planes [ngroups][2 I][F][ld]). */
int pf_estep_multi_plan(int I, int J, int F, int64_t N, int64_t* chunk, int* nsplit,
                        int64_t* workspace_bytes);
int pf_estep_multi(const void* X, const void* V, const void* A, const int* src_of_sub, int R,
                   int J, int I, const double* noise_psd, int F, int64_t N, int64_t ld, void* hatW,
                   void* hat_Rss, void* hat_Rxs, double* ll_f, void* workspace,
                   int64_t workspace_bytes, int64_t N_norm, int dtype, void* stream);
int pf_wiener_multi(const void* X, const void* V, const void* A, const int* src_of_sub, int R,
                    int J, int I, const double* noise_psd, const int* group_of_src, int ngroups,
                    int F, int64_t N, int64_t ld, void* Y, void* workspace,
                    int64_t workspace_bytes, int dtype, void* stream);

/* ---- K3: spatial M-step  (audioModel.py:766-889) ------------------------------- */
/* Instantaneous mixing: f-summed real statistics (audioModel.py:816-826).
 * stats : double [I*n_upd + n_upd*n_upd]; under frequency sharding the caller
 * all-reduces `stats` between the two calls. upd/oth: host int arrays. */
int pf_mix_inst_stats(const void* hat_Rss, const void* hat_Rxs, const void* A, const int* upd,
                      int n_upd, const int* oth, int n_oth, int R, int I, int F, double* stats,
                      void* stream);
/* Solve and broadcast over the local frequencies (audioModel.py:830-841).
 * F_total = number of frequencies the statistics were summed over. */
int pf_mix_inst_solve(const double* stats, double F_total, const int* upd, int n_upd, int I,
                      int F, void* A, int* flags, void* stream);
/* Convolutive mixing: A[:, :, f] = solve(hat_Rss[f]^T, hat_Rxs[f]^T) (audioModel.py:844-863) */
int pf_mix_conv_solve(const void* hat_Rss, const void* hat_Rxs, int R, int I, int F, void* A,
                      int* flags, void* stream);

/* ---- K4: spectral M-step  (audioModel.py:430-498, :1469-1727) ------------------- */
/* V[f,n] (+)= sum_k W[f,k] H[k,n];  W dtype [F][ldw], H dtype [K][ldh], V dtype [F][ldv] */
int pf_spec_power(const void* W, int ldw, const void* H, int64_t ldh, void* V, int64_t ldv,
                  int F, int K, int64_t N, int accumulate, int dtype, void* stream);
/* C[m][n] = sum_k A[m][k] B[k][n] for the small factor products FB.FW (dtype, row-major) */
int pf_small_matmul(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M,
                    int K, int Nc, int dtype, void* stream);
/* FB / FW update contractions over frames (audioModel.py:1521-1631):
 *   num[f,k] = sum_n (hatW/P^2*O)[f,n] G[k,n],  den[f,k] = sum_n (O/P)[f,n] G[k,n]
 * P, O are clamped at eps=1e-10 in the kernel.  Partial sums per frame split:
 * num_partial/den_partial double [nsplit][F][K]. */
int pf_nmf_fb_plan(int F, int K, int64_t N, int dtype, int64_t* chunk, int* nsplit);
int pf_nmf_fb_contract(const void* hatW, const void* P, const void* O, int64_t ld, const void* G,
                       int64_t ldg, int F, int K, int64_t N, double* num_partial,
                       double* den_partial, int64_t chunk, int nsplit, int dtype, void* stream);
/* TW update contractions over frequencies (audioModel.py:1634-1727), with the
 * component's own power P' = max(W H, eps) formed on the fly from the updated W:
 *   num[k,n] = sum_f W[f,k] (O hatW / P'^2)[f,n],  den[k,n] = sum_f W[f,k] (O / P')[f,n]
 * Partial sums per frequency split: double [fsplit][K][ldo]. */
int pf_nmf_tw_plan(int F, int K, int64_t N, int dtype, int* fchunk, int* fsplit);
/* scratch_plane: optional dtype [F][ld] work plane.  With it, float32 planes and K <= 32 the
 * contractions run on the tensor cores (tcgen05 kind::tf32, 3xTF32 split): P' = W H is written
 * to the scratch plane by one kernel and contracted by a second; without it (or for float64)
 * P' is formed on the fly on the CUDA cores. */
int pf_nmf_tw_contract(const void* hatW, const void* O, int64_t ld, const void* W, int ldw,
                       const void* H, int64_t ldh, int F, int K, int64_t N, double* num_partial,
                       double* den_partial, int64_t ldo, int fchunk, int fsplit,
                       void* scratch_plane, int dtype, void* stream);
/* Frequency-sharded TW update (the K x N numerators / denominators are the one large
 * exchange of the frequency partition): out[w][q][k][i] = sum_s part_q[s][k][w c + i], c = ld / world,
 * q = 0 num / 1 den, in the plane type -- the chunk-major input of the reduce-scatter over frames. */
int pf_tw_pack_chunks(const double* num_partial, const double* den_partial, int nsplit,
                      int64_t split_stride, int K, int64_t ld, int world, int Kmax, void* out,
                      int dtype, void* stream);
/* out[i] = sum_s in[s][i] in a fixed order */
int pf_sum_splits(const double* in, int nsplit, int64_t count, double* out, void* stream);
/* theta[r][c] *= (num[r][c] / max(den[r][c], 1e-10))^omega (audioModel.py:1573,:1725) */
int pf_mult_update(void* theta, int64_t ldt, const double* num, const double* den, int64_t ldnd,
                   int rows, int64_t cols, double omega, int dtype, void* stream);

/* the same with the split partial sums reduced on the fly: num/den_partial[s] is a
 * [rows][ldnd] matrix at offset s * split_stride (doubles) */
int pf_mult_update_splits(void* theta, int64_t ldt, const double* num_partial,
                          const double* den_partial, int nsplit, int64_t split_stride,
                          int64_t ldnd, int rows, int64_t cols, double omega, int dtype,
                          void* stream);

/* ---- K4 for general factor structures (several factors per spectral component, free FW, large
 * dictionaries: multiChanSourceF0Filter, audioModel.py:2551-2760).  The planes of
 * update_spectral_components (audioModel.py:1513-1520, :1565-1571) are formed once and contracted
 * with pf_gemm_tf32x3(_splitk). */
/* out [F][2 ld] = ( hatW / max(P,eps)^2 * max(O,eps) | max(O,eps) / max(P,eps) ), eps = 1e-10,
 * zero in the padding columns n >= N */
/* Ptot / Pminus (NULL: none): the correlation penalty lambdaCorr of audioModel.py:1484-1703,
 * c = lambda Pminus / max(Ptot^2, eps): den = O (1/P + c), num = (hatW / P^2 + 2 c P / Ptot) O */
int pf_gem_ratio_planes(const void* hatW, const void* P, const void* O, void* out, int F,
                        int64_t N, int64_t ld, const void* Ptot, const void* Pminus, double lambda,
                        int dtype, void* stream);
/* Ptot = max(sum_j V_j, eps), Pminus = Ptot - max(V_own, eps) (clamped at eps when `clamp`), V planes
 * [J][F][ld] (audioModel.py:1484-1508) */
int pf_corr_planes(const void* V, int J, int own, void* Ptot, void* Pminus, int F, int64_t N,
                   int64_t ld, int clamp, int dtype, void* stream);
/* out[r] = sum_c M[r][c] in float64 (the row means of the time blobs, audioModel.py:2026-2030) */
int pf_row_sums(const void* M, int64_t ldm, int rows, int64_t cols, double* out, int dtype,
                void* stream);
/* out (=, +=) a * b elementwise on [F][ld] planes (b = NULL: a); zero in the padding */
int pf_mul_planes(const void* a, const void* b, void* out, int F, int64_t N, int64_t ld,
                  int accumulate, int dtype, void* stream);
/* theta[r][c] *= (num[r][c] / max(den[r][c], 1e-10))^omega with num / den of theta's type
 * (GEMM outputs)  (audioModel.py:1573, :1631, :1725) */
int pf_mult_update_same(void* theta, int64_t ldt, const void* num, int64_t ldn, const void* den,
                        int64_t ldd, int rows, int64_t cols, double omega, int dtype,
                        void* stream);

/* Sparsity re-weighting of the source activations of multiChanSourceF0Filter
 * (audioModel.py:2981-3014; median filter: tools/signalTools.py:13-24): TW [K][ldt] *= Gaussian
 * mask around the median-filtered (window `length`) barycentre of every frame, of variance
 * sigma = exp(log_sigma0 + slope * (iter_dev[0] - 1)) (the device iteration counter has already
 * advanced when this runs); work: 2 N doubles */
int pf_sparsity_reweigh(void* TW, int64_t ldt, int K, int64_t N, int length, double log_sigma0,
                        double slope, const int* iter_dev, double* work, int dtype, void* stream);

/* ---- K5: renormalisation  (audioModel.py:1980-2040) ----------------------------- */
int pf_spat_energy(const void* A, const int* src_of_sub, int R, int J, int I, int F,
                   double* sums, void* stream);
int pf_spat_scale(void* A, const int* src_of_sub, int R, int I, int F, const double* sums,
                  const double* counts, void* stream);
int pf_fb_scale_colmax(void* FB, int ldw, int F, int K, const double* sums, const double* counts,
                       int j, double* colmax, int dtype, void* stream);
int pf_fw_renorm(void* FW, int ldfw, int Kb, int Kw, const double* colmax, double* w, double* w2,
                 int dtype, void* stream);
int pf_scale_matrix(void* M, int64_t ldm, int rows, int64_t cols, const double* s, int by_row,
                    int divide, double* total, int dtype, void* stream);

/* totals[s] < eps  ->  *flags |= PF_FLAG_TW_RESTART ; totals are reset to zero */
int pf_check_totals(double* totals, int count, double eps, int* flags, const int* iter_dev,
                    int* first_iter, void* stream);

/* ---- dense float32 GEMM on the tensor cores (SIMM: every np.dot of SIMM.py:303-393, :613-941) */
/* C[M x N] = op(A) op(B), row-major float32, tcgen05 kind::tf32 with the 3xTF32 split
 * (float32-class accuracy).  transA: A is given as A^T, i.e. stored [K][M]; transB: B is given
 * as B^T, i.e. stored [N][K].  Leading dimensions multiples of 4, pointers 16-byte aligned;
 * K % 4 == 0 (zero padded) whenever an operand is contiguous along K. */
int pf_gemm_tf32x3(const float* A, int64_t lda, int transA, const float* B, int64_t ldb,
                   int transB, float* C, int64_t ldc, int M, int N, int K, void* stream);

/* The same product with the contraction split over CTAs (few output tiles, long K: the
 * contractions over the frame axis of SIMM.py:354-362, :376-381, :829-843, :909-916).  The partial
 * products go to `workspace` (pf_gemm_splitk_plan gives its size) and are summed in a fixed
 * order.  ldc must equal N rounded up to 4; the padding columns of C are written as zero. */
int pf_gemm_splitk_plan(int M, int N, int K, int* ksplit, int64_t* workspace_bytes);
int pf_gemm_tf32x3_splitk(const float* A, int64_t lda, int transA, const float* B, int64_t ldb,
                          int transB, float* C, int64_t ldc, int M, int N, int K, float* workspace,
                          int64_t workspace_bytes, void* stream);

/* ---- K7 / K8: SIMM and Stereo_SIMM  (SeparateLeadStereo/SIMM/SIMM.py:46-395, :397-943) -----------
 * float32 planes, frames contiguous, ldn = N rounded up to 4; nch = 1 (SIMM) or 2 (Stereo_SIMM):
 *   SF0, SPHI [F][ldn];  SX, hat, SM [F][nch ldn] (channel c in columns [c ldn, c ldn + N));
 *   work planes [F][2 nch ldn].  a2 = alpha^2 (float[nch], device; 1 for the mono model),
 *   b2 = beta^2 (float[nch][ldr], device).  Model: hat_c = a2_c SF0 SPHI + (WM b2_c) HM.
 * Producers write zero into the padding columns of the work planes. */
/* hat_c = max(a2_c SF0 SPHI + SM_c, eps) is formed in registers by every consumer below (the
 * hat planes are not stored).
 * out = (num | den): c_c = a2_c other/hat_c, num = sum_c c_c SX_c/hat_c, den = sum_c c_c, with
 * other = SPHI (other_is_sf0 = 0: HF0 update) or SF0 (1: HPHI / HGAMMA updates)
 * (SIMM.py:304-305, :319-320, :352-353; :622-640, :685-700, :776-790) */
int pf_simm_lead_terms(const float* SM, const float* SF0, const float* SPHI, const float* SX,
                       const float* a2, int other_is_sf0, float* out, int nch, int F, int64_t N,
                       int64_t ldn, void* stream);
/* out = (T_0 .. T_{nch-1} | I_0 .. I_{nch-1}), T = SX/hat^2, I = 1/hat; sq_clamp selects the
 * stereo clamping max(hat^2, eps) (SIMM.py:741-750) against the mono one (:335-337) */
int pf_simm_acc_terms(const float* SM, const float* SF0, const float* SPHI, const float* SX,
                      const float* a2, float* out, int nch, int sq_clamp, int F, int64_t N,
                      int64_t ldn, void* stream);
/* hat_c = max(a2_c SF0 SPHI + SM_c, eps) written out (the separation masks read it)
 * (SIMM.py:313; :655-664) */
int pf_simm_hat(const float* SM, const float* SF0, const float* SPHI, const float* a2, float* hat,
                int nch, int F, int64_t N, int64_t ldn, void* stream);
int64_t pf_simm_reduce_workspace_bytes(void);
/* out[0] = sum_c IS(SX_c | hat_c)   (ISDistortion, SIMM.py:35-44), float64, fixed order */
int pf_simm_is_divergence(const float* SX, const float* SM, const float* SF0, const float* SPHI,
                          const float* a2, int nch, int F, int64_t N, int64_t ldn,
                          double* workspace, double* out, void* stream);
/* alpha update of Stereo_SIMM (SIMM.py:869-896): alpha double[2] and a2 float[2] on the device */
int pf_simm_alpha_update(const float* SX, const float* SM, const float* SF0, const float* SPHI,
                         int F, int64_t N, int64_t ldn, double omega, double* workspace,
                         double* alpha, float* a2, void* stream);
/* theta[r][n] *= (num/max(den,eps))^omega, num = sum_c w[c][r] C[r][c ldn + n],
 * den = sum_c w[c][r] C[r][(nch+c) ldn + n] (w = NULL: 1), then max(theta, floor_value) if > 0
 * (SIMM.py:308-310, :321-323, :338-343; :641-650, :701-708, :752-763) */
int pf_simm_update_rows(float* theta, int64_t ldt, const float* C, int64_t ldc, int nch,
                        int64_t ldn, const float* w, int wld, double omega, double floor_value,
                        int rows, int64_t N, void* stream);
/* HPHI[k][:] *= rowscale[k] (if not NULL); s[n] = sum_k HPHI[k][n]; HPHI[:,n] /= s[n] where s>0
 * (SIMM.py:324-326, :361-365) */
int pf_simm_hphi_normalise(float* HPHI, int64_t ldn, int K, const float* rowscale, int64_t N,
                           float* s_out, void* stream);
/* P[r][n] *= s[n] (s has ld entries) / P[r][n] *= s[r] */
int pf_simm_scale_columns(float* P, int64_t ld, int rows, int64_t N, const float* s, void* stream);
int pf_simm_scale_rows(float* P, int64_t ld, int rows, int64_t N, const float* s, void* stream);
/* HGAMMA update and column normalisation from tn/td = (num | den) HPHI^T  (SIMM.py:354-362) */
int pf_simm_hgamma_update(float* HGAMMA, int ldhg, const float* WGAMMA, int ldwg, const float* tn,
                          const float* td, int ldt, int F, int P, int K, double omega,
                          float* s_out, void* stream);
/* WM update and column normalisation from D = [2 nch][F][ldr] (T_c HM^T, I_c HM^T); the
 * denominator is clamped in the mono model only  (SIMM.py:376-388; :829-866) */
int pf_simm_wm_update(float* WM, int ldr, int R, const float* D, int nch, const float* b2,
                      int clamp_den, double omega, int F, float* s_out, void* stream);
/* beta update of Stereo_SIMM from D = [4][F][ldr]  (SIMM.py:909-941); beta double[2][ldr] */
int pf_simm_beta_update(const float* WM, int ldr, int R, const float* D, int F, double omega,
                        double* beta, float* b2, void* stream);
/* SX_c = |X_c|^2 from STFT planes X [2 nch][F][ldx] (re, im per channel)
 * (SeparateLeadStereoTF.py:843-917) */
int pf_simm_power(const float* X, int64_t ldx, float* SX, int nch, int F, int64_t N, int64_t ldn,
                  void* stream);
/* Wiener masks of the lead / accompaniment separation applied to the mixture STFT
 * (SeparateLeadStereoTF.py:1762-1871; eps_hat = 1e-9 there, SeparateLeadStereoTF.py:31):
 *   Y[s = c]       = a2_c SF0 SPHI / max(hat_c, eps_hat) X_c      (lead, channel c)
 *   Y[s = nch + c] = SM_c          / max(hat_c, eps_hat) X_c      (accompaniment)
 * with hat_c = a2_c SF0 SPHI + SM_c; Y: planes [2 * 2 nch][F][ldx] (re, im per signal) */
int pf_simm_masks(const float* SM, const float* SF0, const float* SPHI, const float* a2,
                  const float* X, int64_t ldx, float* Y, double eps_hat, int nch, int F,
                  int64_t N, int64_t ldn, void* stream);
/* WMs[c][f][r] = WM[f][r] b2[c][r] (b2 = NULL: 1), zero in the padding columns r >= R */
int pf_simm_wm_scaled(const float* WM, int ldr, int R, const float* b2, int nch, int F, float* WMs,
                      void* stream);

/* ---- IS-NMF initialisers  (tools/nmf.py:24-159, audioModel.py:2091-2222; eps = 1e-10) -----------
 * float32, layout as the SIMM kernels; the contractions are pf_gemm_tf32x3(_splitk). */
/* out = (T | I) [F][2 ldn]: T = SX / max(hat^2, eps), I = 1 / max(hat, eps) */
int pf_nmf_is_terms(const float* hat, const float* SX, float* out, double eps, int F, int64_t N,
                    int64_t ldn, void* stream);
/* H[k][n] *= C[k][n] / max(C[k][ldn + n], eps)   (C = W^T (T | I)) */
int pf_nmf_update_rows(float* H, int64_t ldh, const float* C, int64_t ldc, int64_t ldn, double eps,
                       int rows, int64_t N, void* stream);
/* W *= D[0] / max(D[1], eps) (D = [2][F][ldk]: T H^T, I H^T); columns normalised to sum one
 * (a zero sum counts as one); s_out[k] = the sums */
int pf_nmf_w_update(float* W, int ldk, int K, const float* D, double eps, int F, float* s_out,
                    void* stream);
/* out[f][n] = mean_c |X_c[f][n]|^2 from STFT planes X [2 nch][F][ldx]  (audioModel.py:2150-2158) */
int pf_mono_power(const float* X, int64_t ldx, int nch, float* out, int F, int64_t N, int64_t ldn,
                  void* stream);

/* ---- Viterbi melody tracking  (SeparateLeadStereo/tracking/_tracking.pyx:11-93) ----------------
 * log_density_ns: float64 [N][S] (frame major), log_prior [S], log_trans [S][S] (row = from);
 * path: int64 [N].  Same recursion and tie breaking (smallest index among equal maxima) as the
 * reference's Cython module: the path is bit-identical. */
int64_t pf_viterbi_workspace_bytes(int S, int64_t N);
int pf_viterbi(const double* log_density_ns, const double* log_prior, const double* log_trans, int S,
               int64_t N, void* workspace, int64_t workspace_bytes, long long* path, void* stream);

/* ---- Glottal-source F0 dictionary WF0  (SeparateLeadStereo/separateLeadFunctions.py:
 * generate_ODGD_spec :888-949, generate_ODGD_spec_chirped :1010-1072, the column loops of
 * generate_WF0_chirped :237-345 and generate_WF0_TR_chirped :696-886) -----------------------
 * Column c = power spectrum of one windowed frame of the KLGLOTT88 waveform
 *   x_c(t) = Re sum_{h=1}^{npart[c]} A_h exp(2 pi i h (f1[c] tau + (f2[c]-f1[c]) tau^2 / (2 T))),
 *   tau = t / fs, T = Lsig / fs, A_h the glottal amplitudes (:917-929) of F0 = (f1+f2)/2 with
 *   opening coefficient Ot; f1 = f2 gives the plain (unchirped) comb.
 * out[c][r] = | sum_{k < wlen} window[k] x_c(t_begin + k) exp(-2 pi i r k / nfft) |^2, r < rows
 *   (x_c = 0 outside [0, Lsig)).  f1, f2: double [ncols]; npart: int32 [ncols], all <=
 *   max_partials <= 2048; window: double [wlen], wlen <= nfft <= 8192; out: double [ncols][rows];
 *   all device pointers. */
int pf_wf0_combs(const double* f1, const double* f2, const int* npart, int ncols, int max_partials,
                 double fs, double Ot, int64_t Lsig, int64_t t_begin, const double* window,
                 int wlen, int nfft, int rows, double* out, void* stream);

/* ---- tcgen05 self-test (pins the descriptor / layout conventions of csrc/tc.cuh) ------ */
/* D[128][N] = A B^T in tf32 (split3: 3xTF32, fp32-class accuracy).  A: a_mn ? [K][128] :
 * [128][K]; B: b_mn ? [K][N] : [N][K]; float32 device buffers; N, K multiples of 32. */
int pf_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn,
                   int split3, void* stream);

/* ---- GEM loop glue  (audioModel.py:330-382) --------------------------------------- */
/* noise[f] = ((sqrt_lim0[f]*(I-i) + sqrt_lim1[f]*i)/I)^2 with i read from *iter_dev
 * (audioModel.py:368-373), so that a captured CUDA graph can be replayed. */
int pf_noise_anneal(const double* sqrt_lim0, const double* sqrt_lim1, const int* iter_dev,
                    int n_iter, int F, double* noise, void* stream);
/* ll_sum[0] = sum_f ll_f[f]  (fixed order) */
int pf_ll_reduce(const double* ll_f, int F, double* ll_sum, void* stream);
/* logliks[*iter_dev] = -ll_sum[0] / bins ; if advance, ++*iter_dev */
int pf_ll_store(const double* ll_sum, double bins, double* logliks, int* iter_dev, int advance,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYFASST_B200_H_ */
