"""Host orchestration of the SIMM / Stereo_SIMM multiplicative-update loops on the GPU.

Replaces the bodies of pyfasst/SeparateLeadStereo/SIMM/SIMM.py: `SIMM` (:46-395) and
`Stereo_SIMM` (:397-943).  All state lives in HBM as float32 (layout: csrc/simm.cu); one
iteration enqueues ~60 kernels (tensor-core GEMMs for every contraction with an F x N operand,
bandwidth-bound elementwise / reduction kernels for the rest) and never synchronises with the
host.  The mono model is the one-channel case of the stereo one (alpha = beta = 1); the
differences the reference has between the two (clamps, floors: quirk Q12 of SURVEY.md) are
explicit flags below.

Differences from the reference that are deliberate (all within float32 rounding):
  * after a column rescaling of HF0 (`HF0 *= sumHPHI`, SIMM.py:326, :365), SF0 = WF0 HF0 is
    rescaled by the same factors instead of being recomputed with an NF0-sized GEMM;
  * the model power hat is never stored: every kernel that needs it forms
    max(alpha^2 SF0 SPHI + SM, eps) in registers, so it is clamped to eps also before the first
    update (the reference clamps it on use, SIMM.py:268 vs :304) -- identical unless hat < 1e-20.
"""
import numpy as np

EPS = 1e-20


def _ru4(n):
    return (int(n) + 3) // 4 * 4


class SimmEngine(object):
    """nch = 1: SIMM; nch = 2: Stereo_SIMM.  `kernels` is pyfasst_b200._lib.CudaKernels()
    (tests inject the NumPy specification of the kernels instead)."""

    def __init__(self, kernels, SX_list, WF0, WGAMMA, HGAMMA, HPHI, HF0, WM, HM, betaR=None,
                 omega=1.0, update_hgamma=True, compute_error=False, n_iter=0):
        import torch
        self.torch = torch
        self.k = kernels
        self.dev = kernels.device
        self.nch = len(SX_list)
        assert self.nch in (1, 2)
        self.stereo = self.nch == 2
        self.omega = float(omega)
        self.update_hgamma = update_hgamma or not self.stereo
        self.compute_error = compute_error and self.stereo
        F, N = SX_list[0].shape
        self.F, self.N = F, N
        self.NF0, self.P = WF0.shape[1], WGAMMA.shape[1]
        self.K, self.R = HGAMMA.shape[1], WM.shape[1]
        self.ldn, self.ldf0 = _ru4(N), _ru4(self.NF0)
        self.ldp, self.ldk, self.ldr = _ru4(self.P), _ru4(self.K), _ru4(self.R)
        if not self.stereo and self.R != 1 and self.R != N:
            # `HM *= sumWM` (SIMM.py:388) broadcasts [R] against [R, N]
            raise ValueError("operands could not be broadcast together with shapes (%d,%d) (%d,)"
                             % (self.R, N, self.R))
        nch, ldn = self.nch, self.ldn
        f32 = torch.float32

        def zeros(*shape, dtype=f32):
            return torch.zeros(shape, dtype=dtype, device=self.dev)

        def upload(a, rows, cols):
            a = np.asarray(a, dtype=np.float64)
            buf = np.zeros((rows, cols), dtype=np.float32)
            buf[:a.shape[0], :a.shape[1]] = a
            return torch.from_numpy(buf).to(self.dev)

        self.SX = zeros(F, nch * ldn)
        for c, sx in enumerate(SX_list):
            self.SX[:, c * ldn:c * ldn + N] = torch.from_numpy(
                np.ascontiguousarray(sx, dtype=np.float32)).to(self.dev)
        self.SM = zeros(F, nch * ldn)
        self.SF0, self.SPHI = zeros(F, ldn), zeros(F, ldn)
        self.work = zeros(F * 2 * nch * ldn)
        self.work_lead = self.work[:F * 2 * ldn].view(F, 2 * ldn)  # (num | den)
        self.work_acc = self.work.view(F, 2 * nch * ldn)           # (T_c .. | I_c ..)
        self.WF0 = upload(WF0, F, self.ldf0)
        self.HF0 = upload(HF0, self.ldf0, ldn)
        self.WGAMMA = upload(WGAMMA, F, self.ldp)
        self.HGAMMA = upload(HGAMMA, self.ldp, self.ldk)
        self.WPHI = zeros(F, self.ldk)
        self.HPHI = upload(HPHI, self.ldk, ldn)
        self.WM = upload(WM, F, self.ldr)
        self.WMs = zeros(nch, F, self.ldr)
        self.HM = upload(HM, self.ldr, ldn)
        self.C_f0 = zeros(self.NF0, 2 * ldn)
        self.C_phi = zeros(self.ldk, 2 * ldn)
        self.C_hm = zeros(self.ldr, 2 * nch * ldn)
        self.tn, self.td = zeros(F, self.ldk), zeros(F, self.ldk)
        self.D = zeros(2 * nch, F, self.ldr)
        self.s_n, self.s_k, self.s_r = zeros(max(ldn, self.ldr)), zeros(self.ldk), zeros(self.ldr)
        self.a2 = torch.ones(2, dtype=f32, device=self.dev)
        self.alpha = torch.full((2,), 0.5 if self.stereo else 1.0, dtype=torch.float64,
                                device=self.dev)
        self.beta = self.b2 = None
        if self.stereo:
            self.a2.fill_(0.25)
            b = np.zeros((2, self.ldr))
            b[0, :self.R] = np.asarray(betaR, dtype=np.float64)
            b[1, :self.R] = 1 - b[0, :self.R]
            self.beta = torch.from_numpy(b).to(self.dev)
            self.b2 = torch.from_numpy((b ** 2).astype(np.float32)).to(self.dev)
        ws = max(kernels.gemm_splitk_workspace_bytes(F, self.K, ldn),
                 kernels.gemm_splitk_workspace_bytes(F, self.R, ldn), 16)
        self.splitk_ws = zeros(ws // 4)
        self.red_ws = zeros(max(kernels.simm_reduce_workspace_bytes() // 8, 2),
                            dtype=torch.float64)
        self.reco = zeros(n_iter * 5 * 2 + self.NF0 * 2 + 1, dtype=torch.float64)
        self.counter = 1
        # initial model (SIMM.py:263-268; :584-592)
        k = self.k
        k.small_matmul(self.WGAMMA, self.HGAMMA, self.WPHI)
        k.gemm_view(self.WF0, self.HF0, self.SF0, F, N, self.ldf0)
        k.spec_power(self.WPHI, self.HPHI, self.SPHI, N, False)
        self._model_acc()
        if self.compute_error:
            self._error(0)

    # -- building blocks ------------------------------------------------------------------
    def _model_acc(self):
        """SM_c = (WM beta_c^2) HM."""
        k, F, N, ldn = self.k, self.F, self.N, self.ldn
        k.simm_wm_scaled(self.WM, self.R, self.b2, self.nch, F, self.WMs)
        for c in range(self.nch):
            k.gemm_view(self.WMs[c], self.HM, self.SM[:, c * ldn:(c + 1) * ldn], F, N, self.ldr)

    # The model power hat_c = max(a2_c SF0 SPHI + SM_c, eps) is never stored: the kernels below
    # form it in registers from the planes it is made of, which are kept current instead.
    def hat_planes(self):
        """hat as a [F, nch * ldn] device tensor (for the separation masks and tests)."""
        hat = self.torch.zeros_like(self.SM)
        self.k.simm_hat(self.SM, self.SF0, self.SPHI, self.a2, hat, self.nch, self.F, self.N,
                        self.ldn)
        return hat

    def _error(self, slot):
        self.k.simm_is_divergence(self.SX, self.SM, self.SF0, self.SPHI, self.a2, self.nch, self.F,
                                  self.N, self.ldn, self.red_ws, self.reco[slot:slot + 1])

    def _lead_terms(self, other_is_sf0):
        self.k.simm_lead_terms(self.SM, self.SF0, self.SPHI, self.SX, self.a2, other_is_sf0,
                               self.work_lead, self.nch, self.F, self.N, self.ldn)

    def _acc_terms(self):
        self.k.simm_acc_terms(self.SM, self.SF0, self.SPHI, self.SX, self.a2, self.work_acc,
                              self.nch, self.stereo, self.F, self.N, self.ldn)

    def _acc_products(self):
        """D[q] = plane_q HM^T for the 2 nch accompaniment planes (contraction over frames)."""
        k, F, N, ldn = self.k, self.F, self.N, self.ldn
        self._acc_terms()
        for q in range(2 * self.nch):
            k.gemm_view(self.work_acc[:, q * ldn:(q + 1) * ldn], self.HM, self.D[q], F, self.R,
                        ldn, transB=True, workspace=self.splitk_ws)

    def _rescale_lead(self):
        """HF0 *= s_n (SIMM.py:326) and the same column scaling of SF0 = WF0 HF0."""
        self.k.simm_scale_columns(self.HF0, self.NF0, self.N, self.s_n)
        self.k.simm_scale_columns(self.SF0, self.F, self.N, self.s_n)

    # -- one iteration: HF0, HPHI, HM, HGAMMA, WM (, alpha, beta) -----------------------------
    def iterate(self):
        k, F, N, ldn, om = self.k, self.F, self.N, self.ldn, self.omega
        # HF0 (SIMM.py:303-315; :622-664)
        self._lead_terms(False)
        k.gemm_view(self.WF0, self.work_lead, self.C_f0, self.NF0, 2 * ldn, F, transA=True)
        k.simm_update_rows(self.HF0, self.C_f0, 1, ldn, None, om, 0.0, self.NF0, N)
        k.gemm_view(self.WF0, self.HF0, self.SF0, F, N, self.ldf0)
        if self.compute_error:
            self._error(self.counter)
        self.counter += 1
        # HPHI (:319-331; :685-730)
        self._lead_terms(True)
        k.gemm_view(self.WPHI, self.work_lead, self.C_phi, self.K, 2 * ldn, F, transA=True)
        k.simm_update_rows(self.HPHI, self.C_phi, 1, ldn, None, om, 0.0, self.K, N)
        k.simm_hphi_normalise(self.HPHI, self.K, None, N, self.s_n)
        self._rescale_lead()
        k.spec_power(self.WPHI, self.HPHI, self.SPHI, N, False)
        if self.compute_error:
            self._error(self.counter)
        self.counter += 1
        # HM (:335-347; :741-763)
        self._acc_terms()
        k.gemm_view(self.WM, self.work_acc, self.C_hm, self.R, 2 * self.nch * ldn, F, transA=True)
        k.simm_update_rows(self.HM, self.C_hm, self.nch, ldn, self.b2, om,
                           0.0 if self.stereo else EPS, self.R, N)
        self._model_acc()
        self.counter += 1
        # HGAMMA (:351-372; :776-819)
        if self.update_hgamma:
            self._lead_terms(True)
            k.gemm_view(self.work_lead[:, :ldn], self.HPHI, self.tn, F, self.K, ldn, transB=True,
                        workspace=self.splitk_ws)
            k.gemm_view(self.work_lead[:, ldn:], self.HPHI, self.td, F, self.K, ldn, transB=True,
                        workspace=self.splitk_ws)
            k.simm_hgamma_update(self.HGAMMA, self.WGAMMA, self.tn, self.td, F, self.P, self.K, om,
                                 self.s_k)
            k.simm_hphi_normalise(self.HPHI, self.K, self.s_k, N, self.s_n)
            self._rescale_lead()
            k.small_matmul(self.WGAMMA, self.HGAMMA, self.WPHI)
            k.spec_power(self.WPHI, self.HPHI, self.SPHI, N, False)
            self.counter += 1
        # WM (:376-393; :829-866)
        self._acc_products()
        k.simm_wm_update(self.WM, self.R, self.D, self.nch, self.b2, not self.stereo, om, F,
                         self.s_r)
        if self.stereo or self.R == 1:
            k.simm_scale_rows(self.HM, self.R, N, self.s_r)
        else:  # mono quirk (:388): R == N, sumWM scales the COLUMNS of HM
            k.simm_scale_columns(self.HM, self.R, N, self.s_r)
        self._model_acc()
        self.counter += 1
        if not self.stereo:
            return
        # alpha (:869-896)
        k.simm_alpha_update(self.SX, self.SM, self.SF0, self.SPHI, F, N, ldn, om, self.red_ws,
                            self.alpha, self.a2)
        self.counter += 1
        # beta (:909-941)
        self._acc_products()
        k.simm_beta_update(self.WM, self.R, self.D, F, om, self.beta, self.b2)
        self._model_acc()
        self.counter += 1

    # -- results (float64 NumPy, the reference's types) ----------------------------------------
    def _down(self, t, rows, cols):
        return t[:rows, :cols].to("cpu").numpy().astype(np.float64)

    def results(self):
        out = dict(HGAMMA=self._down(self.HGAMMA, self.P, self.K),
                   HPHI=self._down(self.HPHI, self.K, self.N),
                   HF0=self._down(self.HF0, self.NF0, self.N),
                   HM=self._down(self.HM, self.R, self.N),
                   WM=self._down(self.WM, self.F, self.R),
                   recoError=self.reco.to("cpu").numpy().copy())
        if self.stereo:
            al = self.alpha.to("cpu").numpy()
            be = self.beta.to("cpu").numpy()
            out.update(alphaR=float(al[0]), alphaL=float(al[1]), betaR=be[0, :self.R].copy(),
                       betaL=be[1, :self.R].copy())
        return out
