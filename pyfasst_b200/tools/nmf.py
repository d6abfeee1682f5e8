"""IS-NMF initialisers on the GPU, drop-in for pyfasst/tools/nmf.py: `NMF_decomposition` (:24-62)
and `NMF_decomp_init` (:64-159) -- the step before the GEM loop in every realistic FASST workflow
(FASST.initialize_all_spec_comps_with_NMF*, audioModel.py:2091-2222).

SX ~ W H under the Itakura-Saito divergence, multiplicative updates
    W *= ((SX / hat^2) H^T) / ((1 / hat) H^T),  columns of W normalised to sum one (H absorbs it),
    H *= (W^T (SX / hat^2)) / (W^T (1 / hat)),
all clamps at eps = 1e-10.  float32 on the device: the four contractions per iteration are
tensor-core GEMMs (csrc/gemm_tc.cu, split-K for the ones over the frame axis), the rest
bandwidth-bound kernels (csrc/simm.cu).  NumPy in / NumPy out; no CPU fallback.
"""
import numpy as np

eps = 1e-10


def _ru4(n):
    return (int(n) + 3) // 4 * 4


class _IsNmf(object):
    """Device state and one iteration of the IS-NMF updates."""

    def __init__(self, kernels, SX, W, H):
        import torch
        self.torch = torch
        self.k = k = kernels
        dev = k.device
        F = SX.shape[0]
        N = H.shape[1]
        K = W.shape[1]
        self.F, self.N, self.K = F, N, K
        self.ldn, self.ldk = _ru4(N), _ru4(K)

        def up(a, rows, cols):
            buf = np.zeros((rows, cols), dtype=np.float32)
            buf[:a.shape[0], :a.shape[1]] = a
            return torch.from_numpy(buf).to(dev)
        # SX: NumPy [F, N], or a device plane [F, ldn] float32 (padding zero) already in HBM
        self.SX = SX if hasattr(SX, "data_ptr") else up(np.asarray(SX), F, self.ldn)
        self.W = up(np.asarray(W, dtype=np.float64), F, self.ldk)
        self.H = up(np.asarray(H, dtype=np.float64), self.ldk, self.ldn)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.hat = z(F, self.ldn)
        self.work = z(F, 2 * self.ldn)       # (T | I)
        self.D = z(2, F, self.ldk)           # T H^T, I H^T
        self.C = z(self.ldk, 2 * self.ldn)   # W^T (T | I)
        self.s = z(self.ldk)
        self.ws = z(max(k.gemm_splitk_workspace_bytes(F, K, self.ldn) // 4, 4))

    def _terms(self):
        k, F, N, K = self.k, self.F, self.N, self.K
        k.gemm_view(self.W, self.H, self.hat, F, N, self.ldk)           # hat = W H
        k.nmf_is_terms(self.hat, self.SX, self.work, eps, F, N, self.ldn)

    def update_W(self):
        """(ref: nmf.py:39-51, :131-146)"""
        k, F, N, K, ldn = self.k, self.F, self.N, self.K, self.ldn
        self._terms()
        for q in range(2):
            k.gemm_view(self.work[:, q * ldn:(q + 1) * ldn], self.H, self.D[q], F, K, ldn,
                        transB=True, workspace=self.ws)
        k.nmf_w_update(self.W, K, self.D, eps, F, self.s)
        k.simm_scale_rows(self.H, K, N, self.s)                          # H *= vstack(sumW)

    def update_H(self):
        """(ref: nmf.py:53-60, :148-157)"""
        k, F, N, K, ldn = self.k, self.F, self.N, self.K, self.ldn
        self._terms()
        k.gemm_view(self.W, self.work, self.C, K, 2 * ldn, F, transA=True)
        k.nmf_update_rows(self.H, self.C, ldn, eps, K, N)

    def results(self):
        W = self.W[:, :self.K].to("cpu").numpy().astype(np.float64)
        H = self.H[:self.K, :self.N].to("cpu").numpy().astype(np.float64)
        return W, H


def _shape(SX, nframes=None):
    """(freqs, nframes); a device plane [F, ldn] carries padding frames: its caller says how
    many frames are real."""
    freqs, n = SX.shape
    return freqs, (n if nframes is None else int(nframes))


def _kernels(kernels):
    if kernels is not None:
        return kernels
    from ..tftransforms.stft import default_kernels
    return default_kernels()


def NMF_decomposition(SX, nbComps=10, niter=10, verbose=0, kernels=None, nframes=None):
    """W, H = NMF_decomposition(SX, nbComps, niter)   (ref: nmf.py:24-62).  Random initial
    W, H = randn^2 drawn from the global NumPy RNG in the reference's order."""
    freqs, nframes = _shape(SX, nframes)
    W = np.random.randn(freqs, nbComps) ** 2
    H = np.random.randn(nbComps, nframes) ** 2
    W /= W.sum(axis=0)
    eng = _IsNmf(_kernels(kernels), SX, W, H)
    for i in range(niter):
        if verbose:
            print("    NMF iteration %d out of %d" % (i + 1, niter))
        eng.update_W()
        eng.update_H()
    return eng.results()


def NMF_decomp_init(SX, nbComps=10, niter=10, verbose=0, Winit=None, Hinit=None, updateW=True,
                    updateH=True, kernels=None, nframes=None):
    """W, H = NMF_decomp_init(SX, nbComps, niter, Winit=, Hinit=, updateW=, updateH=)
    (ref: nmf.py:64-159).  Hinit may be [nbComps, nframes] or its transpose; anything else raises
    AttributeError like the reference."""
    freqs, nframes = _shape(SX, nframes)
    if Winit is None or (np.shape(Winit) != (freqs, nbComps)):
        W = np.random.randn(freqs, nbComps) ** 2
    else:
        W = np.copy(Winit)
    if Hinit is not None:
        if np.shape(Hinit) == (nbComps, nframes):
            H = np.copy(Hinit)
        elif np.shape(Hinit) == (nframes, nbComps):
            H = np.copy(np.asarray(Hinit).T)
        else:
            raise AttributeError('Hinit not in the right shape.')
    else:
        H = (np.random.randn(nframes, nbComps) ** 2).T
    if updateW:
        W = W / W.sum(axis=0)
    eng = _IsNmf(_kernels(kernels), SX, W, H)
    for i in range(niter):
        if verbose:
            print("    NMF iteration %d out of %d" % (i + 1, niter))
        if updateW:
            eng.update_W()
        if updateH:
            eng.update_H()
    return eng.results()
