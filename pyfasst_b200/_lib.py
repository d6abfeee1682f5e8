"""ctypes binding of libpyfasst_b200.so (the C ABI declared in include/pyfasst_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  `CudaKernels` is the object the GEM engine talks to; every
method takes torch CUDA tensors (device memory + stream plumbing only) and
enqueues hand-written sm_100a kernels on torch's current stream.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PYFASST_B200_LIB: another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("PYFASST_B200_LIB") or os.path.join(HERE, "libpyfasst_b200.so")

PF_F32, PF_F64 = 0, 1
PF_FLAG_SINGULAR, PF_FLAG_TW_RESTART = 1, 2
ABI_VERSION = 10

c_int, c_i64, c_dbl, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
c_ip = ctypes.POINTER(ctypes.c_int)

# name -> argument ctypes (return type is int unless listed in _RESTYPE)
SIGNATURES = {
    "pf_last_error": [],
    "pf_estep_timing": [c_int],
    "pf_estep_timing_read": [ctypes.POINTER(ctypes.c_double), c_ip],
    "pf_copy_2d": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp],
    "pf_abi_version": [],
    "pf_launch_count": [],
    "pf_set_device": [c_int],
    "pf_stft": [c_vp, c_int, c_dbl, c_int, c_i64, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_vp,
                c_i64, c_i64, c_i64, c_vp, c_int, c_vp],
    "pf_pcm_peak": [c_vp, c_int, c_i64, c_vp, c_vp],
    "pf_istft": [c_vp, c_int, c_int, c_i64, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_i64,
                 c_vp, c_dbl, c_i64, c_int, c_int, c_vp],
    "pf_wiener_stereo": [c_vp, c_vp, c_vp, c_ip, c_int, c_int, c_vp, c_ip, c_int, c_int, c_i64,
                         c_i64, c_vp, c_vp, c_i64, c_int, c_vp],
    "pf_estep_plan": [c_int, c_i64, c_int, ctypes.POINTER(c_i64), c_ip, ctypes.POINTER(c_i64),
                      c_int],
    "pf_estep_stereo": [c_vp, c_vp, c_vp, c_ip, c_int, c_int, c_vp, c_int, c_i64, c_i64, c_vp,
                        c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_vp],
    "pf_estep_stereo_inst": [c_vp, c_vp, c_vp, c_ip, c_int, c_int, c_vp, c_int, c_i64, c_i64, c_vp,
                        c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_vp],
    "pf_estep_multi_plan": [c_int, c_int, c_int, c_i64, ctypes.POINTER(c_i64), c_ip,
                            ctypes.POINTER(c_i64)],
    "pf_estep_multi": [c_vp, c_vp, c_vp, c_ip, c_int, c_int, c_int, c_vp, c_int, c_i64, c_i64, c_vp,
                       c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_vp],
    "pf_wiener_multi": [c_vp, c_vp, c_vp, c_ip, c_int, c_int, c_int, c_vp, c_ip, c_int, c_int, c_i64,
                        c_i64, c_vp, c_vp, c_i64, c_int, c_vp],
    "pf_mix_inst_stats": [c_vp, c_vp, c_vp, c_ip, c_int, c_ip, c_int, c_int, c_int, c_int, c_vp,
                          c_vp],
    "pf_mix_inst_solve": [c_vp, c_dbl, c_ip, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "pf_mix_conv_solve": [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "pf_spec_power": [c_vp, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_i64, c_int, c_int,
                      c_vp],
    "pf_small_matmul": [c_vp, c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    "pf_nmf_fb_plan": [c_int, c_int, c_i64, c_int, ctypes.POINTER(c_i64), c_ip],
    "pf_nmf_fb_contract": [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_i64, c_vp, c_vp,
                           c_i64, c_int, c_int, c_vp],
    "pf_nmf_tw_plan": [c_int, c_int, c_i64, c_int, c_ip, c_ip],
    "pf_nmf_tw_contract": [c_vp, c_vp, c_i64, c_vp, c_int, c_vp, c_i64, c_int, c_int, c_i64, c_vp,
                           c_vp, c_i64, c_int, c_int, c_vp, c_int, c_vp],
    "pf_sum_splits": [c_vp, c_int, c_i64, c_vp, c_vp],
    "pf_tw_pack_chunks": [c_vp, c_vp, c_int, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_int, c_vp],
    "pf_mult_update": [c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_i64, c_dbl, c_int, c_vp],
    "pf_mult_update_splits": [c_vp, c_i64, c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_i64, c_dbl,
                              c_int, c_vp],
    "pf_spat_energy": [c_vp, c_ip, c_int, c_int, c_int, c_int, c_vp, c_vp],
    "pf_spat_scale": [c_vp, c_ip, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "pf_fb_scale_colmax": [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_vp],
    "pf_fw_renorm": [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_vp],
    "pf_scale_matrix": [c_vp, c_i64, c_int, c_i64, c_vp, c_int, c_int, c_vp, c_int, c_vp],
    "pf_check_totals": [c_vp, c_int, c_dbl, c_vp, c_vp, c_vp, c_vp],
    "pf_gemm_tf32x3": [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int, c_int,
                       c_vp],
    "pf_gemm_splitk_plan": [c_int, c_int, c_int, c_ip, ctypes.POINTER(c_i64)],
    "pf_gemm_tf32x3_splitk": [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int,
                              c_int, c_vp, c_i64, c_vp],
    "pf_simm_lead_terms": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_i64, c_i64,
                           c_vp],
    "pf_simm_acc_terms": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_i64,
                          c_vp],
    "pf_simm_hat": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_vp],
    "pf_simm_reduce_workspace_bytes": [],
    "pf_simm_is_divergence": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_vp, c_vp,
                              c_vp],
    "pf_simm_alpha_update": [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_dbl, c_vp, c_vp, c_vp,
                             c_vp],
    "pf_simm_update_rows": [c_vp, c_i64, c_vp, c_i64, c_int, c_i64, c_vp, c_int, c_dbl, c_dbl,
                            c_int, c_i64, c_vp],
    "pf_simm_hphi_normalise": [c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_vp],
    "pf_simm_scale_columns": [c_vp, c_i64, c_int, c_i64, c_vp, c_vp],
    "pf_simm_scale_rows": [c_vp, c_i64, c_int, c_i64, c_vp, c_vp],
    "pf_simm_hgamma_update": [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int,
                              c_dbl, c_vp, c_vp],
    "pf_simm_wm_update": [c_vp, c_int, c_int, c_vp, c_int, c_vp, c_int, c_dbl, c_int, c_vp, c_vp],
    "pf_simm_beta_update": [c_vp, c_int, c_int, c_vp, c_int, c_dbl, c_vp, c_vp, c_vp],
    "pf_simm_power": [c_vp, c_i64, c_vp, c_int, c_int, c_i64, c_i64, c_vp],
    "pf_simm_masks": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_int, c_int, c_i64,
                      c_i64, c_vp],
    "pf_nmf_is_terms": [c_vp, c_vp, c_vp, c_dbl, c_int, c_i64, c_i64, c_vp],
    "pf_nmf_update_rows": [c_vp, c_i64, c_vp, c_i64, c_i64, c_dbl, c_int, c_i64, c_vp],
    "pf_nmf_w_update": [c_vp, c_int, c_int, c_vp, c_dbl, c_int, c_vp, c_vp],
    "pf_mono_power": [c_vp, c_i64, c_int, c_vp, c_int, c_i64, c_i64, c_vp],
    "pf_simm_wm_scaled": [c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp],
    "pf_viterbi_workspace_bytes": [c_int, c_i64],
    "pf_viterbi": [c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_i64, c_vp, c_vp],
    "pf_gem_ratio_planes": [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_dbl, c_int,
                            c_vp],
    "pf_corr_planes": [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_int, c_vp],
    "pf_row_sums": [c_vp, c_i64, c_int, c_i64, c_vp, c_int, c_vp],
    "pf_apply_filter": [c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_i64, c_int, c_vp],
    "pf_mul_planes": [c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_int, c_vp],
    "pf_mult_update_same": [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_i64, c_dbl, c_int, c_vp],
    "pf_sparsity_reweigh": [c_vp, c_i64, c_int, c_i64, c_int, c_dbl, c_dbl, c_vp, c_vp, c_int, c_vp],
    "pf_overlap_norm": [c_vp, c_int, c_int, c_i64, c_vp, c_vp],
    "pf_wf0_combs": [c_vp, c_vp, c_vp, c_int, c_int, c_dbl, c_dbl, c_i64, c_i64, c_vp, c_int, c_int,
                     c_int, c_vp, c_vp],
    "pf_tc_selftest": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    "pf_noise_anneal": [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "pf_ll_reduce": [c_vp, c_int, c_vp, c_vp],
    "pf_ll_store": [c_vp, c_dbl, c_vp, c_vp, c_int, c_vp],
}
_RESTYPE = {"pf_last_error": ctypes.c_char_p, "pf_launch_count": ctypes.c_ulonglong,
            "pf_simm_reduce_workspace_bytes": ctypes.c_int64,
            "pf_viterbi_workspace_bytes": ctypes.c_int64}

_lib = None


class KernelError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen the shared library and declare every prototype.  Raises ImportError if
    it has not been built (python -m pyfasst_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise ImportError(
            "pyfasst_b200: %s not found -- build it with `python -m pyfasst_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, c_int)
    if lib.pf_abi_version() != ABI_VERSION:
        raise ImportError("pyfasst_b200: ABI version %d, expected %d -- rebuild the library"
                          % (lib.pf_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def _check(rc, lib):
    if rc == 0:
        return
    msg = lib.pf_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotImplementedError(msg)
    raise KernelError(msg)


def _iarr(values):
    arr = (ctypes.c_int * max(len(values), 1))(*[int(v) for v in values])
    return arr


class CudaKernels(object):
    """The kernels of the C ABI, called with torch CUDA tensors."""

    name = "cuda"

    def __init__(self, device=None):
        import torch
        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("pyfasst_b200 needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()
                                   if device is None else device)
        _check(self.lib.pf_set_device(self.device.index), self.lib)
        self._dev_index = self.device.index
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)

    # -- helpers ----------------------------------------------------------------
    def estep_timing(self, enable):
        _check(self.lib.pf_estep_timing(1 if enable else 0), self.lib)

    def estep_timing_read(self):
        """(summed ms, launches) of the per-bin E-step kernels bracketed since estep_timing(True)."""
        ms, n = ctypes.c_double(), c_int()
        _check(self.lib.pf_estep_timing_read(ctypes.byref(ms), ctypes.byref(n)), self.lib)
        return float(ms.value), int(n.value)

    def copy_cols_to_device(self, host, lo, hi, dev_t):
        """dev_t [rows, hi - lo] (contiguous, dtype of `host`) <- host[:, lo:hi] of a C-contiguous
        2-D host array: one strided DMA, no host-side temporary."""
        rows, isz = host.shape[0], host.itemsize
        assert host.flags.c_contiguous and dev_t.is_contiguous() and dev_t.element_size() == isz
        _check(self.lib.pf_copy_2d(dev_t.data_ptr(), (hi - lo) * isz,
                                   host.ctypes.data + lo * isz, host.shape[1] * isz,
                                   (hi - lo) * isz, rows, self._stream()), self.lib)

    def copy_cols_to_host(self, dev_t, host, lo, hi):
        """host[:, lo:hi] <- dev_t [rows, hi - lo] (contiguous, dtype of `host`)."""
        rows, isz = host.shape[0], host.itemsize
        assert host.flags.c_contiguous and host.flags.writeable and dev_t.is_contiguous() \
            and dev_t.element_size() == isz
        _check(self.lib.pf_copy_2d(host.ctypes.data + lo * isz, host.shape[1] * isz,
                                   dev_t.data_ptr(), (hi - lo) * isz, (hi - lo) * isz, rows,
                                   self._stream()), self.lib)

    def _p(self, t):
        if t is None:
            return None
        # (2-D views with contiguous rows are fine: every matrix argument carries its stride)
        assert t.is_cuda and (t.is_contiguous() or (t.dim() == 2 and t.stride(1) == 1)), \
            "kernel arguments must be CUDA tensors with contiguous rows"
        return t.data_ptr()

    def _stream(self):
        # (the raw handle of torch's current stream: ~0.3 us instead of ~4 us through
        # torch.cuda.current_stream() -- called once per kernel launch, ~60 times per GEM iteration)
        raw = self._raw_stream
        if raw is not None:
            return raw(self._dev_index)
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def dtype_code(self, t):
        if t.dtype == self.torch.float32:
            return PF_F32
        if t.dtype == self.torch.float64:
            return PF_F64
        raise ValueError("unsupported plane dtype %s" % t.dtype)

    def launch_count(self):
        return int(self.lib.pf_launch_count())

    # -- K1 / K6 -------------------------------------------------------------------
    def stft(self, pcm, window, hop, nfft, X, N, psd_sum, pcm_div=1.0, sample0=0, L_total=None,
             frame0=0):
        """pcm: float64 [nch, L] (planar) or int16 / int32 / float32 [L, nch] (interleaved);
        a window [sample0, sample0+L) of a signal of L_total samples.  X receives the frames
        [frame0, frame0+N)."""
        torch = self.torch
        if pcm.dtype == torch.float64:
            fmt, (nch, L) = 0, pcm.shape
        else:
            fmt = {torch.int16: 1, torch.int32: 2, torch.float32: 3}[pcm.dtype]
            L, nch = pcm.shape
        F, ld = X.shape[1], X.shape[2]
        assert X.shape[0] == 2 * nch and F == nfft // 2 + 1
        _check(self.lib.pf_stft(self._p(pcm), fmt, float(pcm_div), nch, L, sample0,
                                L if L_total is None else L_total, self._p(window),
                                window.numel(), hop, nfft, self._p(X), frame0, N, ld,
                                self._p(psd_sum), self.dtype_code(X), self._stream()), self.lib)

    def pcm_peak(self, pcm, peak):
        """peak[0] = max |x| of a device PCM tensor (any layout; the whole tensor is scanned)."""
        torch = self.torch
        fmt = {torch.float64: 0, torch.int16: 1, torch.int32: 2, torch.float32: 3}[pcm.dtype]
        _check(self.lib.pf_pcm_peak(self._p(pcm), fmt, pcm.numel(), self._p(peak),
                                    self._stream()), self.lib)

    def istft(self, Y, N, synth, norm, hop, nfft, out, pcm, maxdata, drop=None, pcm_round=False):
        """drop: overlap-added samples skipped at the start (default wlen/2, the FASST
        convention); pcm_round: round half to even instead of truncating."""
        nsig = Y.shape[0] // 2
        F, ld = Y.shape[1], Y.shape[2]
        assert out.shape[0] == nsig
        drop = synth.numel() // 2 if drop is None else int(drop)
        _check(self.lib.pf_istft(self._p(Y), nsig, F, N, ld, self._p(synth), self._p(norm),
                                 synth.numel(), hop, nfft, self._p(out), out.shape[1],
                                 self._p(pcm), float(maxdata), drop, int(pcm_round),
                                 self.dtype_code(Y), self._stream()), self.lib)

    def wiener_stereo(self, X, V, A, src_of_sub, noise, group_of_src, ngroups, N, Y, workspace):
        J, F, ld = V.shape
        R = A.shape[0]
        _check(self.lib.pf_wiener_stereo(self._p(X), self._p(V), self._p(A), _iarr(src_of_sub), R,
                                         J, self._p(noise), _iarr(group_of_src), ngroups, F, N,
                                         ld, self._p(Y), self._p(workspace),
                                         workspace.numel() * workspace.element_size(),
                                         self.dtype_code(V), self._stream()), self.lib)

    # -- K2 ---------------------------------------------------------------------------
    def estep_workspace_bytes(self, J, F, N, dtype_code):
        chunk, nsplit, nbytes = c_i64(), c_int(), c_i64()
        _check(self.lib.pf_estep_plan(J, N, dtype_code, ctypes.byref(chunk), ctypes.byref(nsplit),
                                      ctypes.byref(nbytes), F), self.lib)
        return nbytes.value

    def estep_stereo(self, X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace,
                     N_norm=0):
        J, F, ld = V.shape
        R = A.shape[0]
        code = self.dtype_code(V)
        _check(self.lib.pf_estep_stereo(self._p(X), self._p(V), self._p(A), _iarr(src_of_sub), R,
                                        J, self._p(noise), F, N, ld, self._p(hatW), self._p(Rss),
                                        self._p(Rxs), self._p(ll_f), self._p(workspace),
                                        workspace.numel() * workspace.element_size(),
                                        int(N_norm), code, self._stream()), self.lib)

    def estep_stereo_inst(self, X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace,
                          N_norm=0):
        """estep_stereo for REAL mixing vectors (instantaneous mixing): the imaginary parts of
        hat_Rss / hat_Rxs are not formed (returned as zero)."""
        J, F, ld = V.shape
        R = A.shape[0]
        code = self.dtype_code(V)
        _check(self.lib.pf_estep_stereo_inst(self._p(X), self._p(V), self._p(A), _iarr(src_of_sub),
                                             R, J, self._p(noise), F, N, ld, self._p(hatW),
                                             self._p(Rss), self._p(Rxs), self._p(ll_f),
                                             self._p(workspace),
                                             workspace.numel() * workspace.element_size(),
                                             int(N_norm), code, self._stream()), self.lib)

    # -- K2 / K6 for I = 2..4 channels (X: [2 I, F, ld], A: [R, I, F], Rxs: [F, I, R]) --------
    def estep_multi_workspace_bytes(self, I, J, F, N):
        chunk, nsplit, nbytes = c_i64(), c_int(), c_i64()
        _check(self.lib.pf_estep_multi_plan(I, J, F, N, ctypes.byref(chunk), ctypes.byref(nsplit),
                                            ctypes.byref(nbytes)), self.lib)
        return nbytes.value

    def estep_multi(self, X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace,
                    N_norm=0):
        J, F, ld = V.shape
        R, I = A.shape[0], A.shape[1]
        assert X.shape[0] == 2 * I
        _check(self.lib.pf_estep_multi(self._p(X), self._p(V), self._p(A), _iarr(src_of_sub), R, J,
                                       I, self._p(noise), F, N, ld, self._p(hatW), self._p(Rss),
                                       self._p(Rxs), self._p(ll_f), self._p(workspace),
                                       workspace.numel() * workspace.element_size(), int(N_norm),
                                       self.dtype_code(V), self._stream()), self.lib)

    def wiener_multi(self, X, V, A, src_of_sub, noise, group_of_src, ngroups, N, Y, workspace):
        J, F, ld = V.shape
        R, I = A.shape[0], A.shape[1]
        _check(self.lib.pf_wiener_multi(self._p(X), self._p(V), self._p(A), _iarr(src_of_sub), R, J,
                                        I, self._p(noise), _iarr(group_of_src), ngroups, F, N, ld,
                                        self._p(Y), self._p(workspace),
                                        workspace.numel() * workspace.element_size(),
                                        self.dtype_code(V), self._stream()), self.lib)

    # -- K3 ---------------------------------------------------------------------------
    def mix_inst_stats(self, Rss, Rxs, A, upd, oth, stats):
        R, I, F = A.shape
        _check(self.lib.pf_mix_inst_stats(self._p(Rss), self._p(Rxs), self._p(A), _iarr(upd),
                                          len(upd), _iarr(oth), len(oth), R, I, F,
                                          self._p(stats), self._stream()), self.lib)

    def mix_inst_solve(self, stats, F_total, upd, A, flags):
        R, I, F = A.shape
        _check(self.lib.pf_mix_inst_solve(self._p(stats), float(F_total), _iarr(upd), len(upd), I,
                                          F, self._p(A), self._p(flags), self._stream()), self.lib)

    def mix_conv_solve(self, Rss, Rxs, A, flags):
        R, I, F = A.shape
        _check(self.lib.pf_mix_conv_solve(self._p(Rss), self._p(Rxs), R, I, F, self._p(A),
                                          self._p(flags), self._stream()), self.lib)

    # -- K4 ---------------------------------------------------------------------------
    def spec_power(self, W, H, V, N, accumulate):
        F, K = W.shape
        _check(self.lib.pf_spec_power(self._p(W), W.stride(0), self._p(H), H.stride(0),
                                      self._p(V), V.stride(0), F, K, N, int(accumulate),
                                      self.dtype_code(V), self._stream()), self.lib)

    def small_matmul(self, A, B, C):
        M, K = A.shape
        Nc = B.shape[1]
        _check(self.lib.pf_small_matmul(self._p(A), A.stride(0), self._p(B), B.stride(0),
                                        self._p(C), C.stride(0), M, K, Nc, self.dtype_code(C),
                                        self._stream()), self.lib)

    def fb_plan(self, F, K, N, dtype_code):
        chunk, nsplit = c_i64(), c_int()
        _check(self.lib.pf_nmf_fb_plan(F, K, N, dtype_code, ctypes.byref(chunk),
                                       ctypes.byref(nsplit)), self.lib)
        return chunk.value, nsplit.value

    def fb_contract(self, hatW, P, O, G, N, num_partial, den_partial, chunk, nsplit):
        F, ld = hatW.shape
        K = G.shape[0]
        _check(self.lib.pf_nmf_fb_contract(self._p(hatW), self._p(P), self._p(O), ld, self._p(G),
                                           G.stride(0), F, K, N, self._p(num_partial),
                                           self._p(den_partial), chunk, nsplit,
                                           self.dtype_code(hatW), self._stream()), self.lib)

    def tw_plan(self, F, K, N, dtype_code):
        fchunk, fsplit = c_int(), c_int()
        _check(self.lib.pf_nmf_tw_plan(F, K, N, dtype_code, ctypes.byref(fchunk),
                                       ctypes.byref(fsplit)), self.lib)
        return fchunk.value, fsplit.value

    def tw_contract(self, hatW, O, W, H, N, num_partial, den_partial, fchunk, fsplit,
                    scratch=None):
        F, ld = hatW.shape
        K = W.shape[1]
        _check(self.lib.pf_nmf_tw_contract(self._p(hatW), self._p(O), ld, self._p(W), W.stride(0),
                                           self._p(H), H.stride(0), F, K, N,
                                           self._p(num_partial), self._p(den_partial),
                                           num_partial.shape[-1], fchunk, fsplit,
                                           self._p(scratch), self.dtype_code(hatW),
                                           self._stream()), self.lib)

    def tw_pack_chunks(self, num_partial, den_partial, out, world):
        """out [world, 2, Kmax, c] (plane type) <- the split sums of num/den_partial [nsplit, K, ld]."""
        nsplit, K, ld = num_partial.shape
        _check(self.lib.pf_tw_pack_chunks(self._p(num_partial), self._p(den_partial), nsplit,
                                          num_partial.stride(0), K, ld, world, out.shape[2],
                                          self._p(out), self.dtype_code(out), self._stream()),
               self.lib)

    def sum_splits(self, parts, out):
        nsplit = parts.shape[0]
        _check(self.lib.pf_sum_splits(self._p(parts), nsplit, out.numel(), self._p(out),
                                      self._stream()), self.lib)

    def mult_update(self, theta, num, den, rows, cols, omega):
        _check(self.lib.pf_mult_update(self._p(theta), theta.stride(0), self._p(num), self._p(den),
                                       num.stride(0), rows, cols, float(omega),
                                       self.dtype_code(theta), self._stream()), self.lib)

    def mult_update_splits(self, theta, num_partial, den_partial, rows, cols, omega):
        """num/den_partial: [nsplit, rows_alloc, ld] split partial sums."""
        nsplit = num_partial.shape[0]
        _check(self.lib.pf_mult_update_splits(self._p(theta), theta.stride(0),
                                              self._p(num_partial), self._p(den_partial), nsplit,
                                              num_partial.stride(0), num_partial.stride(1), rows,
                                              cols, float(omega), self.dtype_code(theta),
                                              self._stream()), self.lib)

    def gemm(self, A, B, C, M, N, K, transA=False, transB=False):
        """C[M,N] = op(A) op(B) on the tensor cores (float32, 3xTF32); see pf_gemm_tf32x3."""
        _check(self.lib.pf_gemm_tf32x3(self._p(A), A.stride(0), int(transA), self._p(B),
                                       B.stride(0), int(transB), self._p(C), C.stride(0), M, N, K,
                                       self._stream()), self.lib)

    def gemm_splitk_workspace_bytes(self, M, N, K):
        ks, nbytes = c_int(), c_i64()
        _check(self.lib.pf_gemm_splitk_plan(M, N, K, ctypes.byref(ks), ctypes.byref(nbytes)),
               self.lib)
        return nbytes.value

    def gemm_splitk(self, A, B, C, M, N, K, workspace, transA=False, transB=False):
        """The same product with the contraction split over CTAs (long K, few output tiles)."""
        _check(self.lib.pf_gemm_tf32x3_splitk(self._p(A), A.stride(0), int(transA), self._p(B),
                                              B.stride(0), int(transB), self._p(C), C.stride(0),
                                              M, N, K, self._p(workspace),
                                              workspace.numel() * workspace.element_size(),
                                              self._stream()), self.lib)

    # -- K7 / K8: SIMM (float32; arguments may be row-major VIEWS with unit column stride) --
    def _pv(self, t):
        if t is None:
            return None
        assert t.is_cuda and t.dtype in (self.torch.float32, self.torch.float64)
        assert t.dim() == 1 or t.is_contiguous() or t.stride(-1) == 1, "rows must be contiguous"
        return t.data_ptr()

    def _call(self, name, *args):
        _check(getattr(self.lib, name)(*args, self._stream()), self.lib)

    def gemm_view(self, A, B, C, M, N, K, transA=False, transB=False, workspace=None):
        """C = op(A) op(B) on row-major views; split-K when a workspace is given."""
        if workspace is None:
            self._call("pf_gemm_tf32x3", self._pv(A), A.stride(0), int(transA), self._pv(B),
                       B.stride(0), int(transB), self._pv(C), C.stride(0), M, N, K)
        else:
            self._call("pf_gemm_tf32x3_splitk", self._pv(A), A.stride(0), int(transA),
                       self._pv(B), B.stride(0), int(transB), self._pv(C), C.stride(0), M, N, K,
                       self._pv(workspace), workspace.numel() * workspace.element_size())

    def simm_reduce_workspace_bytes(self):
        return int(self.lib.pf_simm_reduce_workspace_bytes())

    def simm_lead_terms(self, SM, SF0, SPHI, SX, a2, other_is_sf0, out, nch, F, N, ldn):
        self._call("pf_simm_lead_terms", self._pv(SM), self._pv(SF0), self._pv(SPHI),
                   self._pv(SX), self._pv(a2), int(other_is_sf0), self._pv(out), nch, F, N, ldn)

    def simm_acc_terms(self, SM, SF0, SPHI, SX, a2, out, nch, sq_clamp, F, N, ldn):
        self._call("pf_simm_acc_terms", self._pv(SM), self._pv(SF0), self._pv(SPHI),
                   self._pv(SX), self._pv(a2), self._pv(out), nch, int(sq_clamp), F, N, ldn)

    def simm_hat(self, SM, SF0, SPHI, a2, hat, nch, F, N, ldn):
        self._call("pf_simm_hat", self._pv(SM), self._pv(SF0), self._pv(SPHI), self._pv(a2),
                   self._pv(hat), nch, F, N, ldn)

    def simm_is_divergence(self, SX, SM, SF0, SPHI, a2, nch, F, N, ldn, workspace, out):
        self._call("pf_simm_is_divergence", self._pv(SX), self._pv(SM), self._pv(SF0),
                   self._pv(SPHI), self._pv(a2), nch, F, N, ldn, self._pv(workspace),
                   self._pv(out))

    def simm_alpha_update(self, SX, SM, SF0, SPHI, F, N, ldn, omega, workspace, alpha, a2):
        self._call("pf_simm_alpha_update", self._pv(SX), self._pv(SM), self._pv(SF0),
                   self._pv(SPHI), F, N, ldn, float(omega), self._pv(workspace), self._pv(alpha),
                   self._pv(a2))

    def simm_update_rows(self, theta, C, nch, ldn, w, omega, floor_value, rows, N):
        self._call("pf_simm_update_rows", self._pv(theta), theta.stride(0), self._pv(C),
                   C.stride(0), nch, ldn, self._pv(w), 0 if w is None else w.stride(0),
                   float(omega), float(floor_value), rows, N)

    def simm_hphi_normalise(self, HPHI, K, rowscale, N, s_out):
        self._call("pf_simm_hphi_normalise", self._pv(HPHI), HPHI.stride(0), K,
                   self._pv(rowscale), N, self._pv(s_out))

    def simm_scale_columns(self, P, rows, N, s):
        self._call("pf_simm_scale_columns", self._pv(P), P.stride(0), rows, N, self._pv(s))

    def simm_scale_rows(self, P, rows, N, s):
        self._call("pf_simm_scale_rows", self._pv(P), P.stride(0), rows, N, self._pv(s))

    def simm_hgamma_update(self, HGAMMA, WGAMMA, tn, td, F, P, K, omega, s_out):
        self._call("pf_simm_hgamma_update", self._pv(HGAMMA), HGAMMA.stride(0), self._pv(WGAMMA),
                   WGAMMA.stride(0), self._pv(tn), self._pv(td), tn.stride(0), F, P, K,
                   float(omega), self._pv(s_out))

    def simm_wm_update(self, WM, R, D, nch, b2, clamp_den, omega, F, s_out):
        self._call("pf_simm_wm_update", self._pv(WM), WM.stride(0), R, self._pv(D), nch,
                   self._pv(b2), int(clamp_den), float(omega), F, self._pv(s_out))

    def simm_beta_update(self, WM, R, D, F, omega, beta, b2):
        self._call("pf_simm_beta_update", self._pv(WM), WM.stride(0), R, self._pv(D), F,
                   float(omega), self._pv(beta), self._pv(b2))

    def simm_power(self, X, SX, nch, F, N, ldn):
        self._call("pf_simm_power", self._pv(X), X.stride(1), self._pv(SX), nch, F, N, ldn)

    def simm_masks(self, SM, SF0, SPHI, a2, X, Y, eps_hat, nch, F, N, ldn):
        self._call("pf_simm_masks", self._pv(SM), self._pv(SF0), self._pv(SPHI), self._pv(a2),
                   self._pv(X), X.stride(1), self._pv(Y), float(eps_hat), nch, F, N, ldn)

    # -- Viterbi ------------------------------------------------------------------------------
    def viterbi(self, log_density, log_prior, log_trans):
        """log_density: device float64 [S, N]; returns the device int64 path [N]."""
        torch = self.torch
        S, N = log_density.shape
        dens = log_density.t().contiguous()  # frame major (layout plumbing)
        ws = torch.empty((int(self.lib.pf_viterbi_workspace_bytes(S, N)) + 7) // 8,
                         dtype=torch.float64, device=log_density.device)
        path = torch.empty(N, dtype=torch.int64, device=log_density.device)
        self._call("pf_viterbi", self._pv(dens), self._pv(log_prior), self._pv(log_trans), S, N,
                   self._pv(ws), ws.numel() * 8, path.data_ptr())
        return path

    def overlap_norm(self, prod, hop, N):
        """Device float64 [hop (N-1) + wlen]: the overlap-added window product of the inverse
        STFT (prod: host float64 [wlen])."""
        torch = self.torch
        pd = torch.tensor(np.ascontiguousarray(prod, dtype=np.float64)).to(self.device)
        norm = torch.empty(int(hop) * (int(N) - 1) + pd.numel(), dtype=torch.float64,
                           device=self.device)
        self._call("pf_overlap_norm", self._pv(pd), pd.numel(), int(hop), int(N), self._pv(norm))
        return norm

    # -- glottal-source F0 dictionary ------------------------------------------------------------
    def wf0_combs(self, f1, f2, npart, fs, Ot, Lsig, t_begin, window, nfft, rows):
        """f1, f2 (float64), npart (int32): host arrays [ncols]; window: host float64 [wlen].
        Returns the device float64 array [ncols, rows] of comb power spectra."""
        torch = self.torch
        dev = self.device
        f1d = torch.tensor(np.ascontiguousarray(f1, dtype=np.float64)).to(dev)
        f2d = torch.tensor(np.ascontiguousarray(f2, dtype=np.float64)).to(dev)
        npd = torch.tensor(np.ascontiguousarray(npart, dtype=np.int32)).to(dev)
        wd = torch.tensor(np.ascontiguousarray(window, dtype=np.float64)).to(dev)
        out = torch.empty((f1d.numel(), int(rows)), dtype=torch.float64, device=dev)
        self._call("pf_wf0_combs", self._pv(f1d), self._pv(f2d), npd.data_ptr(), f1d.numel(),
                   int(np.max(npart)) if len(npart) else 0, float(fs), float(Ot), int(Lsig),
                   int(t_begin), self._pv(wd), wd.numel(), int(nfft), int(rows), self._pv(out))
        return out

    # -- IS-NMF initialisers ---------------------------------------------------------------
    def nmf_is_terms(self, hat, SX, out, eps, F, N, ldn):
        self._call("pf_nmf_is_terms", self._pv(hat), self._pv(SX), self._pv(out), float(eps), F, N,
                   ldn)

    def nmf_update_rows(self, H, C, ldn, eps, rows, N):
        self._call("pf_nmf_update_rows", self._pv(H), H.stride(0), self._pv(C), C.stride(0), ldn,
                   float(eps), rows, N)

    def nmf_w_update(self, W, K, D, eps, F, s_out):
        self._call("pf_nmf_w_update", self._pv(W), W.stride(0), K, self._pv(D), float(eps), F,
                   self._pv(s_out))

    def mono_power(self, X, out, F, N, ldn):
        self._call("pf_mono_power", self._pv(X), X.stride(1), X.shape[0] // 2, self._pv(out), F, N,
                   ldn)

    def simm_wm_scaled(self, WM, R, b2, nch, F, WMs):
        self._call("pf_simm_wm_scaled", self._pv(WM), WM.stride(0), R, self._pv(b2), nch, F,
                   self._pv(WMs))

    # -- K5 ---------------------------------------------------------------------------
    def spat_energy(self, A, src_of_sub, J, sums):
        R, I, F = A.shape
        _check(self.lib.pf_spat_energy(self._p(A), _iarr(src_of_sub), R, J, I, F, self._p(sums),
                                       self._stream()), self.lib)

    def spat_scale(self, A, src_of_sub, sums, counts):
        R, I, F = A.shape
        _check(self.lib.pf_spat_scale(self._p(A), _iarr(src_of_sub), R, I, F, self._p(sums),
                                      self._p(counts), self._stream()), self.lib)

    def fb_scale_colmax(self, FB, sums, counts, j, colmax):
        F, K = FB.shape
        _check(self.lib.pf_fb_scale_colmax(self._p(FB), FB.stride(0), F, K, self._p(sums),
                                           self._p(counts), j, self._p(colmax),
                                           self.dtype_code(FB), self._stream()), self.lib)

    def fw_renorm(self, FW, colmax, w, w2):
        Kb, Kw = FW.shape
        _check(self.lib.pf_fw_renorm(self._p(FW), FW.stride(0), Kb, Kw, self._p(colmax),
                                     self._p(w), self._p(w2), self.dtype_code(FW),
                                     self._stream()), self.lib)

    def scale_matrix(self, M, rows, cols, s, by_row, divide, total=None):
        _check(self.lib.pf_scale_matrix(self._p(M), M.stride(0), rows, cols, self._p(s),
                                        int(by_row), int(divide), self._p(total),
                                        self.dtype_code(M), self._stream()), self.lib)

    # -- general factor structures -----------------------------------------------------------------
    def gem_ratio_planes(self, hatW, P, O, out, N, Ptot=None, Pminus=None, lam=0.0):
        F, ld = hatW.shape
        _check(self.lib.pf_gem_ratio_planes(self._p(hatW), self._p(P), self._p(O), self._p(out), F,
                                            N, ld, self._p(Ptot), self._p(Pminus), float(lam),
                                            self.dtype_code(hatW), self._stream()), self.lib)

    def corr_planes(self, V, own, Ptot, Pminus, N, clamp=True):
        J, F, ld = V.shape
        _check(self.lib.pf_corr_planes(self._p(V), J, own, self._p(Ptot), self._p(Pminus), F, N, ld,
                                       int(clamp), self.dtype_code(V), self._stream()), self.lib)

    def row_sums(self, M, rows, cols, out):
        _check(self.lib.pf_row_sums(self._p(M), M.stride(0), rows, cols, self._p(out),
                                    self.dtype_code(M), self._stream()), self.lib)

    def apply_filter(self, X, W, Y, N):
        """Y[c1] = sum_c2 W[c1, c2] X[c2] per bin; W complex128 [nc, nc, F] or [nc, nc, F, >= N]."""
        nc, F, ld = X.shape[0] // 2, X.shape[1], X.shape[2]
        wn = W.shape[3] if W.dim() == 4 else 0
        _check(self.lib.pf_apply_filter(self._p(X), self._p(W), self._p(Y), nc, F, N, ld, wn,
                                        self.dtype_code(X), self._stream()), self.lib)

    def mul_planes(self, a, b, out, N, accumulate=False):
        F, ld = a.shape
        _check(self.lib.pf_mul_planes(self._p(a), self._p(b), self._p(out), F, N, ld,
                                      int(accumulate), self.dtype_code(a), self._stream()), self.lib)

    def mult_update_same(self, theta, num, den, rows, cols, omega):
        _check(self.lib.pf_mult_update_same(self._p(theta), theta.stride(0), self._p(num),
                                            num.stride(0), self._p(den), den.stride(0), rows, cols,
                                            float(omega), self.dtype_code(theta), self._stream()),
               self.lib)

    def sparsity_reweigh(self, TW, K, N, length, log_sigma0, slope, iter_dev, work):
        _check(self.lib.pf_sparsity_reweigh(self._p(TW), TW.stride(0), K, N, int(length),
                                            float(log_sigma0), float(slope), iter_dev.data_ptr(),
                                            self._p(work), self.dtype_code(TW), self._stream()),
               self.lib)

    def check_totals(self, totals, eps, flags, iter_dev=None, first_iter=None):
        """flags |= 2 where totals < eps (and first_iter = min(first_iter, *iter_dev)); totals = 0."""
        _check(self.lib.pf_check_totals(self._p(totals), totals.numel(), float(eps),
                                        self._p(flags), self._p(iter_dev), self._p(first_iter),
                                        self._stream()), self.lib)

    def noise_anneal(self, sqrt0, sqrt1, iter_dev, n_iter, noise):
        _check(self.lib.pf_noise_anneal(self._p(sqrt0), self._p(sqrt1), self._p(iter_dev), n_iter,
                                        noise.numel(), self._p(noise), self._stream()), self.lib)

    def ll_reduce(self, ll_f, ll_sum):
        _check(self.lib.pf_ll_reduce(self._p(ll_f), ll_f.numel(), self._p(ll_sum),
                                     self._stream()), self.lib)

    def ll_store(self, ll_sum, bins, logliks, iter_dev, advance):
        _check(self.lib.pf_ll_store(self._p(ll_sum), float(bins), self._p(logliks),
                                    self._p(iter_dev), int(advance), self._stream()), self.lib)
