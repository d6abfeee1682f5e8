"""STFT / inverse STFT on the device (drop-in for pyfasst/tftransforms/stft.py).

Same functions, class and argument meaning as the reference (`stft` :3-69, `istft`
:71-131, class `STFT` :339-394); the frame loops are replaced by the K1 / K6 kernels
(pyfasst_b200/csrc/stft.cu) called through the C ABI.  The host-facing methods return
NumPy arrays like the reference; `stft_planes` / `istft_planes` keep everything in HBM
for the FASST classes.
"""
import numpy as np

from ..tools.utils import *  # noqa: F401,F403
from ..tools.utils import sinebell

_kernels = None


def default_kernels():
    """The CUDA kernels (created on first use; raises without a GPU / built library)."""
    global _kernels
    if _kernels is None:
        from .._lib import CudaKernels
        _kernels = CudaKernels()
    return _kernels


def _tdtype(torch, name):
    return {"float32": torch.float32, "float64": torch.float64}[name]


def number_of_frames(length, hopsize):
    """N = ceil(L / hop) + 2 (ref: stft.py:40)"""
    return int(np.ceil(length / float(hopsize)) + 2)


def overlap_norm(window, analysisWindow, hopsize, nframes):
    """Overlap-added window product of istft (ref: stft.py:106-129), zeros -> 1."""
    wlen = window.size
    prod = window * analysisWindow
    norm = np.zeros(hopsize * (nframes - 1) + wlen)
    for n in range(nframes):
        norm[n * hopsize:n * hopsize + wlen] += prod
    norm[norm == 0] = 1.0
    return norm


def stft_planes(kernels, pcm, window, hopsize, nfft, dtype="float64", psd_sum=None,
                pcm_div=1.0, frames=None, sample0=0, L_total=None):
    """pcm: device float64 [nch, L] (planar) or int16 / int32 / float32 [L, nch] as read
    from a WAV file; every sample is divided by `pcm_div` on the device.  `pcm` may be the
    window [sample0, sample0+L) of a signal of L_total samples and `frames` = (n_lo, n_hi)
    the range of frames to compute (frame sharding).
    Returns (X planes [2*nch, F, ld] of `dtype`, number of frames computed)."""
    import torch
    if pcm.dtype == torch.float64:
        nch, L = pcm.shape
    else:
        L, nch = pcm.shape
    L_total = L if L_total is None else L_total
    n_lo, n_hi = (0, number_of_frames(L_total, hopsize)) if frames is None else frames
    N = n_hi - n_lo
    ld = (N + 31) // 32 * 32
    F = nfft // 2 + 1
    X = torch.zeros([2 * nch, F, ld], dtype=_tdtype(torch, dtype), device=pcm.device)
    win = torch.tensor(np.asarray(window, dtype=np.float64)).to(pcm.device)
    kernels.stft(pcm, win, int(hopsize), int(nfft), X, N, psd_sum, pcm_div, sample0, L_total, n_lo)
    return X, N


def istft_planes(kernels, Y, N, window, analysisWindow, hopsize, nfft, length=None,
                 maxdata=None):
    """Y: device planes [2*nsig, F, ld].  Returns device float64 [nsig, length] (and int16
    PCM [length, nsig] when `maxdata` is given)."""
    import torch
    nsig = Y.shape[0] // 2
    wlen = window.size
    total = hopsize * (N - 1) + wlen
    if length is None:
        length = total - wlen // 2
    dev = Y.device
    # the overlap-added window product is formed on the device (a host loop over the frames of a
    # 10-minute signal plus the upload of its 212 MB took longer than the transform itself)
    norm = kernels.overlap_norm(np.asarray(window) * np.asarray(analysisWindow), hopsize, N)
    norm[norm == 0] = 1.0
    synth = torch.tensor(np.asarray(window, dtype=np.float64)).to(dev)
    out = torch.zeros([nsig, length], dtype=torch.float64, device=dev)
    pcm = None
    if maxdata is not None:
        pcm = torch.zeros([length, nsig], dtype=torch.int16, device=dev)
    kernels.istft(Y, N, synth, norm, int(hopsize), int(nfft), out, pcm,
                  1.0 if maxdata is None else maxdata)
    return out, pcm


def stft(data, window=sinebell(2048), hopsize=256.0, nfft=2048.0, fs=44100.0, kernels=None):
    """X, F, N = stft(data, window, hopsize, nfft, fs)   (ref: stft.py:3-69)"""
    import torch
    k = kernels or default_kernels()
    hopsize, nfft = int(hopsize), int(nfft)
    data = np.asarray(data, dtype=np.float64).reshape(1, -1)
    pcm = torch.tensor(data).to(k.device)
    X, N = stft_planes(k, pcm, np.asarray(window), hopsize, nfft, "float64")
    Xh = X[:, :, :N].cpu().numpy()
    freqs = np.arange(nfft // 2 + 1) / np.double(nfft) * fs
    times = np.arange(N) * hopsize / np.double(fs)
    return Xh[0] + 1j * Xh[1], freqs, times


def istft(X, window=sinebell(2048), analysisWindow=None, hopsize=256.0, nfft=2048.0,
          kernels=None):
    """data = istft(X, window, analysisWindow, hopsize, nfft)   (ref: stft.py:71-131)"""
    import torch
    k = kernels or default_kernels()
    if analysisWindow is None:
        analysisWindow = window
    hopsize, nfft = int(hopsize), int(nfft)
    X = np.asarray(X)
    F, N = X.shape
    ld = (N + 31) // 32 * 32
    planes = np.zeros([2, F, ld])
    planes[0, :, :N], planes[1, :, :N] = X.real, X.imag
    Y = torch.tensor(planes).to(k.device)
    out, _ = istft_planes(k, Y, N, np.asarray(window), np.asarray(analysisWindow), hopsize, nfft)
    return out[0].cpu().numpy()


def filter_stft(data, W, analysisWindow=None, synthWindow=sinebell(2048), hopsize=256.0,
                nfft=2048.0, fs=44100.0, kernels=None):
    """ndata = filter_stft(data, W, ...)   (ref: stft.py:133-227): STFT of every channel of
    `data` [T, M], multiplication by the M x M x F (x N) filter W in every bin, inverse STFT with
    overlap-add.  `ceil(T / hop)` frames, the first one centred on the first sample; the result
    has (frames - 1) * hop + window / 2 samples, like the reference's."""
    import torch
    k = kernels or default_kernels()
    data = np.asarray(data, dtype=np.float64)
    W = np.asarray(W)
    ns, nc = data.shape
    if nc != W.shape[0]:
        raise AttributeError("W does not have the right number of channels")
    synthWindow = np.asarray(synthWindow, dtype=np.float64)
    if analysisWindow is None or len(analysisWindow) != len(synthWindow):
        analysisWindow = synthWindow
    analysisWindow = np.asarray(analysisWindow, dtype=np.float64)
    hopsize, nfft = int(hopsize), int(nfft)
    N = int(np.ceil(ns / np.double(hopsize)))
    if nfft // 2 + 1 != W.shape[2]:
        raise AttributeError("W not the right size")
    if W.ndim not in (3, 4):
        raise NotImplementedError("For W.ndim== 3 or 4" + str(W.ndim))
    if W.ndim == 4 and W.shape[3] < N:
        raise IndexError("W holds %d frames, the signal has %d" % (W.shape[3], N))
    pcm = torch.tensor(np.ascontiguousarray(data.T)).to(k.device)
    X, _ = stft_planes(k, pcm, analysisWindow, hopsize, nfft, "float64", frames=(0, N))
    Wd = torch.tensor(np.ascontiguousarray(W.astype(np.complex128))).to(k.device)
    Y = torch.zeros_like(X)
    k.apply_filter(X, Wd, Y, N)
    out, _ = istft_planes(k, Y, N, synthWindow, analysisWindow, hopsize, nfft)
    return np.ascontiguousarray(out.cpu().numpy().T)


class STFT(object):
    """ref: stft.py:339-394 (same constructor, attributes and methods)."""
    transformname = 'stft'

    def __init__(self, linFTLen=2048, atomHopFactor=0.25, winFunc=np.hanning, fs=44100,
                 synthWinFunc=None, **kwargs):
        self.ftlen = linFTLen
        self.freqbins = self.ftlen // 2 + 1
        self.atomHopFactor = atomHopFactor
        self.fthop = int(linFTLen * atomHopFactor)
        if winFunc is None:
            winFunc = np.hanning
        self.winFunc = winFunc
        self.window = self.winFunc(self.ftlen)
        self.synthWinFunc = synthWinFunc if synthWinFunc is not None else self.winFunc
        self.synthWindow = self.synthWinFunc(self.ftlen)
        self.fs = fs
        self._kernels = kwargs.get("kernels")

    def computeTransform(self, data):
        self.transfo, self.freq_stamps, self.time_stamps = stft(
            data=data, window=self.window, hopsize=self.fthop, fs=self.fs, nfft=self.ftlen,
            kernels=self._kernels)
        self.datalen_init = np.size(data)
        self.time_stamps *= self.fs  # in samples, like the reference (stft.py:385)

    def invertTransform(self):
        return istft(X=self.transfo, window=self.synthWindow, analysisWindow=self.window,
                     hopsize=self.fthop, nfft=self.ftlen,
                     kernels=self._kernels)[:self.datalen_init]
