"""Time-frequency front / back end of the SeparateLeadStereo (SIMM) path, drop-in for the
functions of pyfasst/SeparateLeadStereo/separateLeadFunctions.py that the path uses:
`stft` (:90-161), `istft` (:163-233), `sinebell`, `nextpow2`, `generateHannBasis`
(pyfasst/sourcefilter/filter.py:9-73).  The transforms run on the GPU (csrc/stft.cu, shared
with the FASST front end -- the two conventions differ only in the window, in the patched
normalisation sequence of the inverse and in whether the leading half window is kept, quirk
Q9 of SURVEY.md); there is no CPU fallback.

Out of scope: generate_WF0_chirped / the KLGLOTT88 source model.
"""
import numpy as np

from ..tftransforms import stft as _stft
from ..tools.utils import nextpow2, sinebell  # noqa: F401  (re-exported like the reference)


def stft(data, window=sinebell(2048), hopsize=256.0, nfft=2048.0, fs=44100.0, start=0,
         stop=None, kernels=None):
    """X, F, N = stft(data, window, hopsize, nfft, fs, start, stop)
    (ref: separateLeadFunctions.py:90-161).  The signal is padded by half a window on both
    sides; `ceil(L / hop + 1) + 1` frames; X holds the frames [start, stop)."""
    import torch
    k = kernels or _stft.default_kernels()
    hopsize, nfft = int(hopsize), int(nfft)
    data = np.asarray(data, dtype=np.float64).reshape(1, -1)
    nframes = int(np.ceil(data.shape[1] / float(hopsize) + 1) + 1)
    stop = nframes if stop is None else int(stop)
    pcm = torch.tensor(data).to(k.device)
    X, n = _stft.stft_planes(k, pcm, np.asarray(window), hopsize, nfft, "float64",
                             frames=(int(start), stop))
    Xh = X[:, :, :n].cpu().numpy()
    F = np.arange(nfft // 2 + 1) / float(nfft) * fs
    N = np.arange(nframes) * hopsize / float(fs)
    return Xh[0] + 1j * Xh[1], F, N


def overlap_norm(window, analysisWindow, hopsize, nframes):
    """Normalisation sequence of `istft` (ref: separateLeadFunctions.py:205-222): the
    overlap-added window product, whose first and last window lengths are REPLACED by the next /
    previous ones (the reference's way of undoing the fade at the edges), zeros -> 1."""
    wlen = window.size
    prod = window * analysisWindow
    norm = np.zeros(hopsize * (nframes - 1) + wlen)
    for n in range(nframes):
        norm[n * hopsize:n * hopsize + wlen] += prod
    norm[:wlen] = norm[wlen:2 * wlen]
    norm[-wlen:] = norm[-2 * wlen:-wlen]
    norm[norm == 0] = 1.0
    return norm


def istft_planes(kernels, Y, N, window, analysisWindow, hopsize, nfft, originalDataLen=None,
                 scale=None):
    """Y: device planes [2 * nsig, F, ld] -> device float64 [nsig, length] and, when `scale`
    is given, int16 PCM [length, nsig] = round(y * scale) (SeparateLeadStereoTF.py:1826)."""
    import torch
    nsig = Y.shape[0] // 2
    total = hopsize * (N - 1) + window.size
    length = total if originalDataLen is None else min(int(originalDataLen), total)
    dev = Y.device
    norm = torch.tensor(overlap_norm(window, analysisWindow, hopsize, N)).to(dev)
    synth = torch.tensor(np.asarray(window, dtype=np.float64)).to(dev)
    out = torch.zeros([nsig, length], dtype=torch.float64, device=dev)
    pcm = None if scale is None else torch.zeros([length, nsig], dtype=torch.int16, device=dev)
    kernels.istft(Y, N, synth, norm, int(hopsize), int(nfft), out, pcm,
                  1.0 if scale is None else scale, drop=0, pcm_round=True)
    return out, pcm


def istft(X, analysisWindow=None, window=sinebell(2048), hopsize=256.0, nfft=2048.0,
          originalDataLen=None, start=-1, stop=None, kernels=None):
    """data = istft(X, analysisWindow, window, hopsize, nfft, originalDataLen)
    (ref: separateLeadFunctions.py:163-233): overlap-add with the synthesis window; the leading
    half window is NOT removed ("better do the cutting outside", :224-228)."""
    import torch
    k = kernels or _stft.default_kernels()
    window = np.asarray(window, dtype=np.float64)
    if analysisWindow is None:
        analysisWindow = window
    hopsize, nfft = int(hopsize), int(nfft)
    X = np.asarray(X)
    F, N = X.shape
    ld = (N + 31) // 32 * 32
    planes = np.zeros([2, F, ld])
    planes[0, :, :N], planes[1, :, :N] = X.real, X.imag
    Y = torch.tensor(planes).to(k.device)
    out, _ = istft_planes(k, Y, N, window, np.asarray(analysisWindow, dtype=np.float64), hopsize,
                          nfft, originalDataLen)
    return out[0].cpu().numpy()


def generateHannBasis(numberFrequencyBins, sizeOfFourier, Fs, frequencyScale='linear',
                      numberOfBasis=20, overlap=.75):
    """WGAMMA: overlapping Hann windows on a linear frequency scale, the smooth-filter
    dictionary of the source/filter model (ref: sourcefilter/filter.py:9-73).  Host side, a
    one-off F x P table."""
    if frequencyScale != 'linear':
        raise NotImplementedError("The desired feature for frequencyScale "
                                  "is not recognized yet...")
    nwin = np.ceil(1.0 / (1.0 - overlap))
    overlap = 1.0 - 1.0 / np.double(nwin)
    length = np.ceil(numberFrequencyBins / ((1.0 - overlap) * (numberOfBasis - 1) + 1
                                            - 2.0 * overlap))
    length = int(2.0 * np.floor(length / 2.0))
    big = 2 * int(numberFrequencyBins)
    centers = np.round(np.arange(-nwin + 1, numberOfBasis - nwin + 1) * (1 - overlap)
                       * np.double(length) + length / 2.0)
    bigWindow = np.zeros(big * 2)
    bigWindow[big - length // 2:big + length // 2] = np.hanning(length)
    WGAMMA = np.zeros([numberFrequencyBins, numberOfBasis])
    freq = np.arange(numberFrequencyBins)
    for p in range(numberOfBasis):
        WGAMMA[:, p] = bigWindow[np.int32(freq - centers[p] + big)]
    return WGAMMA
