"""Time-frequency front / back end of the SeparateLeadStereo (SIMM) path, drop-in for the
functions of pyfasst/SeparateLeadStereo/separateLeadFunctions.py that the path uses:
`stft` (:90-161), `istft` (:163-233), `sinebell`, `nextpow2`, `generateHannBasis`
(pyfasst/sourcefilter/filter.py:9-73), and the glottal-source F0 dictionary generators
`generate_WF0_chirped` (:237-345) / `generate_WF0_TR_chirped` (:696-886) with their `.npz`
cache files.  The transforms run on the GPU (csrc/stft.cu, shared
with the FASST front end -- the two conventions differ only in the window, in the patched
normalisation sequence of the inverse and in whether the leading half window is kept, quirk
Q9 of SURVEY.md); there is no CPU fallback.

Out of scope: the constant-Q variants generate_WF0_MinQT_chirped / _NSGTMinQT_chirped (other
transforms).
"""
import os

import numpy as np

from ..tftransforms import stft as _stft
from ..tools.utils import hann, nextpow2, sinebell  # noqa: F401  (re-exported like the reference)


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
    # overlap_norm() on the device: sums, then the reference's edge patch and zeros -> 1
    norm = kernels.overlap_norm(np.asarray(window) * np.asarray(analysisWindow), hopsize, N)
    wlen = window.size
    norm[:wlen] = norm[wlen:2 * wlen].clone()
    norm[-wlen:] = norm[-2 * wlen:-wlen].clone()
    norm[norm == 0] = 1.0
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


# ---- glottal-source F0 dictionary (KLGLOTT88) ------------------------------------------------
def _f0_table(minF0, maxF0, stepNotes):
    """F0 candidates, `stepNotes` per semitone (ref: separateLeadFunctions.py:313-316)."""
    minF0, maxF0, stepNotes = np.double(minF0), np.double(maxF0), np.double(stepNotes)
    numberOfF0 = int(np.ceil(12.0 * stepNotes * np.log2(maxF0 / minF0)) + 1)
    return minF0 * (2 ** (np.arange(numberOfF0, dtype=np.double) / (12 * stepNotes)))


def _comb_columns(F0Table, Fs, perF0, depthChirpInSemiTone):
    """(F1, F2, partials) of every column: the plain comb of each F0 followed by its perF0 - 1
    chirps, F0 the mean of F1 and F2 (ref: :318-341; partialMax :911, :1038)."""
    f1, f2 = [], []
    for F0 in F0Table:
        f1.append(F0)
        f2.append(F0)
        for chirpNumber in range(int(perF0) - 1):
            F2 = F0 * (2 ** ((chirpNumber + 1.0) * depthChirpInSemiTone / (12.0 * (perF0 - 1.0))))
            f1.append(2.0 * F0 - F2)
            f2.append(F2)
    f1, f2 = np.array(f1), np.array(f2)
    npart = np.floor((np.double(Fs) / 2) / np.maximum(f1, f2)).astype(np.int32)
    return f1, f2, npart


def _combs(kernels, F0Table, Fs, perF0, depth, Ot, Lsig, t_begin, window, nfft, rows, plain_window=None):
    """WF0 [rows, NF0 * perF0] from the comb kernel (csrc/wf0.cu)."""
    k = kernels or _stft.default_kernels()
    f1, f2, npart = _comb_columns(F0Table, Fs, perF0, depth)
    W = k.wf0_combs(f1, f2, npart, Fs, Ot, Lsig, t_begin, window, nfft, rows)
    W = np.ascontiguousarray(W.cpu().numpy().T)
    if plain_window is not None and perF0 > 1:
        # generate_WF0_chirped windows the plain combs with `analysisWindow` but the chirped
        # ones with the sinebell default of generate_ODGD_spec_chirped (:335-339)
        plain = np.arange(F0Table.size) * int(perF0)
        Wp = k.wf0_combs(f1[plain], f2[plain], npart[plain], Fs, Ot, Lsig, t_begin, plain_window,
                         nfft, rows)
        W[:, plain] = Wp.cpu().numpy().T
    return W


def generate_WF0_chirped(minF0, maxF0, Fs, Nfft=2048, stepNotes=4, lengthWindow=2048, Ot=0.5,
                         perF0=1, depthChirpInSemiTone=0.5, loadWF0=True, analysisWindow='hanning',
                         kernels=None):
    """F0Table, WF0 = generate_WF0_chirped(...)  (ref: separateLeadFunctions.py:237-345): the
    `Nfft` x (NF0 perF0) dictionary of glottal harmonic combs, |FFT|^2 of the windowed KLGLOTT88
    waveform of every F0 (and of its perF0 - 1 chirps), cached in the working directory under
    the reference's file name."""
    filename = str('').join(['wf0_', '_minF0-', str(minF0), '_maxF0-', str(maxF0), '_Fs-', str(Fs),
                             '_Nfft-', str(Nfft), '_stepNotes-', str(stepNotes), '_Ot-', str(Ot),
                             '_perF0-', str(perF0), '_depthChirp-', str(depthChirpInSemiTone),
                             '_analysisWindow-', analysisWindow, '.npz'])
    if os.path.isfile(filename) and loadWF0:
        struc = np.load(filename)
        return struc['F0Table'], struc['WF0']
    windows = {'sinebell': sinebell, 'hanning': hann, 'rectangular': np.ones}
    if analysisWindow not in windows:
        raise ValueError("Analysis window not understood.")
    lengthWindow, Nfft = int(lengthWindow), int(Nfft)
    F0Table = _f0_table(minF0, maxF0, stepNotes)
    chirp_win = sinebell(lengthWindow)
    plain_win = windows[analysisWindow](lengthWindow)
    WF0 = _combs(kernels, F0Table, Fs, perF0, depthChirpInSemiTone, Ot, lengthWindow, 0,
                 chirp_win if perF0 > 1 else plain_win, Nfft, Nfft,
                 plain_window=None if analysisWindow == 'sinebell' else plain_win)
    np.savez(filename, F0Table=F0Table, WF0=WF0)
    return F0Table, WF0


def generate_WF0_TR_chirped(transform, minF0, maxF0, stepNotes=4, Ot=0.5, perF0=1,
                            depthChirpInSemiTone=0.5, loadWF0=True, verbose=False, kernels=None):
    """F0Table, WF0, transform = generate_WF0_TR_chirped(transform, ...)  (ref: :696-886) for an
    STFT transform object (tftransforms.stft.STFT): every comb is the power of the STFT frame
    in the middle of a waveform of 2 linFTLen samples (old NumPy's rfft kept the real part of
    the complex waveform).  Cached under the reference's file name (F0Table and WF0 only)."""
    if hasattr(transform, 'cqtkernel') or hasattr(transform, 'octaveNr'):
        raise NotImplementedError("pyfasst_b200: only the STFT transform (DESIGN.md section 8)")
    try:
        lengthWindow = (transform.freqbins - 1) * 2 * 2  # "just to be sure" (:763)
    except AttributeError:
        raise AttributeError('There is something utterly wrong with the desired TF '
                             'representation...\nNo freqbins attribute!')
    # the reference's cache name: the scalar / function attributes of the transform (:768-799)
    keep = ('fmin', 'fmax', 'bins', 'fs', 'winFunc', 'freqbins', 'atomHopFactor')
    attributes = []
    for key, v in transform.__dict__.items():
        if key in keep and np.isscalar(v):
            attributes.append(key.lower() + '-' + str(v))
        elif key in keep and callable(v):
            attributes.append(v.__name__)
    attributes.sort()
    filename = str('').join(['wf0_%s_' % transform.transformname, '_minF0-', str(minF0),
                             '_maxF0-', str(maxF0), '_stepNotes-', str(int(stepNotes)),
                             '_Ot-', str(Ot), '_perF0-', str(int(perF0)),
                             '_depthChirp-', str(depthChirpInSemiTone),
                             '_lengthWindow-%d' % lengthWindow, '_', str('_').join(attributes),
                             '.npz'])
    if os.path.isfile(filename) and loadWF0:
        struc = np.load(filename, allow_pickle=True)
        return struc['F0Table'], struc['WF0'], transform
    Nfft = (transform.freqbins - 1) * 2
    hop = int(transform.fthop)
    window = np.asarray(transform.window, dtype=np.float64)
    # the frame nearest to the middle of the waveform; frame n is centred on sample n hop
    # (tftransforms/stft.py:40-63; midindex :849-850)
    nframes = int(np.ceil(lengthWindow / np.double(hop)) + 2)
    midindex = int(np.argmin((lengthWindow / 2. - np.arange(nframes) * float(hop)) ** 2))
    t_begin = midindex * hop - window.size // 2
    F0Table = _f0_table(minF0, maxF0, stepNotes)
    WF0 = _combs(kernels, F0Table, transform.fs, perF0, depthChirpInSemiTone, Ot, lengthWindow,
                 t_begin, window, Nfft, transform.freqbins)
    np.savez(filename, F0Table=F0Table, WF0=WF0)
    return F0Table, WF0, transform
