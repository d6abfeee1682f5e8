"""CPU oracle for the glottal-source F0 dictionary WF0 (KLGLOTT88).  TEST INFRASTRUCTURE ONLY.

float64 NumPy restatement of pyfasst/SeparateLeadStereo/separateLeadFunctions.py:
`generate_ODGD_spec` (:888-949), `generate_ODGD_spec_chirped` (:1010-1072),
`generate_WF0_chirped` (:237-345, the `stftold` dictionary) and `generate_WF0_TR_chirped`
(:696-886) for an STFT transform object (the only transform on this path), plus the
normalisation of `SeparateLeadProcess.computeWF0` (SeparateLeadStereoTF.py:587-700).

Parity status: PINNED by tests/golden/wf0.npz, produced by running the reference itself
(oracle/make_golden.py: run_wf0) -- see tests/test_wf0_cpu.py.

Old-NumPy behaviours restated deliberately:
  * generate_WF0_TR_chirped hands the COMPLEX waveform to `transform.computeTransform`, whose
    `np.fft.rfft` (NumPy of 2013) cast it to float, i.e. kept the real part
    (tftransforms/stft.py:59-63);
  * float sizes (`numberOfF0`) were truncated to integers.
"""
import numpy as np


def sinebell(n):
    """ref: tools/utils.py:43-57"""
    return np.sin(np.pi * np.arange(n) / (1.0 * n))


def glottal_amplitudes(F0, partial_max, Ot):
    """Complex amplitudes of the partials of the KLGLOTT88 glottal flow derivative
    (ref: separateLeadFunctions.py:917-929; :1044-1054)."""
    h = np.arange(1, partial_max + 1)
    t = 1j * 2.0 * np.pi * h * Ot
    return F0 * 27 / 4 * (np.exp(-t) + (2 * (1 + 2 * np.exp(-t)) / t)
                          - (6 * (1 - np.exp(-t)) / (t ** 2))) / t


def odgd(F0, Fs, length, Ot=0.5, t0=0.0):
    """Complex waveform of generate_ODGD_spec (ref: :912-941)."""
    F0, Fs = np.double(F0), np.double(Fs)
    pmax = int(np.floor((Fs / 2) / F0))
    h = np.arange(1, pmax + 1)
    amp = glottal_amplitudes(F0, pmax, Ot)
    ts = np.arange(length) / Fs + t0 / F0
    return np.sum(np.exp(np.outer(2.0 * 1j * np.pi * F0 * h, ts)) * amp[:, None], axis=0)


def odgd_chirped(F1, F2, Fs, length, Ot=0.5, t0=0.0):
    """Complex waveform of generate_ODGD_spec_chirped (ref: :1021-1067): linear chirp from
    F1 to F2 over the window, amplitudes of the mean F0."""
    F1, F2, Fs = np.double(F1), np.double(F2), np.double(Fs)
    F0 = (F1 + F2) / 2.0
    pmax = int(np.floor((Fs / 2) / max(F1, F2)))
    h = np.arange(1, pmax + 1)
    amp = glottal_amplitudes(F0, pmax, Ot)
    ts = np.arange(length) / Fs + t0 / F0
    ph = np.outer(F1 * h, ts) + np.outer((F2 - F1) * h, ts ** 2) / (2 * length / Fs)
    return np.sum(np.exp(2.0 * 1j * np.pi * ph) * amp[:, None], axis=0)


def f0_table(minF0, maxF0, stepNotes):
    """ref: :313-316 / :826-829"""
    minF0, maxF0, stepNotes = np.double(minF0), np.double(maxF0), np.double(stepNotes)
    n = int(np.ceil(12.0 * stepNotes * np.log2(maxF0 / minF0)) + 1)
    return minF0 * (2 ** (np.arange(n, dtype=np.double) / (12 * stepNotes)))


def chirp_pair(F0, chirp, perF0, depth):
    """F1, F2 of chirp number `chirp` (0-based) of a fundamental (ref: :328-333)."""
    F2 = F0 * (2 ** ((chirp + 1.0) * depth / (12.0 * (perF0 - 1.0))))
    return 2.0 * F0 - F2, F2


def generate_WF0_chirped(minF0, maxF0, Fs, Nfft=2048, stepNotes=4, lengthWindow=2048, Ot=0.5,
                         perF0=1, depthChirpInSemiTone=0.5, analysisWindow='hanning'):
    """ref: :237-345 (without the .npz cache).  NOTE the reference's window asymmetry: the
    plain combs use `analysisWindow`, the chirped ones always the default sinebell of
    generate_ODGD_spec_chirped (:335-339 pass no window)."""
    win = {'sinebell': sinebell, 'hanning': np.hanning, 'hann': np.hanning,
           'rectangular': np.ones}[analysisWindow](lengthWindow)
    table = f0_table(minF0, maxF0, stepNotes)
    WF0 = np.zeros([Nfft, table.size * perF0])
    for i, F0 in enumerate(table):
        WF0[:, i * perF0] = np.abs(np.fft.fft(np.real(odgd(F0, Fs, lengthWindow, Ot) * win),
                                              n=Nfft)) ** 2
        for c in range(perF0 - 1):
            F1, F2 = chirp_pair(F0, c, perF0, depthChirpInSemiTone)
            x = odgd_chirped(F1, F2, Fs, lengthWindow, Ot)
            WF0[:, i * perF0 + c + 1] = np.abs(np.fft.fft(np.real(x * sinebell(lengthWindow)),
                                                          n=Nfft)) ** 2
    return table, WF0


def stft_mid_frame(x, window, hop, nfft):
    """The frame of tftransforms/stft.py: stft (:3-69) nearest to the middle of `x`, as
    generate_WF0_TR_chirped picks it (:849-854): frame n is centred on sample n hop."""
    wlen = window.size
    nframes = int(np.ceil(x.size / np.double(hop)) + 2)
    n = int(np.argmin((x.size / 2. - np.arange(nframes) * float(hop)) ** 2))
    pad = np.concatenate((np.zeros(wlen // 2), x,
                          np.zeros((nframes - 1) * hop + wlen - wlen // 2 - x.size)))
    return np.fft.rfft(window * pad[n * hop:n * hop + wlen], nfft)


def generate_WF0_TR_chirped(ftlen, hop, winFunc, fs, minF0, maxF0, stepNotes=4, Ot=0.5, perF0=1,
                            depthChirpInSemiTone=0.5):
    """ref: :696-886 for `transform` = tftransforms.stft.STFT(linFTLen=ftlen,
    atomHopFactor=hop/ftlen, winFunc=winFunc, fs=fs): every comb is the power of the STFT
    frame in the middle of a 2 ftlen long waveform (lengthWindow = (freqbins - 1) * 2 * 2)."""
    length = (ftlen // 2) * 2 * 2
    window = winFunc(ftlen)
    table = f0_table(minF0, maxF0, stepNotes)
    WF0 = np.zeros([ftlen // 2 + 1, table.size * perF0])
    for i, F0 in enumerate(table):
        x = np.real(odgd(F0, fs, length, Ot))  # old rfft: complex input cast to float
        WF0[:, i * perF0] = np.abs(stft_mid_frame(x, window, hop, ftlen)) ** 2
        for c in range(perF0 - 1):
            F1, F2 = chirp_pair(F0, c, perF0, depthChirpInSemiTone)
            x = np.real(odgd_chirped(F1, F2, fs, length, Ot))
            WF0[:, i * perF0 + c + 1] = np.abs(stft_mid_frame(x, window, hop, ftlen)) ** 2
    return table, WF0


def normalise(WF0):
    """computeWF0: columns sum to one (ref: SeparateLeadStereoTF.py:611-613, :684)."""
    return WF0 / np.sum(WF0, axis=0)
