"""SeparateLeadProcess: lead / accompaniment separation with the (stereo) SIMM model -- the part
of pyfasst/SeparateLeadStereo/SeparateLeadStereoTF.py that lies on the hot path (SURVEY.md 8a
rows a14-a16): construction (WAV, 1.2 max scaling, STFT parameters :388-420), `computeMonoX`
(:702-739), `computeStereoX` / `computeStereoSX` (:761-917), `estimSIMMParams` (:919-957),
`estimStereoSIMMParams` (:1677-1713) and `writeSeparatedSignals` (:1762-1871), with the
reference's attribute names (`files`, `stftParams`, `SIMMParams`, `XR`, `XL`, `scaleData`).

The spectrograms, the SIMM parameter estimation, the Wiener masks and the inverse transforms
run on the GPU; there is no CPU fallback.

Not here (SURVEY.md 8f "next" rows 3 and 4): the Viterbi melody tracking between the two
estimation stages (`runViterbi`, `autoMelSepAndWrite`), the chunked processing, and the
generation of the glottal F0 dictionary -- `WF0` (F x NF0) must be given to the constructor;
`computeWF0` raises NotImplementedError without it.  Only `tfrepresentation='stft'`.
"""
import os

import numpy as np
import scipy.io.wavfile as wav

from . import separateLeadFunctions as slf
from .SIMM import SIMM

eps = 10 ** -9  # SeparateLeadStereoTF.py:31


class SeparateLeadProcess(object):
    def __init__(self, inputAudioFilename, windowSize=0.0464, hopsize=None, NFT=None, nbIter=10,
                 numCompAccomp=40, minF0=39, maxF0=2000, stepNotes=16, chirpPerF0=1,
                 K_numFilters=4, P_numAtomFilters=30, imageCanvas=None, wavCanvas=None,
                 progressBar=None, verbose=True, outputDirSuffix='/', minF0search=None,
                 maxF0search=None, tfrepresentation='stft', initHF00='random', freeMemory=True,
                 WF0=None, kernels=None):
        if tfrepresentation != 'stft':
            raise NotImplementedError("pyfasst_b200: only tfrepresentation='stft'")
        # per-instance (the reference shares these dicts between instances, :259-261)
        self.files, self.stftParams, self.SIMMParams = {}, {}, {}
        self.verbose = verbose
        self.tfrepresentation = tfrepresentation
        self.displayEvolution = False
        self.imageCanvas = None
        self._kernels = kernels
        self.files['inputAudioFilename'] = str(inputAudioFilename)
        self.setOutputFileNames(outputDirSuffix)
        # the WAV file, rescaled to +-1/1.2 (:388-395)
        self.fs, data = wav.read(self.files['inputAudioFilename'])
        self.scaleData = 1.2 * np.abs(data).max()
        self.dataType = data.dtype
        self.numberChannels = 1 if data.ndim == 1 else min(data.shape[1], 2)
        self.stftParams['windowSizeInSamples'] = slf.nextpow2(np.round(windowSize * self.fs))
        self.stftParams['hopsize'] = (self.stftParams['windowSizeInSamples'] / 8.
                                      if hopsize is None else np.double(hopsize))
        self.stftParams['NFT'] = self.stftParams['windowSizeInSamples'] if NFT is None else NFT
        self.stftParams['offsets'] = {'stft': self.stftParams['windowSizeInSamples'] / 2}
        self.SIMMParams.update(niter=nbIter, R=numCompAccomp, minF0=minF0, maxF0=maxF0,
                               stepNotes=stepNotes, K=K_numFilters, P=P_numAtomFilters,
                               chirpPerF0=chirpPerF0, initHF00=initHF00, HF00=None,
                               F0Table=None)
        self.F = int(self.stftParams['NFT']) // 2 + 1
        self.SIMMParams['WF0'] = None if WF0 is None else np.asarray(WF0, dtype=np.float64)
        self.computeWF0()
        self.SIMMParams['WGAMMA'] = slf.generateHannBasis(
            numberFrequencyBins=self.F, sizeOfFourier=self.stftParams['NFT'], Fs=self.fs,
            frequencyScale='linear', numberOfBasis=self.SIMMParams['P'], overlap=.75)
        self.freeMemory = freeMemory

    # -- files -------------------------------------------------------------------------------
    def setOutputFileNames(self, outputDirSuffix):
        """Output names next to the input file (ref: :540-585)."""
        f = self.files
        f['outputDirSuffix'] = outputDirSuffix
        f['outputDir'] = str('/').join(f['inputAudioFilename'].split('/')[:-1]) + '/' + \
            outputDirSuffix + '/'
        if not os.path.isdir(f['outputDir']):
            os.mkdir(f['outputDir'])
        f['pathBaseName'] = f['outputDir'] + f['inputAudioFilename'].split('/')[-1][:-4]
        f['mus_output_file'] = str(f['pathBaseName'] + '_acc.wav')
        f['voc_output_file'] = str(f['pathBaseName'] + '_lead.wav')
        f['pitch_output_file'] = str(f['pathBaseName'] + '_pitches.txt')

    def computeWF0(self):
        """The F0 dictionary.  Its generation (KLGLOTT88 glottal source, ref: :587-700,
        separateLeadFunctions.py:696-949) is not on this path: it must be supplied."""
        WF0 = self.SIMMParams['WF0']
        if WF0 is None:
            raise NotImplementedError(
                "pyfasst_b200: the glottal F0 dictionary generator is not implemented; pass "
                "WF0 (F x NF0) to SeparateLeadProcess")
        if WF0.shape[0] != self.F:
            raise ValueError("WF0 must have NFT/2+1 = %d rows, got %d" % (self.F, WF0.shape[0]))
        self.SIMMParams['NF0'] = WF0.shape[1]

    # -- time-frequency front end ---------------------------------------------------------------
    def _k(self):
        if self._kernels is None:
            from ..tftransforms.stft import default_kernels
            self._kernels = default_kernels()
        return self._kernels

    def _window(self):
        return slf.sinebell(self.stftParams['windowSizeInSamples'])

    def _read(self):
        _, data = wav.read(self.files['inputAudioFilename'])
        return np.double(data) / self.scaleData

    def _stft(self, x, start, stop):
        X, _, _ = slf.stft(x, fs=self.fs, hopsize=self.stftParams['hopsize'],
                           window=self._window(), nfft=self.stftParams['NFT'], start=start,
                           stop=stop, kernels=self._k())
        return X

    def computeMonoX(self, start=0, stop=None):
        """SX of the mean of the channels, floored at 1e-8 (ref: :702-739)."""
        data = self._read()
        if data.ndim > 1 and data.shape[1] > 1:
            data = data.mean(axis=1)
        X = self._stft(data, start, stop)
        self.F, _ = X.shape
        return np.maximum(np.abs(X) ** 2, 10 ** -8)

    def computeNFrames(self):
        """(ref: :741-759)"""
        data = self._read()
        self.totFrames = int(np.ceil(data.shape[0] / float(self.stftParams['hopsize']) + 1) + 1)
        return self.totFrames

    def computeStereoX(self, start=0, stop=None):
        """XR, XL: the STFT of each channel (ref: :761-841; mono files are duplicated)."""
        data = self._read()
        self.originalDataLen = data.shape[0]
        if data.ndim == 1:
            data = np.vstack([data, data]).T
        self.XR = self._stft(data[:, 0], start, stop)
        self.XL = self._stft(data[:, 1], start, stop)
        self.F, _ = self.XR.shape

    def computeStereoSX(self, start=0, stop=None):
        """SXR, SXL = max(|X|^2, 1e-8) (ref: :843-917)."""
        self.computeStereoX(start, stop)
        SXR = np.maximum(np.abs(self.XR) ** 2, 10 ** -8)
        SXL = np.maximum(np.abs(self.XL) ** 2, 10 ** -8)
        del self.XR, self.XL
        return SXR, SXL

    # -- parameter estimation ----------------------------------------------------------------------
    def estimSIMMParams(self, R=1):
        """Mono SIMM on the mean of the channels (ref: :919-957)."""
        SX = self.computeMonoX()
        p = self.SIMMParams
        HGAMMA, HPHI, HF0, HM, WM, _ = SIMM.SIMM(
            SX, WF0=p['WF0'], WGAMMA=p['WGAMMA'], numberOfFilters=p['K'],
            numberOfAccompanimentSpectralShapes=R, numberOfIterations=p['niter'],
            updateRulePower=1., stepNotes=p['stepNotes'], verbose=self.verbose,
            kernels=self._kernels)
        p.update(HGAMMA=HGAMMA, HPHI=HPHI, HF0=HF0, HM=HM, WM=WM)

    def estimStereoSIMMParams(self):
        """Stereo SIMM on the two channels, HF0 initialised with SIMMParams['HF00']
        (ref: :1677-1713)."""
        self.computeStereoX()
        SXR, SXL = np.abs(self.XR) ** 2, np.abs(self.XL) ** 2
        p = self.SIMMParams
        res = SIMM.Stereo_SIMM(
            SXR, SXL, WF0=p['WF0'], WGAMMA=p['WGAMMA'], numberOfFilters=p['K'],
            numberOfAccompanimentSpectralShapes=p['R'], HF00=p['HF00'],
            numberOfIterations=p['niter'], updateRulePower=1.0, stepNotes=p['stepNotes'],
            verbose=self.verbose, kernels=self._kernels)
        for nm, v in zip(('alphaR', 'alphaL', 'HGAMMA', 'HPHI', 'HF0', 'betaR', 'betaL', 'HM',
                          'WM'), res):
            p[nm] = v

    # -- separation -------------------------------------------------------------------------------
    def separated_pcm(self, suffix='.wav'):
        """The four separated signals as int16 PCM: (lead [L, 2], accompaniment [L, 2]).
        Masks, inverse STFTs and the rounding to PCM run on the device
        (ref: writeSeparatedSignals, :1762-1871)."""
        import torch
        from ..simm_engine import SimmEngine
        p = self.SIMMParams
        WF0, HF0 = (p['WUF0'], p['HUF0']) if 'VUIMM' in suffix else (p['WF0'], p['HF0'])
        k = self._k()
        betaR = np.diag(p['betaR']) if np.ndim(p['betaR']) == 2 else np.asarray(p['betaR'])
        F, N = self.XR.shape
        # the model planes SF0, SPHI, SM_c of the estimated parameters (no iteration is run)
        eng = SimmEngine(k, [np.zeros((F, N), np.float32)] * 2, WF0, p['WGAMMA'], p['HGAMMA'],
                         p['HPHI'], HF0, p['WM'], p['HM'], betaR=betaR)
        al = torch.tensor([p['alphaR'], p['alphaL']], dtype=torch.float64)
        eng.a2.copy_((al ** 2).to(torch.float32))
        hop, nfft = int(self.stftParams['hopsize']), int(self.stftParams['NFT'])
        ldx = (N + 31) // 32 * 32
        Xh = np.zeros((4, F, ldx), np.float32)
        Xh[0, :, :N], Xh[1, :, :N] = self.XR.real, self.XR.imag
        Xh[2, :, :N], Xh[3, :, :N] = self.XL.real, self.XL.imag
        X = torch.from_numpy(Xh).to(k.device)
        Y = torch.zeros((8, F, ldx), dtype=torch.float32, device=k.device)
        k.simm_masks(eng.SM, eng.SF0, eng.SPHI, eng.a2, X, Y, eps, 2, F, N, eng.ldn)
        win = self._window()
        _, pcm = slf.istft_planes(k, Y, N, win, win, hop, nfft, scale=float(self.scaleData))
        pcm = pcm.cpu().numpy().astype(self.dataType)
        return pcm[:, 0:2], pcm[:, 2:4]

    def writeSeparatedSignals(self, suffix='.wav'):
        """Writes the lead and the accompaniment to files['voc_output_file'] /
        files['mus_output_file'] (with `suffix` replacing '.wav'); 'VUIMM' in the suffix selects
        the dictionary with the unvoiced elements (ref: :1762-1871)."""
        voc, mus = self.separated_pcm(suffix)
        wav.write(self.files['voc_output_file'][:-4] + suffix, self.fs, voc)
        wav.write(self.files['mus_output_file'][:-4] + suffix, self.fs, mus)

    def writeSeparatedSignalsWithUnvoice(self):
        """(ref: :1873-1878)"""
        self.writeSeparatedSignals(suffix='_VUIMM.wav')
