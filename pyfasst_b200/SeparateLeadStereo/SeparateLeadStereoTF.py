"""SeparateLeadProcess: lead / accompaniment separation with the (stereo) SIMM model -- the part
of pyfasst/SeparateLeadStereo/SeparateLeadStereoTF.py that lies on the hot path (SURVEY.md 8a
rows a14-a16): construction (WAV, 1.2 max scaling, STFT parameters :388-420), `computeMonoX`
(:702-739), `computeStereoX` / `computeStereoSX` (:761-917), `estimSIMMParams` (:919-957),
`estimStereoSIMMParams` (:1677-1713) and `writeSeparatedSignals` (:1762-1871), with the
reference's attribute names (`files`, `stftParams`, `SIMMParams`, `XR`, `XL`, `scaleData`).

The spectrograms, the SIMM parameter estimation, the Wiener masks and the inverse transforms
run on the GPU; there is no CPU fallback.

Also here: the melody tracking between the two estimation stages --
`estimHF0` (:959-1072), `runViterbi` (:1150-1319, the Viterbi decoding itself on the GPU,
tracking/_tracking.py), `initiateHF0WithIndexBestPath` (:1321-1368) -- and the chunked second
stage `estimStereoSIMMParamsWriteSeps` / `overlapAddChunks` / `checkChunkSize` (:1370-1583,
:1880-1899), i.e. the whole of `autoMelSepAndWrite` (:1142-1148).

And (8f row 4) the glottal F0 dictionary: `computeWF0` (:587-700) generates it on the GPU
(separateLeadFunctions.generate_WF0_TR_chirped, csrc/wf0.cu) with the reference's `.npz` cache,
unless `WF0` (F x NF0, and optionally `F0Table`) is given to the constructor.  Only
`tfrepresentation='stft'`.
"""
import os

import numpy as np
import scipy.io.wavfile as wav

from . import separateLeadFunctions as slf
from ..tftransforms.stft import STFT
from ..tools.utils import sqrt_blackmanharris
from .SIMM import SIMM
from .tracking._tracking import viterbiTracking as viterbiTrackingArray

eps = 10 ** -9  # SeparateLeadStereoTF.py:31


class SeparateLeadProcess(object):
    def __init__(self, inputAudioFilename, windowSize=0.0464, hopsize=None, NFT=None, nbIter=10,
                 numCompAccomp=40, minF0=39, maxF0=2000, stepNotes=16, chirpPerF0=1,
                 K_numFilters=4, P_numAtomFilters=30, imageCanvas=None, wavCanvas=None,
                 progressBar=None, verbose=True, outputDirSuffix='/', minF0search=None,
                 maxF0search=None, tfrepresentation='stft', cqtfmax=4000, cqtfmin=50, cqtbins=48,
                 cqtWinFunc=sqrt_blackmanharris, cqtAtomHopFactor=0.25, initHF00='random',
                 freeMemory=True, WF0=None, F0Table=None, kernels=None):
        if tfrepresentation != 'stft':
            raise NotImplementedError("pyfasst_b200: only tfrepresentation='stft'")
        # per-instance (the reference shares these dicts between instances, :259-261)
        self.files, self.stftParams, self.SIMMParams = {}, {}, {}
        # the transform object of the dictionary (ref: :334-339; the cqt* values other than the
        # window and the hop factor are ignored by the STFT, tftransforms/stft.py:360-362)
        self.stftParams.update(cqtfmin=cqtfmin, cqtfmax=cqtfmax, cqtbins=cqtbins,
                               cqtWinFunc=cqtWinFunc, cqtAtomHopFactor=cqtAtomHopFactor)
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
        self.stftParams['offsets'] = {'stft': self.stftParams['windowSizeInSamples'] // 2}
        self.SIMMParams.update(niter=nbIter, R=numCompAccomp, minF0=minF0, maxF0=maxF0,
                               stepNotes=stepNotes, K=K_numFilters, P=P_numAtomFilters,
                               chirpPerF0=chirpPerF0, initHF00=initHF00, HF00=None,
                               F0Table=None if F0Table is None else np.asarray(F0Table))
        self.scopeAllowedHF0 = 4.0 / 1.0
        self.trackingParams = {
            'minF0search': minF0 if minF0search is None else minF0search,
            'maxF0search': maxF0 if maxF0search is None else maxF0search}
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
        """The F0 dictionary (ref: :587-700, the `stft` branch :661-684): glottal harmonic combs
        through the STFT transform object, columns normalised to sum one -- generated on the
        GPU (or read from the reference's .npz cache in the working directory) unless WF0 was
        given to the constructor."""
        WF0 = self.SIMMParams['WF0']
        if WF0 is None:
            self.mqt = STFT(linFTLen=int(self.stftParams['NFT']),
                            atomHopFactor=self.stftParams['cqtAtomHopFactor'],
                            winFunc=self.stftParams['cqtWinFunc'], fs=self.fs,
                            kernels=self._kernels)
            self.SIMMParams['F0Table'], WF0, self.mqt = slf.generate_WF0_TR_chirped(
                transform=self.mqt, minF0=self.SIMMParams['minF0'],
                maxF0=self.SIMMParams['maxF0'], stepNotes=self.SIMMParams['stepNotes'], Ot=0.5,
                perF0=self.SIMMParams['chirpPerF0'], depthChirpInSemiTone=0.5, loadWF0=True,
                verbose=self.verbose, kernels=self._kernels)
            WF0 = self.SIMMParams['WF0'] = WF0 / np.sum(WF0, axis=0)
            self.SIMMParams['NF0'] = self.SIMMParams['F0Table'].size
            self.F = WF0.shape[0]
            return
        if WF0.shape[0] != self.F:
            raise ValueError("WF0 must have NFT/2+1 = %d rows, got %d" % (self.F, WF0.shape[0]))
        self.SIMMParams['NF0'] = WF0.shape[1] // self.SIMMParams['chirpPerF0']
        if self.SIMMParams['F0Table'] is None:
            # the table generate_WF0_chirped builds (separateLeadFunctions.py:313-315)
            self.SIMMParams['F0Table'] = self.SIMMParams['minF0'] * 2 ** (
                np.arange(self.SIMMParams['NF0'], dtype=np.double)
                / (12 * self.SIMMParams['stepNotes']))

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
        """Number of frames of the whole file, as slf.stft counts them (ref: :741-759)."""
        if not hasattr(self, 'totFrames'):
            _, data = wav.read(self.files['inputAudioFilename'])
            self.lengthData = data.shape[0]
            self.totFrames = np.int32(
                np.ceil((self.lengthData - 0) / self.stftParams['hopsize'] + 1) + 1)
            self.N = self.totFrames
        return self.totFrames

    def checkChunkSize(self, maxFrames):
        """Number of chunks of maxFrames frames; evens the chunks out when the last one would be
        shorter than a window (ref: :1880-1899)."""
        totFrames = np.int32(self.computeNFrames())
        nChunks = int(totFrames // maxFrames + 1)
        if (totFrames - (nChunks - 1) * maxFrames <
                self.stftParams['windowSizeInSamples'] / self.stftParams['hopsize']):
            maxFrames = int(np.ceil(np.double(totFrames) / nChunks))
            nChunks = int(totFrames // maxFrames)
        return totFrames, nChunks, maxFrames

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

    def estimHF0(self, R=1, maxFrames=1000):
        """First stage: mono SIMM chunk by chunk, keeping only HF0 (ref: :959-1072)."""
        totFrames, nChunks, maxFrames = self.checkChunkSize(maxFrames)
        p = self.SIMMParams
        p['HF0'] = np.zeros([p['NF0'] * p['chirpPerF0'], totFrames])
        for n in range(nChunks):
            start = n * maxFrames
            stop = int(np.minimum((n + 1) * maxFrames, totFrames))
            SX = self.computeMonoX(start=start, stop=stop)
            HF00 = None
            if p['initHF00'] == 'nnls':
                import scipy.optimize
                HF00 = np.ones((p['NF0'] * p['chirpPerF0'], stop - start))
                for framenb in range(stop - start):
                    HF00[:, framenb], _ = scipy.optimize.nnls(p['WF0'], SX[:, framenb])
                HF00 += eps
            _, _, HF0, _, _, _ = SIMM.SIMM(
                SX, WF0=p['WF0'], WGAMMA=p['WGAMMA'], numberOfFilters=p['K'],
                numberOfAccompanimentSpectralShapes=R, HF00=HF00,
                numberOfIterations=p['niter'], updateRulePower=1., stepNotes=p['stepNotes'],
                verbose=self.verbose, kernels=self._kernels)
            p['HF0'][:, start:stop] = np.copy(HF0)

    def autoMelSepAndWrite(self, maxFrames=1000):
        """Fully automated estimation of the melody and separation (ref: :1142-1148)."""
        self.estimHF0(maxFrames=maxFrames)
        self.runViterbi()
        self.initiateHF0WithIndexBestPath()
        self.estimStereoSIMMParamsWriteSeps(maxFrames=maxFrames)

    def runViterbi(self):
        """Viterbi decoding of the predominant F0 line from HF0 (ref: :1150-1319): banded
        transition model (geometric decay per note, flat beyond 10 notes, an extra silence state
        that the decoder -- like the reference's call -- never visits), pitch file written to
        files['pitch_output_file']."""
        p = self.SIMMParams
        if 'HF0' not in p:
            raise AttributeError("HF0 has probably not been estimated yet.")
        self.computeNFrames()
        scale = 1.0
        NF0 = p['NF0'] * p['chirpPerF0']
        nmaxF0, nminF0 = NF0, 0
        minF0, maxF0 = p['minF0'], p['maxF0']
        minF0search = self.trackingParams['minF0search']
        maxF0search = self.trackingParams['maxF0search']
        if minF0search > minF0 and minF0search < maxF0:
            nminF0 = np.where(p['F0Table'] >= minF0search)[0][0] * p['chirpPerF0']
        if maxF0search > minF0 and maxF0search < maxF0 and maxF0search > minF0search:
            nmaxF0 = (np.where(p['F0Table'] >= maxF0search)[0][0] + 1) * p['chirpPerF0']
        NF0 = int(nmaxF0 - nminF0)
        transitions = np.exp(-np.floor(np.arange(0, NF0) / p['stepNotes']) * scale)
        cutoffnote = int(np.minimum(NF0, 2 * 5 * p['stepNotes']))
        transitions[cutoffnote:] = transitions[cutoffnote - 1]
        T = np.zeros([NF0 + 1, NF0 + 1])  # Toeplitz
        b = np.arange(NF0)
        T[0:NF0, 0:NF0] = transitions[
            np.array(np.abs(np.outer(np.ones(NF0), b) - np.outer(b, np.ones(NF0))), dtype=int)]
        T[0:NF0, NF0] = transitions[cutoffnote - 1] * 10 ** (-90)
        T[NF0, 0:NF0] = transitions[cutoffnote - 1] * 10 ** (-80)
        T[NF0, NF0] = transitions[cutoffnote - 1] * 10 ** (-100)
        T = T / np.outer(np.sum(T, axis=1), np.ones(NF0 + 1))
        priorProbabilities = 1 / (NF0 + 1.0) * np.ones([NF0 + 1])
        logHF0 = np.zeros([NF0 + 1, self.N])
        normHF0 = np.amax(p['HF0'][nminF0:nmaxF0], axis=0)
        with np.errstate(divide='ignore'):
            logHF0[0:NF0, :] = np.log(p['HF0'][nminF0:nmaxF0])
        logHF0[0:NF0, normHF0 == 0] = np.amin(logHF0[logHF0 > -np.inf])
        logHF0[NF0, :] = np.maximum(np.amin(logHF0[logHF0 > -np.inf]), -100)
        indexBestPath = viterbiTrackingArray(NF0, self.N, logHF0, np.log(priorProbabilities),
                                             np.log(T), verbose=False, kernels=self._kernels)
        indexBestPath = indexBestPath + nminF0
        freqMelody = p['F0Table'][np.array(indexBestPath // p['chirpPerF0'], dtype=int)]
        freqMelody[indexBestPath == 0] = - freqMelody[indexBestPath == 0]
        np.savetxt(self.files['pitch_output_file'],
                   np.array([np.arange(self.N) * self.stftParams['hopsize'] / np.double(self.fs),
                             freqMelody]).T)
        self.indexBestPath = indexBestPath
        self.freqMelody = freqMelody

    def initiateHF0WithIndexBestPath(self):
        """HF00 for the second stage: a band of +-stepNotes/4 around the decoded line, at the
        maximum of HF0; silent where the line sits on the first or last state (ref: :1321-1368)."""
        p = self.SIMMParams
        NF0, chirpPerF0, stepNotes = p['NF0'], p['chirpPerF0'], p['stepNotes']
        HF00 = np.zeros([NF0 * chirpPerF0, self.N])
        half = np.floor(stepNotes / self.scopeAllowedHF0)
        width = int(chirpPerF0 * (2 * half + 1))
        dim1index = np.array(
            np.maximum(np.minimum(
                np.outer(self.indexBestPath, np.ones(width))
                + np.outer(np.ones(self.N), np.arange(-chirpPerF0 * half, chirpPerF0 * (half + 1))),
                chirpPerF0 * NF0 - 1), 0), dtype=int)
        dim1index = dim1index[self.indexBestPath != 0, :]
        dim1index = dim1index.reshape(1, dim1index.size)
        dim2index = np.outer(np.arange(self.N), np.ones(width, dtype=int))
        dim2index = dim2index[self.indexBestPath != 0, :]
        dim2index = dim2index.reshape(1, dim2index.size)
        HF00[dim1index, dim2index] = p['HF0'].max()
        HF00[:, self.indexBestPath == (NF0 - 1)] = 0.0
        HF00[:, self.indexBestPath == 0] = 0.0
        p['HF00'] = HF00

    def estimStereoSIMMParamsWriteSeps(self, maxFrames=1000):
        """Second stage chunk by chunk: Stereo_SIMM initialised with HF00 (and the previous
        chunk's HGAMMA), separation of the chunk, then overlap-add of the chunks' WAV files
        (ref: :1370-1467)."""
        totFrames, nChunks, maxFrames = self.checkChunkSize(maxFrames)
        p = self.SIMMParams
        p['HGAMMA'] = None
        for n in range(nChunks):
            start = n * maxFrames
            stop = int(np.minimum((n + 1) * maxFrames, totFrames))
            SXR, SXL = self.computeStereoSX(start=start, stop=stop)
            HF00 = np.zeros([p['NF0'] * p['chirpPerF0'], SXR.shape[1]])
            HF00[:, 0:stop - start] = p['HF00'][:, start:stop]
            res = SIMM.Stereo_SIMM(
                SXR, SXL, WF0=p['WF0'], WGAMMA=p['WGAMMA'], numberOfFilters=p['K'],
                numberOfAccompanimentSpectralShapes=p['R'], HGAMMA0=p['HGAMMA'], HF00=HF00,
                numberOfIterations=p['niter'], updateRulePower=1.0, stepNotes=p['stepNotes'],
                verbose=self.verbose, kernels=self._kernels)
            for nm, v in zip(('alphaR', 'alphaL', 'HGAMMA', 'HPHI', 'HF0', 'betaR', 'betaL', 'HM',
                              'WM'), res):
                p[nm] = v
            p['HF00'][:, start:stop] = np.copy(p['HF0'][:, 0:stop - start])
            self.computeStereoX(start=start, stop=stop)
            self.writeSeparatedSignals(suffix='%05d.wav' % n)
            del self.XR, self.XL
            if self.freeMemory:
                for nm in ('HM', 'HF0', 'HPHI', 'alphaR', 'alphaL', 'betaR', 'betaL'):
                    del p[nm]
        self.overlapAddChunks(nChunks=nChunks, suffixIsSUIMM='.wav')

    def overlapAddChunks(self, nChunks, suffixIsSUIMM='.wav'):
        """Concatenates the chunks' separated WAV files with overlap-add of the frames shared
        by consecutive chunks, removes them, and writes the final lead / accompaniment files
        (ref: :1469-1583)."""
        wlen = int(self.stftParams['windowSizeInSamples'])
        offsetTF = int(self.stftParams['offsets'][self.tfrepresentation])
        hopsize = int(self.stftParams['hopsize'])
        overlapSamp = wlen - hopsize
        overlapFunc = np.ones(overlapSamp)
        nuDataLen = int(self.totFrames * hopsize + 2 * wlen)
        for key in ('voc_output_file', 'mus_output_file'):
            data = np.zeros([nuDataLen, 2], np.int16)
            cumulframe = 0
            for n in range(nChunks):
                fname = self.files[key][:-4] + '%05d%s' % (n, suffixIsSUIMM)
                _, datatmp = wav.read(fname)
                datatmp = np.array(datatmp)
                datatype = type(datatmp[0][0])
                if n == 0 and nChunks != 1:
                    datatmp[-overlapSamp:, 0] = datatype(datatmp[-overlapSamp:, 0] * overlapFunc)
                    datatmp[-overlapSamp:, 1] = datatype(datatmp[-overlapSamp:, 1] * overlapFunc)
                    lendatatmp = datatmp.shape[0] - offsetTF
                    data[:lendatatmp, :] = np.copy(datatmp[offsetTF:, :])
                    cumulframe = lendatatmp
                elif nChunks != 1:
                    if n != nChunks - 1:
                        datatmp[-overlapSamp:, 0] = datatype(datatmp[-overlapSamp:, 0] * overlapFunc)
                        datatmp[-overlapSamp:, 1] = datatype(datatmp[-overlapSamp:, 1] * overlapFunc)
                    datatmp[:overlapSamp, 0] = datatype(datatmp[:overlapSamp, 0] * overlapFunc[::-1])
                    datatmp[:overlapSamp, 1] = datatype(datatmp[:overlapSamp, 1] * overlapFunc[::-1])
                    start = cumulframe - wlen + hopsize
                    stop = start + datatmp.shape[0]
                    data[start:stop, :] += datatmp
                    cumulframe = stop
                else:  # a single chunk
                    lendatatmp = datatmp.shape[0] - offsetTF
                    data[:lendatatmp] = datatmp[offsetTF:, :]
                os.remove(fname)
            wav.write(self.files[key][:-4] + suffixIsSUIMM, self.fs, data[:self.lengthData, :])

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
