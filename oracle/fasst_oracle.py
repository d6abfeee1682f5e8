"""CPU oracle for the FASST GEM / Wiener hot path.

TEST INFRASTRUCTURE ONLY.  This module is a float64 numpy restatement of the
reference algorithm (s-ben/pyfasst, Python 2) used as the *checker* for the CUDA
path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it; the product package `pyfasst_b200` never
does (it fails loudly when the CUDA library is missing).

Parity status: PINNED.  The reference cannot be imported as-is (Python 2), but
oracle/_py2shim.py makes its unmodified arithmetic executable here and
oracle/make_golden.py records its outputs in tests/golden/*.npz;
tests/test_oracle_golden.py checks every function below against those vectors
(and against the reference's own known-answer tests for the 2x2 inverse and the
window functions).

All `ref:` citations are file:line under /root/reference/pyfasst/.
Quirks of the reference that results depend on are kept on purpose (SURVEY.md
7.4, Q1-Q10); they are flagged with `QUIRK`.
"""
import os
import warnings

import numpy as np
import scipy.io.wavfile as wavfile

EPS = 1e-10  # ref: audioModel.py:61 and tools/signalTools.py (module eps)


# --------------------------------------------------------------------------- #
# small helpers (ref: tools/utils.py:30-72)
# --------------------------------------------------------------------------- #
def nextpow2(i):
    """Smallest power of two >= i, never below 2 (ref: tools/utils.py:30-41)."""
    n = 2
    while n < i:
        n *= 2
    return n


def sinebell(length):
    """sin(pi t / L), t = 0..L-1 (ref: tools/utils.py:43-57)."""
    return np.sin(np.pi * np.arange(length) / (1.0 * length))


def hann(length):
    """ref: tools/utils.py:59-65 (numpy's symmetric hanning)."""
    return np.hanning(length)


def inv_herm_mat_2d(diag, off):
    """Closed-form inverse of stacked 2x2 Hermitian matrices.

    ref: tools/signalTools.py:132-196.  `diag` is [2, ...] real, `off` the (0,1)
    element.  QUIRK (Q5): determinant clamp sign(det+eps)*max(|det|, eps).
    """
    det = diag[0] * diag[1] - np.abs(off) ** 2
    det = np.sign(det + EPS) * np.maximum(np.abs(det), EPS)
    inv_diag = np.zeros_like(diag)
    inv_diag[0] = diag[1] / det
    inv_diag[1] = diag[0] / det
    return inv_diag, -off / det, det


# --------------------------------------------------------------------------- #
# STFT / iSTFT (ref: tftransforms/stft.py:3-131, :339-394)
# --------------------------------------------------------------------------- #
def stft(data, window, hopsize, nfft, fs=44100.0):
    """Framed, windowed rfft.  ref: tftransforms/stft.py:3-69.

    N = ceil(L/hop) + 2 frames; half a window of zeros is prepended so frame 0
    is centred on sample 0; the tail is zero filled.
    """
    hopsize = int(hopsize)
    nfft = int(nfft)
    wlen = window.size
    nframes = int(np.ceil(data.size / float(hopsize)) + 2)
    total = (nframes - 1) * hopsize + wlen
    buf = np.zeros(total)
    buf[wlen // 2: wlen // 2 + data.size] = data
    idx = hopsize * np.arange(nframes)[:, None] + np.arange(wlen)[None, :]
    X = np.fft.rfft(window[None, :] * buf[idx], nfft, axis=1).T
    F = np.arange(nfft // 2 + 1) / float(nfft) * fs
    N = np.arange(nframes) * hopsize / float(fs)
    return np.ascontiguousarray(X), F, N


def istft(X, window, analysisWindow=None, hopsize=256, nfft=2048):
    """Overlap-add inverse with sum(synth*analysis) normalisation.

    ref: tftransforms/stft.py:71-131.  The first half window is dropped.
    """
    if analysisWindow is None:
        analysisWindow = window
    hopsize = int(hopsize)
    wlen = window.size
    nframes = X.shape[1]
    total = hopsize * (nframes - 1) + wlen
    frames = np.fft.irfft(X.T, int(nfft), axis=1)[:, :wlen] * window[None, :]
    data = np.zeros(total)
    norm = np.zeros(total)
    wprod = window * analysisWindow
    for n in range(nframes):  # sequential OLA keeps the reference's add order
        b = n * hopsize
        data[b:b + wlen] += frames[n]
        norm[b:b + wlen] += wprod
    data = data[wlen // 2:]
    norm = norm[wlen // 2:]
    norm[norm == 0] = 1.0
    return data / norm


class STFT(object):
    """ref: tftransforms/stft.py:339-394."""

    def __init__(self, linFTLen=2048, atomHopFactor=0.25, winFunc=np.hanning,
                 fs=44100, synthWinFunc=None, **kwargs):
        self.ftlen = linFTLen
        self.fthop = int(linFTLen * atomHopFactor)
        self.freqbins = linFTLen // 2 + 1
        winFunc = np.hanning if winFunc is None else winFunc
        self.window = winFunc(linFTLen)
        self.synthWindow = (synthWinFunc or winFunc)(linFTLen)
        self.fs = fs

    def computeTransform(self, data):
        self.transfo, self.freq_stamps, self.time_stamps = stft(
            data, self.window, self.fthop, self.ftlen, self.fs)
        self.datalen_init = data.size
        self.time_stamps = self.time_stamps * self.fs  # ref: stft.py:385

    def invertTransform(self):
        return istft(self.transfo, self.synthWindow, self.window, self.fthop,
                     self.ftlen)[:self.datalen_init]


# --------------------------------------------------------------------------- #
# audio I/O scaling (ref: audioObject.py:112-127, :130-147, :76-98)
# --------------------------------------------------------------------------- #
def median_filter(x, length=10):
    """ref: tools/signalTools.py:13-24 -- the window [n - length, min(n + length, N - 1)) is
    asymmetric and never holds the last sample; an empty window keeps the input value."""
    N = x.size
    out = np.zeros_like(x)
    for n in range(N):
        win = x[max(n - length, 0):min(n + length, N - 1)]
        out[n] = np.median(win) if win.size else np.nan
        if np.isnan(out[n]):
            out[n] = x[n]
    return out


def read_audio(filename):
    """Returns (fs, data[L, nc] float64 scaled by 1/maxdata, maxdata)."""
    fs, raw = wavfile.read(filename)
    if raw.ndim == 1:
        raw = raw[:, None]
    maxdata = np.maximum(1.1 * np.abs(raw).max(), 1e-10)
    return fs, raw / maxdata, maxdata


def pcm_from_float(ndata, maxdata):
    """QUIRK (Q10): np.int16() truncates toward zero (ref: audioModel.py:1227)."""
    return np.int16(ndata * maxdata)


def write_pcm(filename, fs, pcm):
    """ref: audioObject.py:83-98 with formatenc='pcm16' (scipy fall-back branch):
    the encoding is re-derived from the peak value."""
    peak = np.abs(pcm).max() if pcm.size else 0
    enc = "int32" if peak > 2 ** 15 else ("int16" if peak > 2 ** 7 else "int8")
    wavfile.write(filename, fs, np.array(pcm, dtype=enc))


# --------------------------------------------------------------------------- #
# FASST core
# --------------------------------------------------------------------------- #
class OracleFASST(object):
    """Restatement of `FASST` + `MultiChanNMFInst_FASST` / `MultiChanNMFConv`.

    ref: audioModel.py:66-2294 (core), :2296-2508 (model structures).  Stereo
    only, like the reference (audioModel.py:394,605,1127), unless
    `generalised=True` (default for I != 2): the E-step and the Wiener gains then
    use the batched I x I inverse of `estep_general` -- an EXTENSION with no
    reference counterpart, proven equal to the stereo path at I = 2.
    """

    def __init__(self, audio, nbComps=3, nbNMFComps=4, spatial_rank=2,
                 wlen=2048, hopsize=512, iter_num=50, sim_ann_opt="ann",
                 ann_PSD_lim=None, nmfUpdateCoeff=1.0, lambdaCorr=0.0,
                 init=True, generalised=None):
        if isinstance(audio, str):
            self.filename = audio
            self.fs, self.data, self.maxdata = read_audio(audio)
        else:  # (fs, data[L,nc] already scaled, maxdata)
            self.filename = "mix.wav"
            self.fs, self.data, self.maxdata = audio
        self.nframes_audio, self.channels = self.data.shape
        self.generalised = (self.channels != 2) if generalised is None else generalised
        self.wlen = nextpow2(wlen)  # ref: audioModel.py:192-193
        self.hopsize = hopsize
        self.tft = STFT(linFTLen=self.wlen,
                        atomHopFactor=1.0 * hopsize / self.wlen, fs=self.fs)
        self.noise = {"PSD": np.zeros(self.wlen // 2 + 1),
                      "sim_ann_opt": sim_ann_opt,
                      "ann_PSD_lim": [None, None] if ann_PSD_lim is None
                      else ann_PSD_lim}
        self.iter_num = iter_num
        self.nmfUpdateCoeff = nmfUpdateCoeff
        self.lambdaCorr = lambdaCorr
        self.spat_comps, self.spec_comps = {}, {}
        self.comp_transf_Cx()
        self.nbComps, self.nbNMFComps = nbComps, nbNMFComps
        rank = np.atleast_1d(spatial_rank)
        self.rank = [rank[0]] * nbComps if rank.size < nbComps else list(rank)
        if init:
            self._initialize_structures()

    # -- ref: audioModel.py:250-328 ----------------------------------------- #
    def comp_transf_Cx(self):
        nc = self.channels
        Xchan = []
        for c in range(nc):
            self.tft.computeTransform(self.data[:, c])
            Xchan.append(self.tft.transfo)
        self.X = np.array(Xchan)
        self.nbFreqsSigRepr, self.nbFramesSigRepr = Xchan[0].shape
        F, N = Xchan[0].shape
        self.Cx = np.zeros([nc * (nc + 1) // 2, F, N], dtype=complex)
        for n1 in range(nc):
            for n2 in range(n1, nc):
                n = n2 - n1 + int(np.sum(np.arange(nc, nc - n1, -1)))
                self.Cx[n] = Xchan[n1] * np.conj(Xchan[n2])
        lim = self.noise["ann_PSD_lim"]
        if lim[0] is None or lim[1] is None:
            mix_psd = 0
            for n1 in range(nc):
                n = int(np.sum(np.arange(nc, nc - n1, -1)))
                mix_psd = mix_psd + np.mean(self.Cx[n], axis=1)
            mix_psd = mix_psd / nc
            if lim[0] is None:
                lim[0] = np.real(mix_psd) / 100.0
            if lim[1] is None:
                lim[1] = np.real(mix_psd) / 10000.0
        if self.noise["sim_ann_opt"] in "ann":  # QUIRK: substring test (:324)
            self.noise["PSD"] = lim[0]

    # -- ref: audioModel.py:2349-2393 ---------------------------------------- #
    def _initialize_structures(self):
        nc, F, N = self.channels, self.nbFreqsSigRepr, self.nbFramesSigRepr
        K = self.nbNMFComps
        for j in range(self.nbComps):
            # F5: RNG call order randn(nc,rank); [randn(rank)]*2; randn(F,K); randn(K,N)
            params = np.random.randn(nc, self.rank[j])
            if nc == 2:
                ang = (j + 1) * np.pi / (2.0 * (self.nbComps + 1))
                params = np.array(
                    [np.sin(ang) + np.random.randn(self.rank[j]) * np.sqrt(0.01),
                     np.cos(ang) + np.random.randn(self.rank[j]) * np.sqrt(0.01)])
            self.spat_comps[j] = {"time_dep": "indep", "mix_type": "inst",
                                  "frdm_prior": "free", "params": params}
            fac = {"FB": 0.75 * np.abs(np.random.randn(F, K)) + 0.25,
                   "FW": np.eye(K),
                   "TW": 0.75 * np.abs(np.random.randn(K, N)) + 0.25,
                   "TB": [], "FB_frdm_prior": "free", "FW_frdm_prior": "fixed",
                   "TW_frdm_prior": "free", "TB_frdm_prior": [],
                   "TW_constr": "NMF"}
            self.spec_comps[j] = {"spat_comp_ind": j, "factor": {0: fac}}
        self.renormalize_parameters()

    # -- ref: audioModel.py:2488-2508 ---------------------------------------- #
    def makeItConvolutive(self):
        F = self.nbFreqsSigRepr
        for j, sc in self.spat_comps.items():
            if sc["mix_type"] != "inst":
                warnings.warn("Spatial component %d already not instantaneous,"
                              " skipping..." % j)
                continue
            inst = sc["params"]
            sc["mix_type"] = "conv"
            sc["params"] = np.zeros([self.rank[j], self.channels, F],
                                    dtype=complex)
            sc["params"][:] = inst.T[:, :, None]

    # -- ref: audioModel.py:330-382 ------------------------------------------ #
    def estim_param_a_post_model(self):
        logliks = np.ones(self.iter_num)
        opt, lim = self.noise["sim_ann_opt"], self.noise["ann_PSD_lim"]
        if opt in ["ann"]:
            self.noise["PSD"] = lim[0]
        elif opt == "no_ann":
            self.noise["PSD"] = lim[1]
        else:
            warnings.warn("To add noise to the signal, provide the sim_ann_opt"
                          " from any of 'ann', 'no_ann' or 'ann_ns_inj' ")
        I = self.iter_num
        for i in range(I):
            if opt in ["ann", "ann_ns_inj"]:
                # QUIRK (Q8): never reaches lim[1] since i <= I-1
                self.noise["PSD"] = ((np.sqrt(lim[0]) * (I - i)
                                      + np.sqrt(lim[1]) * i) / I) ** 2
            logliks[i] = self.GEM_iteration()
            # multiChanSourceF0Filter.estim_param_a_post_model (audioModel.py:2933-2979): the
            # source activations are re-weighted after every iteration when a component
            # carries a 'sparsity' entry; sigma goes from K^2 down to 9 geometrically
            if any(sp.get("sparsity") for sp in self.spec_comps.values()):
                log0 = np.log(np.max([sp["factor"][0]["TW"].shape[0]
                                      for sp in self.spec_comps.values()]) ** 2)
                sigma = np.exp(log0 + (np.log(9.0) - log0) / max(I - 1.0, 1.) * i)
                self.reweigh_sparsity_constraint(sigma)
        return logliks

    # -- ref: audioModel.py:2981-3014, tools/signalTools.py:13-24 ---------------- #
    def reweigh_sparsity_constraint(self, sigma):
        for sp in self.spec_comps.values():
            TW = sp["factor"][0]["TW"]
            K = TW.shape[0]
            if not sp.get("sparsity") or K <= 2:
                continue
            w = np.arange(K - 1, 0, -1) ** 2
            mu = np.dot(np.arange(K - 1) * w, TW[:-1]) / np.dot(w, np.maximum(TW[:-1], EPS))
            mu = median_filter(mu, length=sp["sparsity"])
            mask = np.exp(-0.5 * ((np.vstack(np.arange(K)) - mu) ** 2) / sigma)
            mask[-1] = mask.max(axis=0)
            pos = mask[-1] > 0
            mask[:, pos] /= mask[-1][pos]
            TW *= mask

    # -- ref: audioModel.py:384-428 ------------------------------------------ #
    def GEM_iteration(self):
        if self.channels != 2 and not self.generalised:
            raise AttributeError("Nb channels %d not implemented yet"
                                 % self.channels)
        powers, mix, ranks = self.retrieve_subsrc_params()
        _, hat_Rxs, hat_Rss, hat_Ws, loglik = self.compute_suff_stat(powers, mix)
        self.update_mix_matrix(hat_Rxs, hat_Rss, mix, ranks)
        hat_W = np.array([np.mean(hat_Ws[ranks[w]], axis=0)
                          for w in range(len(ranks))])
        self.update_spectral_components(hat_W)
        self.renormalize_parameters()
        return float(np.real(loglik))  # Q4: complex mean cast to float (:376)

    # -- ref: audioModel.py:430-498 ------------------------------------------ #
    def comp_spat_comp_power(self, spat_comp_ind, spec_comp_ind=(),
                             factor_ind=()):
        V = np.zeros([self.nbFreqsSigRepr, self.nbFramesSigRepr])
        specs = spec_comp_ind if len(spec_comp_ind) else list(self.spec_comps)
        for k in specs:
            if self.spec_comps[k]["spat_comp_ind"] != spat_comp_ind:
                continue
            facs = self.spec_comps[k]["factor"]
            # QUIRK (Q1): an empty factor list means *all* factors (:478-481)
            which = factor_ind if len(factor_ind) else list(facs)
            Vc = np.ones_like(V)
            for f in which:
                fac = facs[f]
                W = np.dot(fac["FB"], fac["FW"])
                H = np.dot(fac["TW"], fac["TB"]) if len(fac["TB"]) else fac["TW"]
                Vc *= np.dot(W, H)
            V += Vc
        return V

    # -- ref: audioModel.py:514-578 ------------------------------------------ #
    def retrieve_subsrc_params(self):
        ranks, total = {}, 0
        for j in range(len(self.spat_comps)):
            sc = self.spat_comps[j]
            r = sc["params"].shape[1 if sc["mix_type"] == "inst" else 0]
            ranks[j] = total + np.arange(r)
            total += r
        F, N = self.nbFreqsSigRepr, self.nbFramesSigRepr
        powers = np.zeros([total, F, N])
        mix = np.zeros([total, self.channels, F], dtype=complex)
        for j, sc in self.spat_comps.items():
            powers[ranks[j]] = self.comp_spat_comp_power(j)[None]
            if sc["mix_type"] == "inst":
                mix[ranks[j]] = sc["params"].T[:, :, None]
            else:
                mix[ranks[j]] = sc["params"]
        return powers, mix, ranks

    # -- ref: audioModel.py:580-764 (E-step) --------------------------------- #
    def compute_suff_stat(self, spat_comp_powers, mix_matrix):
        if self.generalised:
            # EXTENSION (no reference for I != 2): the same definitions with a batched I x I
            # inverse; identical to the stereo path at I = 2 (tests/test_oracle_golden.py)
            hat_Rxs, hat_Rss, hat_Ws, loglik = estep_general(
                self.X, spat_comp_powers, mix_matrix, self.noise["PSD"])
            return None, hat_Rxs, hat_Rss, hat_Ws, loglik
        if self.channels != 2:
            raise ValueError("Nb channels not supported:%d" % self.channels)
        R = spat_comp_powers.shape[0]
        F, N = self.nbFreqsSigRepr, self.nbFramesSigRepr
        V, A, Cx = spat_comp_powers, mix_matrix, self.Cx
        col = lambda a: a[:, None]  # np.vstack on a 1-D array
        # Sigma_x = sum_r a_r a_r^H v_r + noise (:613-652); r=0 first, then noise
        sd = np.empty([2, F, N])
        sd[0] = col(np.abs(A[0][0]) ** 2) * V[0]
        sd[1] = col(np.abs(A[0][1]) ** 2) * V[0]
        so = col(A[0][0] * np.conj(A[0][1])) * V[0]
        sd += col(self.noise["PSD"])[None]
        for r in range(1, R):
            sd[0] += col(np.abs(A[r][0]) ** 2) * V[r]
            sd[1] += col(np.abs(A[r][1]) ** 2) * V[r]
            so += col(A[r][0] * np.conj(A[r][1])) * V[r]
        idg, iof, det = inv_herm_mat_2d(sd, so)
        # QUIRK (Q4): log(det*pi), not pi^2; complex mean (:660-664)
        loglik = -np.mean(np.log(det * np.pi) + idg[0] * Cx[0] + idg[1] * Cx[2]
                          + 2.0 * np.real(iof * np.conj(Cx[1])))
        # Wiener gains G[c, r] = v_r (a_r^H Sigma^-1)[c] (:666-684)
        G = np.empty([2, R, F, N], dtype=complex)
        for r in range(R):
            a0c, a1c = col(np.conj(A[r][0])), col(np.conj(A[r][1]))
            G[0, r] = (a0c * idg[0] + a1c * np.conj(iof)) * V[r]
            G[1, r] = (a0c * iof + a1c * idg[1]) * V[r]
        hat_Rss = np.empty([F, R, R], dtype=complex)
        hat_Ws = np.empty([R, F, N])
        for r1 in range(R):
            for r2 in range(R):
                # G_r1 Cx G_r2^H - G_r1 a_r2 v_r2 (+ v_r1 on the diagonal) (:698-731)
                t = G[0, r1] * (Cx[0] * np.conj(G[0, r2])
                                + np.conj(G[1, r2]) * Cx[1])
                t += G[1, r1] * (Cx[2] * np.conj(G[1, r2])
                                 + np.conj(G[0, r2] * Cx[1]))
                t -= (G[0, r1] * col(A[r2, 0]) + G[1, r1] * col(A[r2, 1])) * V[r2]
                if r1 == r2:
                    t += V[r1]
                    hat_Ws[r1] = np.abs(np.real(t))
                hat_Rss[:, r1, r2] = np.mean(t, axis=1)
        hat_Rss = 0.5 * (hat_Rss + np.conj(np.transpose(hat_Rss, (0, 2, 1))))
        hat_Rxs = np.empty([F, 2, R], dtype=complex)
        for r in range(R):
            g0c, g1c = np.conj(G[0, r]), np.conj(G[1, r])
            hat_Rxs[:, 0, r] = np.mean(g0c * Cx[0] + g1c * Cx[1], axis=1)
            hat_Rxs[:, 1, r] = np.mean(g0c * np.conj(Cx[1]) + g1c * Cx[2],
                                       axis=1)
        hat_Rxx = np.mean(Cx, axis=-1)
        return hat_Rxx, hat_Rxs, hat_Rss, hat_Ws, loglik

    # -- ref: audioModel.py:766-889 (spatial M-step) -------------------------- #
    def update_mix_matrix(self, hat_Rxs, hat_Rss, mix_matrix, rank_part_ind):
        F = self.nbFreqsSigRepr
        sel = {"inst": ([], []), "conv": ([], [])}
        for j, sc in self.spat_comps.items():
            for kind in ("inst", "conv"):
                hit = sc["frdm_prior"] == "free" and sc["mix_type"] == kind
                sel[kind][0 if hit else 1].extend(rank_part_ind[j])
        upd, oth = sel["inst"]
        if len(upd):
            rxs = hat_Rxs[:, :, upd]
            if len(oth):
                for f in range(F):
                    rxs[f] -= np.dot(mix_matrix[oth, :, f].T,
                                     hat_Rss[f][np.ix_(oth, upd)])
            # QUIRK (Q6): real part of the f-averaged statistics (:824-830)
            rxs = np.real(np.mean(rxs, axis=0))
            rss = np.real(np.mean(hat_Rss[:, np.vstack(upd), upd], axis=0))
            sol = np.linalg.solve(rss.T, rxs.T)
            mix_matrix[upd] = sol[:, :, None]
        upd, oth = sel["conv"]
        if len(upd):
            rxs = hat_Rxs[:, :, upd]
            if len(oth):
                for f in range(F):
                    rxs[f] -= np.dot(mix_matrix[oth, :, f].T,
                                     hat_Rss[f][np.ix_(oth, upd)])
            for f in range(F):
                # QUIRK (Q7): solves with the full hat_Rss[f] (:857)
                try:
                    mix_matrix[upd, :, f] = np.linalg.solve(hat_Rss[f].T,
                                                            rxs[f].T)
                except np.linalg.LinAlgError:
                    raise np.linalg.LinAlgError("Singular Matrix")
        for k, sc in self.spat_comps.items():
            if sc["frdm_prior"] != "free":
                continue
            if sc["mix_type"] == "inst":
                sc["params"] = np.mean(mix_matrix[rank_part_ind[k]], axis=2).T
            else:
                sc["params"] = mix_matrix[rank_part_ind[k]]

    # -- ref: audioModel.py:1469-1727, :1931-1978 (spectral M-step) ----------- #
    def update_spectral_components(self, hat_W):
        om = self.nmfUpdateCoeff
        if self.lambdaCorr > 0:
            raise NotImplementedError("oracle: lambdaCorr > 0 not restated")
        for s, spec in self.spec_comps.items():
            nfac = len(spec["factor"])
            j = spec["spat_comp_ind"]
            for fi, fac in spec["factor"].items():
                others = [x for x in range(nfac) if x != fi]
                # QUIRK (Q1, Q2): own power when single factor; computed once
                other = np.maximum(self.comp_spat_comp_power(j, [s], others),
                                   EPS)
                H0 = lambda: (np.dot(fac["TW"], fac["TB"]) if len(fac["TB"])
                              else fac["TW"])
                if fac["FB_frdm_prior"] == "free":
                    # QUIRK (Q3): power of *all* spec comps of this spat comp
                    P = np.maximum(self.comp_spat_comp_power(j), EPS)
                    WH = np.dot(fac["FW"], H0()).T
                    den = np.dot(other * (1.0 / P), WH)
                    num = np.dot(hat_W[j] / P ** 2 * other, WH)
                    fac["FB"] *= (num / np.maximum(den, EPS)) ** om
                if fac["FW_frdm_prior"] == "free":
                    P = np.maximum(self.comp_spat_comp_power(j, [s]), EPS)
                    H = H0()
                    den = np.dot(fac["FB"].T, np.dot(other * (1.0 / P), H.T))
                    num = np.dot(fac["FB"].T,
                                 np.dot(hat_W[j] / P ** 2 * other, H.T))
                    fac["FW"] *= (num / np.maximum(den, EPS)) ** om
                if fac["TW_frdm_prior"] == "free":
                    if fac["TW_constr"] != "NMF":
                        raise NotImplementedError("discrete-state TW (:1728)")
                    P = np.maximum(self.comp_spat_comp_power(j, [s]), EPS)
                    W = np.dot(fac["FB"], fac["FW"])
                    dplane = other * (1.0 / P)
                    nplane = other * (hat_W[j] / P ** 2)
                    if len(fac["TB"]):
                        dplane = np.dot(dplane, fac["TB"].T)
                        nplane = np.dot(hat_W[j] / P ** 2 * other, fac["TB"].T)
                    den, num = np.dot(W.T, dplane), np.dot(W.T, nplane)
                    fac["TW"] *= (num / np.maximum(den, EPS)) ** om
                if len(fac["TB"]) and fac["TB_frdm_prior"] == "free":
                    P = np.maximum(self.comp_spat_comp_power(j, [s]), EPS)
                    W = np.dot(np.dot(fac["FB"], fac["FW"]), fac["TW"])
                    den = np.dot(W.T, other * (1.0 / P))
                    num = np.dot(W.T, hat_W[j] / np.maximum(P ** 2, EPS) * other)
                    fac["TB"] *= (num / np.maximum(den, EPS)) ** om

    # -- ref: audioModel.py:1980-2040 ------------------------------------------ #
    def renormalize_parameters(self):
        energy = np.zeros(len(self.spat_comps))
        for j, sc in self.spat_comps.items():
            energy[j] = np.mean(np.abs(sc["params"]) ** 2)
            sc["params"] = sc["params"] / np.sqrt(energy[j])
        for s, spec in self.spec_comps.items():
            g = energy[spec["spat_comp_ind"]]
            nfac = len(spec["factor"])
            for fi, fac in spec["factor"].items():
                if fac["TW_constr"] in ("GMM", "HMM"):
                    raise NotImplementedError(
                        "Temporal discrete state mngmt not done yet. ")
                fac["FB"] *= g
                w = fac["FB"].max(axis=0)
                w[w == 0] = 1.0
                fac["FB"] /= w
                fac["FW"] *= w[:, None]
                w = fac["FW"].mean(axis=0)
                w[w == 0] = 1.0
                fac["FW"] /= w
                fac["TW"] *= w[:, None]
                if np.sum(fac["TW"]) < EPS:  # random restart (:2023-2025)
                    fac["TW"] = np.random.randn(*fac["TW"].shape) ** 2
                    fac["TW"] *= 1e3 * EPS
                if len(fac["TB"]):
                    w = fac["TB"].mean(axis=1)
                    w[w == 0] = 1.0
                    fac["TB"] /= w[:, None]
                    fac["TW"] *= w
                g = fac["TW"].mean()
                if fi < nfac - 1:
                    fac["TW"] /= g

    # -- ref: audioModel.py:1063-1236, :1327-1467 (Wiener separation) ---------- #
    def separation_gains(self, spec_comp_ind=None):
        """Returns WG[nsrc, 2, 2, F, N] (complex) -- compute_sigma_comp_2d,
        compute_inv_sigma_mix_2d, compute_Wiener_gain_2d."""
        if spec_comp_ind is None:
            spec_comp_ind = {j: [] for j in range(len(self.spat_comps))}
            for s, spec in self.spec_comps.items():
                spec_comp_ind[spec["spat_comp_ind"]].append(s)
        nsrc = len(spec_comp_ind)
        F, N = self.nbFreqsSigRepr, self.nbFramesSigRepr
        sdiag = np.zeros([nsrc, 2, F, N])
        soff = np.zeros([nsrc, F, N], dtype=complex)
        for n in range(nsrc):
            spats = np.unique([self.spec_comps[s]["spat_comp_ind"]
                               for s in spec_comp_ind[n]])
            for j in spats:
                V = self.comp_spat_comp_power(j, spec_comp_ind[n])
                sc = self.spat_comps[j]
                mc = sc["params"].T if sc["mix_type"] == "inst" else sc["params"]
                R0 = np.atleast_1d((np.abs(mc[:, 0]) ** 2).sum(axis=0))
                R1 = np.atleast_1d((np.abs(mc[:, 1]) ** 2).sum(axis=0))
                Ro = np.atleast_1d((mc[:, 0] * np.conj(mc[:, 1])).sum(axis=0))
                sdiag[n, 0] += R0[:, None] * V
                sdiag[n, 1] += R1[:, None] * V
                soff[n] += Ro[:, None] * V
        sx = sdiag.sum(axis=0)
        sxo = soff.sum(axis=0)
        sx += self.noise["PSD"][None, :, None]  # last-iteration PSD (Q8, :1385)
        idg, iof, _ = inv_herm_mat_2d(sx, sxo)
        WG = np.zeros([nsrc, 2, 2, F, N], dtype=complex)
        for n in range(nsrc):
            WG[n, 0, 0] = soff[n] * np.conj(iof)
            WG[n, 1, 1] = np.conj(WG[n, 0, 0])
            WG[n, 0, 0] += sdiag[n, 0] * idg[0]
            WG[n, 1, 1] += sdiag[n, 1] * idg[1]
            WG[n, 0, 1] = sdiag[n, 0] * iof + soff[n] * idg[1]
            WG[n, 1, 0] = np.conj(soff[n]) * idg[0] + sdiag[n, 1] * np.conj(iof)
        return WG

    def separation_gains_general(self, spec_comp_ind=None):
        """EXTENSION: WG[nsrc, I, I, F, N] = Sigma_n Sigma_x^-1 with a batched inverse (the clamp of
        the determinant as in estep_general); equals separation_gains at I = 2."""
        if spec_comp_ind is None:
            spec_comp_ind = {j: [] for j in range(len(self.spat_comps))}
            for s, spec in self.spec_comps.items():
                spec_comp_ind[spec["spat_comp_ind"]].append(s)
        nsrc, I = len(spec_comp_ind), self.channels
        F, N = self.nbFreqsSigRepr, self.nbFramesSigRepr
        Sn = np.zeros([nsrc, F, N, I, I], dtype=complex)
        for n in range(nsrc):
            spats = np.unique([self.spec_comps[s]["spat_comp_ind"] for s in spec_comp_ind[n]])
            for j in spats:
                V = self.comp_spat_comp_power(j, spec_comp_ind[n])
                sc = self.spat_comps[j]
                mc = sc["params"].T if sc["mix_type"] == "inst" else sc["params"]  # [rank, I(, F)]
                mc = mc[:, :, None] * np.ones(F) if mc.ndim == 2 else mc
                Rj = np.einsum("raf,rbf->fab", mc, np.conj(mc))
                Sn[n] += V[:, :, None, None] * Rj[:, None]
        Sx = Sn.sum(axis=0) + self.noise["PSD"][:, None, None, None] * np.eye(I)
        det = np.real(np.linalg.det(Sx))
        detc = np.sign(det + EPS) * np.maximum(np.abs(det), EPS)
        Sinv = np.linalg.inv(Sx) * (det / detc)[..., None, None]
        WG = np.einsum("sfnab,fnbc->sacfn", Sn, Sinv)
        return WG

    def separate_signals(self, spec_comp_ind=None):
        """Float separated signals [nsrc, L, I] before PCM conversion."""
        if self.generalised:
            WG = self.separation_gains_general(spec_comp_ind)
            outs = []
            for n in range(WG.shape[0]):
                chans = []
                for c1 in range(self.channels):
                    self.tft.transfo = sum(WG[n, c1, c2] * self.X[c2]
                                           for c2 in range(self.channels))
                    self.tft.datalen_init = self.nframes_audio
                    chans.append(self.tft.invertTransform())
                outs.append(np.array(chans).T)
            return np.array(outs)
        WG = self.separation_gains(spec_comp_ind)
        outs = []
        for n in range(WG.shape[0]):
            chans = []
            for c1 in range(2):
                self.tft.transfo = WG[n, c1, 0] * self.X[0] + WG[n, c1, 1] * self.X[1]
                self.tft.datalen_init = self.nframes_audio
                chans.append(self.tft.invertTransform())
            outs.append(np.array(chans).T)
        return np.array(outs)

    def separate_spat_comps(self, dir_results=None, suffix=None):
        sig = self.separate_signals()
        nsrc = sig.shape[0]
        if dir_results is None:
            dir_results = "/".join(self.filename.split("/")[:-1])
        root = self.filename.split("/")[-1][:-4]
        self.files = {"spat_comp": []}
        pcm_all = []
        for n in range(nsrc):
            sfx = "_" + suffix[n] if (suffix is not None and n in suffix) else ""
            name = "%s/%s_%d-%d%s.wav" % (dir_results, root, n, nsrc, sfx)
            pcm = pcm_from_float(sig[n][:self.nframes_audio], self.maxdata)
            write_pcm(name, self.fs, pcm)
            self.files["spat_comp"].append(name)
            pcm_all.append(pcm)
        return pcm_all


# --------------------------------------------------------------------------- #
# generalised-I E-step (no reference implementation for I != 2).
# Validated against `compute_suff_stat` at I = 2 in tests/test_oracle_golden.py.
# --------------------------------------------------------------------------- #
def estep_general(X, V_sub, A, noise_psd):
    """X[I,F,N] complex, V_sub[R,F,N], A[R,I,F] complex, noise_psd[F].

    Same definitions as audioModel.py:580-764 with a batched np.linalg.inv
    instead of the 2x2 closed form (the determinant clamp Q5 is applied to the
    generic determinant in the same way).  Returns hat_Rxs[F,I,R],
    hat_Rss[F,R,R], hat_Ws[R,F,N], loglik.
    """
    I, F, N = X.shape
    R = V_sub.shape[0]
    Af = np.transpose(A, (2, 1, 0))  # [F, I, R]
    Sig = np.einsum("fir,rfn,fjr->fnij", Af, V_sub, np.conj(Af))
    Sig = Sig + (noise_psd[:, None, None, None]
                 * np.eye(I)[None, None])
    det = np.real(np.linalg.det(Sig))
    detc = np.sign(det + EPS) * np.maximum(np.abs(det), EPS)
    Sinv = np.linalg.inv(Sig) * (det / detc)[..., None, None]
    xs = np.transpose(X, (1, 2, 0))  # [F,N,I]
    y = np.einsum("fnij,fnj->fni", Sinv, xs)
    quad = np.real(np.einsum("fni,fni->fn", np.conj(xs), y))
    loglik = -np.mean(np.log(detc * np.pi) + quad)
    M = y[..., :, None] * np.conj(y[..., None, :]) - Sinv  # [F,N,I,I]
    # hat_Rss[f,r1,r2] = mean_n v_r1 v_r2 a_r1^H M a_r2 + delta mean_n v_r1
    aMa = np.einsum("fir,fnij,fjs->fnrs", np.conj(Af), M, Af)
    Vt = np.transpose(V_sub, (1, 2, 0))  # [F,N,R]
    full = Vt[..., :, None] * Vt[..., None, :] * aMa
    hat_Ws = np.abs(np.real(np.einsum("fnrr->rfn", full)) + V_sub)
    hat_Rss = np.mean(full, axis=1)
    hat_Rss[:, np.arange(R), np.arange(R)] += np.mean(V_sub, axis=2).T
    hat_Rss = 0.5 * (hat_Rss + np.conj(np.transpose(hat_Rss, (0, 2, 1))))
    # hat_Rxs[f,:,r] = mean_n v_r x y^H a_r
    ya = np.einsum("fni,fir->fnr", np.conj(y), Af)
    hat_Rxs = np.mean(xs[..., :, None] * (Vt * ya)[..., None, :], axis=1)
    return hat_Rxs, hat_Rss, hat_Ws, loglik
