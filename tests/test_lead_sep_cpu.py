"""The SIMM front / back end (pyfasst_b200/SeparateLeadStereo: separateLeadFunctions.stft / istft /
generateHannBasis, SeparateLeadProcess.computeStereoX / estimStereoSIMMParams /
writeSeparatedSignals) on the NumPy specification of the kernels, against golden vectors made
by executing the reference (tests/golden/lead_sep.npz, oracle/make_golden.py: run_lead_sep).
CPU only; the `-m gpu` twin is tests/test_lead_sep_gpu.py."""
import os
import shutil

import numpy as np
import pytest
import scipy.io.wavfile as wavfile
from numpy.testing import assert_allclose

from oracle import simm_oracle as so
from pyfasst_b200.SeparateLeadStereo import SeparateLeadStereoTF as sls
from pyfasst_b200.SeparateLeadStereo import separateLeadFunctions as slf
from tests.fake_kernels import FakeKernels
from tests.fake_simm_kernels import FakeSimmKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FS, WLEN, HOP = 8000, 256, 32


class AllFakeKernels(FakeSimmKernels, FakeKernels):
    """STFT / iSTFT stand-ins of fake_kernels.py + the SIMM stand-ins."""


def load():
    return np.load(os.path.join(GOLDEN, "lead_sep.npz"))


def make_process(tmp_path, kernels, g=None, **kw):
    wav = os.path.join(str(tmp_path), "mix_lead.wav")
    shutil.copy(os.path.join(GOLDEN, "mix_lead.wav"), wav)
    g = load() if g is None else g
    return sls.SeparateLeadProcess(wav, windowSize=WLEN / float(FS), hopsize=HOP, nbIter=3,
                                   numCompAccomp=4, K_numFilters=3, P_numAtomFilters=8,
                                   WF0=g["p_WF0"], verbose=False, kernels=kernels, **kw)


def check_stft_istft(kernels):
    g = load()
    data = np.double(g["pcm"]) / (1.2 * np.abs(g["pcm"]).max())
    X, F, N = slf.stft(data[:, 0], window=slf.sinebell(WLEN), hopsize=HOP, nfft=WLEN, fs=FS,
                       kernels=kernels)
    assert X.shape == g["XR"].shape
    assert_allclose(X, g["XR"], rtol=0, atol=1e-10)
    assert_allclose(F, g["F"], rtol=1e-12)
    assert_allclose(N, g["N"], rtol=1e-12)
    # a range of frames (chunked processing, SeparateLeadStereoTF.py:761-841)
    Xc, _, _ = slf.stft(data[:, 0], window=slf.sinebell(WLEN), hopsize=HOP, nfft=WLEN, fs=FS,
                        start=10, stop=50, kernels=kernels)
    assert_allclose(Xc, g["XR"][:, 10:50], rtol=0, atol=1e-10)
    y = slf.istft(g["XR"], window=slf.sinebell(WLEN), hopsize=HOP, nfft=WLEN, kernels=kernels)
    assert y.shape == g["yR"].shape
    assert_allclose(y, g["yR"], rtol=0, atol=1e-10)


def check_separation(tmp_path, kernels):
    g = load()
    proc = make_process(tmp_path, kernels, g)
    assert proc.stftParams["windowSizeInSamples"] == WLEN and proc.F == WLEN // 2 + 1
    assert proc.scaleData == 1.2 * np.abs(g["pcm"]).max()
    for nm in ("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR", "betaL", "HM", "WM",
               "WGAMMA"):
        proc.SIMMParams[nm] = g["p_" + nm]
    proc.SIMMParams["alphaR"] = float(g["p_alphaR"])
    proc.SIMMParams["alphaL"] = float(g["p_alphaL"])
    proc.computeStereoX()
    assert_allclose(proc.XR, g["XR"], rtol=0, atol=1e-10)
    assert_allclose(proc.XL, g["XL"], rtol=0, atol=1e-10)
    proc.writeSeparatedSignals()
    for key, ref in (("voc_output_file", g["voc"]), ("mus_output_file", g["mus"])):
        fs, got = wavfile.read(proc.files[key])
        assert fs == FS and got.dtype == np.int16 and got.shape == ref.shape
        # float32 masks against the float64 reference, then rounding to int16
        assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1, key
        assert np.mean(got != ref) < 0.02, key
    assert proc.files["voc_output_file"].endswith("mix_lead_lead.wav")
    assert proc.files["mus_output_file"].endswith("mix_lead_acc.wav")


def check_estimation(tmp_path, kernels):
    g = load()
    proc = make_process(tmp_path, kernels, g)
    F, N = g["XR"].shape
    P, K, R, NF0 = 8, 3, 4, g["p_WF0"].shape[1]
    proc.SIMMParams["WGAMMA"] = g["p_WGAMMA"]
    proc.SIMMParams["HF00"] = g["HF00"]
    np.random.seed(4)
    proc.estimStereoSIMMParams()
    # the same draws, in the reference's order (SIMM.py:523-581), for the oracle
    np.random.seed(4)
    HG0, HPHI0 = np.abs(np.random.randn(P, K)), np.abs(np.random.randn(K, N))
    HM0, WM0 = np.abs(np.random.randn(R, N)), np.abs(np.random.randn(F, R))
    beta0 = np.random.rand(R)
    ref = so.stereo_simm(np.abs(g["XR"]) ** 2, np.abs(g["XL"]) ** 2, g["p_WF0"], g["p_WGAMMA"],
                         HG0, HPHI0, g["HF00"], WM0, HM0, beta0, numberOfIterations=3)
    names = ("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR", "betaL", "HM", "WM")
    for nm, b in zip(names, ref):
        a = np.asarray(proc.SIMMParams[nm])
        assert np.abs(a - b).max() / np.abs(b).max() < 3e-4, nm


def test_stft_istft_match_reference():
    check_stft_istft(AllFakeKernels())


def test_hann_basis_matches_reference():
    g = load()
    assert_allclose(slf.generateHannBasis(1025, 2048, 44100, numberOfBasis=30, overlap=.75),
                    g["hann_1025_30"], rtol=0, atol=1e-14)
    assert_allclose(slf.generateHannBasis(129, 256, 8000, numberOfBasis=8), g["hann_129_8"],
                    rtol=0, atol=1e-14)
    with pytest.raises(NotImplementedError):
        slf.generateHannBasis(129, 256, 8000, frequencyScale='log')


def test_write_separated_signals_matches_reference(tmp_path):
    check_separation(tmp_path, AllFakeKernels())


def test_estim_stereo_simm_params(tmp_path):
    check_estimation(tmp_path, AllFakeKernels())


def test_constructor_builds_the_f0_dictionary(tmp_path, monkeypatch):
    """Without a WF0 argument the constructor generates the glottal F0 dictionary like the
    reference's computeWF0 (SeparateLeadStereoTF.py:661-684): STFT transform object with the
    sqrt-Blackman-Harris window, combs normalised to sum one."""
    from oracle import wf0_oracle as wo
    from pyfasst_b200.tools.utils import sqrt_blackmanharris
    monkeypatch.chdir(tmp_path)
    p = make_process(tmp_path, AllFakeKernels(), dict(p_WF0=None), minF0=100, maxF0=800,
                     stepNotes=2)
    table, W = wo.generate_WF0_TR_chirped(WLEN, WLEN // 4, sqrt_blackmanharris, FS, 100, 800, 2,
                                          0.5, 1, 0.5)
    assert_allclose(p.SIMMParams['F0Table'], table, rtol=0, atol=0)
    assert_allclose(p.SIMMParams['WF0'], wo.normalise(W), rtol=0, atol=1e-12)
    assert p.SIMMParams['NF0'] == table.size and p.F == WLEN // 2 + 1


def test_constructor_errors(tmp_path):
    g = load()
    with pytest.raises(ValueError):
        make_process(tmp_path, AllFakeKernels(), dict(p_WF0=g["p_WF0"][:-1]))
    with pytest.raises(NotImplementedError):
        make_process(tmp_path, AllFakeKernels(), g, tfrepresentation='minqt')
