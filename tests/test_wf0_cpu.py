"""The glottal-source F0 dictionary on the CPU: the oracle against the golden
vectors made by the reference (oracle/make_golden.py: run_wf0), and the host side
(pyfasst_b200 separateLeadFunctions.generate_WF0_chirped / generate_WF0_TR_chirped with the NumPy
specification of the comb kernel) against both, cache files included."""
import os

import numpy as np
import pytest

from oracle import wf0_oracle as wo
from pyfasst_b200.SeparateLeadStereo import separateLeadFunctions as slf
from pyfasst_b200.tftransforms.stft import STFT
from pyfasst_b200.tools.utils import sqrt_blackmanharris
from tests.fake_simm_kernels import FakeSimmKernels

GOLD = os.path.join(os.path.dirname(__file__), "golden", "wf0.npz")


def relmax(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_against_reference(gold):
    t, w = wo.generate_WF0_TR_chirped(256, 64, sqrt_blackmanharris, 8000, 100, 800, 2, 0.5, 1, 0.5)
    assert np.array_equal(t, gold["t1"]) and relmax(w, gold["w1"]) < 1e-12
    t, w = wo.generate_WF0_TR_chirped(256, 64, sqrt_blackmanharris, 8000, 100, 800, 1, 0.5, 3, 0.5)
    assert np.array_equal(t, gold["t2"]) and relmax(w, gold["w2"]) < 1e-12
    t, w = wo.generate_WF0_chirped(100, 800, 8000, 256, 1, 256, 0.5, 2, .15, 'sinebell')
    assert np.array_equal(t, gold["t3"]) and relmax(w, gold["w3"]) < 1e-12
    t, w = wo.generate_WF0_TR_chirped(512, 64, np.hanning, 16000, 60, 500, 1, 0.25, 2, 0.5)
    assert np.array_equal(t, gold["t4"]) and relmax(w, gold["w4"]) < 1e-12


def test_host_side_against_reference(gold, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    fk = FakeSimmKernels()
    tr = STFT(linFTLen=256, atomHopFactor=0.25, winFunc=sqrt_blackmanharris, fs=8000, kernels=fk)
    t, w, tr_out = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=2, Ot=0.5,
                                               perF0=1, depthChirpInSemiTone=0.5, kernels=fk)
    assert tr_out is tr
    assert np.array_equal(t, gold["t1"]) and relmax(w, gold["w1"]) < 1e-10
    t, w, _ = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=1, Ot=0.5, perF0=3,
                                          depthChirpInSemiTone=0.5, kernels=fk)
    assert np.array_equal(t, gold["t2"]) and relmax(w, gold["w2"]) < 1e-10
    t, w = slf.generate_WF0_chirped(100, 800, 8000, Nfft=256, stepNotes=1, lengthWindow=256,
                                    Ot=0.5, perF0=2, depthChirpInSemiTone=.15,
                                    analysisWindow='sinebell', kernels=fk)
    assert np.array_equal(t, gold["t3"]) and relmax(w, gold["w3"]) < 1e-10
    tr2 = STFT(linFTLen=512, atomHopFactor=0.125, winFunc=np.hanning, fs=16000, kernels=fk)
    t, w, _ = slf.generate_WF0_TR_chirped(tr2, minF0=60, maxF0=500, stepNotes=1, Ot=0.25, perF0=2,
                                          depthChirpInSemiTone=0.5, kernels=fk)
    assert np.array_equal(t, gold["t4"]) and relmax(w, gold["w4"]) < 1e-10
    # the cache files carry the reference's names, and are read back instead of recomputed
    assert sorted(os.listdir(".")) == sorted(str(n) for n in gold["cache_names"])
    launches = fk.launches
    t, w, _ = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=2, Ot=0.5, perF0=1,
                                          depthChirpInSemiTone=0.5, kernels=fk)
    assert fk.launches == launches and relmax(w, gold["w1"]) < 1e-10


def test_plain_and_chirped_windows_differ_like_the_reference(tmp_path, monkeypatch):
    """generate_WF0_chirped windows the plain combs with `analysisWindow` and the chirped ones
    with the sinebell default of generate_ODGD_spec_chirped (separateLeadFunctions.py:335-339)."""
    monkeypatch.chdir(tmp_path)
    fk = FakeSimmKernels()
    t, w = slf.generate_WF0_chirped(100, 400, 8000, Nfft=256, stepNotes=1, lengthWindow=256,
                                    Ot=0.5, perF0=2, depthChirpInSemiTone=.15,
                                    analysisWindow='hanning', kernels=fk)
    t2, w2 = wo.generate_WF0_chirped(100, 400, 8000, 256, 1, 256, 0.5, 2, .15, 'hanning')
    assert np.array_equal(t, t2) and relmax(w, w2) < 1e-10
    with pytest.raises(ValueError):
        slf.generate_WF0_chirped(100, 400, 8000, Nfft=256, lengthWindow=256,
                                 analysisWindow='blackman', kernels=fk)
