"""GPU twin of tests/test_wf0_cpu.py: the comb kernel (csrc/wf0.cu) against its NumPy
specification, the dictionary generators against the golden vectors made by the reference, and
SeparateLeadProcess.computeWF0 at the reference's default size against a property of the
dictionary (every comb peaks at its own fundamental)."""
import os
import time

import numpy as np
import pytest

from oracle import wf0_oracle as wo
from pyfasst_b200.SeparateLeadStereo import separateLeadFunctions as slf
from pyfasst_b200.tftransforms.stft import STFT
from pyfasst_b200.tools.utils import sqrt_blackmanharris
from tests.fake_simm_kernels import FakeSimmKernels

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "wf0.npz")


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


def relmax(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("wlen,nfft,rows,tb", [(256, 256, 129, 64), (200, 256, 256, 0),
                                               (512, 1024, 513, -100), (64, 64, 33, 700)])
def test_comb_kernel_against_spec(wlen, nfft, rows, tb):
    rng = np.random.default_rng(wlen + rows)
    fs, Lsig, n = 8000.0, 800, 37
    f0 = 60.0 * 2 ** (rng.random(n) * 5)
    f1 = f0 * (1 - 0.02 * rng.random(n) * (rng.random(n) < 0.5))
    f2 = 2 * f0 - f1
    f1[-1] = f2[-1] = 5000.0  # above Nyquist: no partial, a zero column
    npart = np.floor((fs / 2) / np.maximum(f1, f2)).astype(np.int32)
    window = np.abs(rng.standard_normal(wlen)) + 0.1
    outs = [k.wf0_combs(f1, f2, npart, fs, 0.4, Lsig, tb, window, nfft, rows).cpu().numpy()
            for k in (FakeSimmKernels(), ck())]
    assert outs[1].shape == (n, rows)
    assert np.all(outs[1][-1] == 0)
    assert relmax(outs[1], outs[0]) < 1e-10


def test_generators_against_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gold = np.load(GOLD)
    k = ck()
    tr = STFT(linFTLen=256, atomHopFactor=0.25, winFunc=sqrt_blackmanharris, fs=8000, kernels=k)
    t, w, _ = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=2, Ot=0.5, perF0=1,
                                          depthChirpInSemiTone=0.5, kernels=k)
    assert np.array_equal(t, gold["t1"]) and relmax(w, gold["w1"]) < 1e-10
    t, w, _ = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=1, Ot=0.5, perF0=3,
                                          depthChirpInSemiTone=0.5, kernels=k)
    assert np.array_equal(t, gold["t2"]) and relmax(w, gold["w2"]) < 1e-10
    t, w = slf.generate_WF0_chirped(100, 800, 8000, Nfft=256, stepNotes=1, lengthWindow=256,
                                    Ot=0.5, perF0=2, depthChirpInSemiTone=.15,
                                    analysisWindow='sinebell', kernels=k)
    assert np.array_equal(t, gold["t3"]) and relmax(w, gold["w3"]) < 1e-10
    tr2 = STFT(linFTLen=512, atomHopFactor=0.125, winFunc=np.hanning, fs=16000, kernels=k)
    t, w, _ = slf.generate_WF0_TR_chirped(tr2, minF0=60, maxF0=500, stepNotes=1, Ot=0.25, perF0=2,
                                          depthChirpInSemiTone=0.5, kernels=k)
    assert np.array_equal(t, gold["t4"]) and relmax(w, gold["w4"]) < 1e-10
    assert sorted(os.listdir(".")) == sorted(str(n) for n in gold["cache_names"])


def test_separate_lead_process_builds_its_dictionary(tmp_path, monkeypatch):
    """SeparateLeadProcess without a WF0 argument: the reference's default dictionary (39 Hz -
    2 kHz, 16 F0 per semitone = 1092 combs of 1025 bins at 44.1 kHz) is generated on the GPU;
    columns sum to one, every comb peaks at a multiple of its F0, and a sample of columns equals
    the oracle's."""
    import scipy.io.wavfile as wavfile
    from pyfasst_b200.SeparateLeadStereo.SeparateLeadStereoTF import SeparateLeadProcess
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    wav = str(tmp_path / "mix.wav")
    wavfile.write(wav, 44100, np.int16(3000 * rng.standard_normal((8192, 2))))
    t0 = time.perf_counter()
    p = SeparateLeadProcess(wav, verbose=False, outputDirSuffix="out", kernels=ck())
    dt = time.perf_counter() - t0
    WF0, table = p.SIMMParams['WF0'], p.SIMMParams['F0Table']
    assert WF0.shape == (1025, 1092) and table.size == 1092 and p.SIMMParams['NF0'] == 1092
    assert np.allclose(WF0.sum(axis=0), 1.0, atol=1e-12)
    freqs = np.arange(1025) * 44100.0 / 2048
    for i in (300, 700, 1091):  # F0 well above the frequency resolution
        peak = freqs[np.argmax(WF0[:, i])]
        ratio = peak / table[i]
        assert abs(ratio - round(ratio)) < 0.02 + 21.6 / table[i], (i, table[i], peak)
    for i in (0, 511, 1091):
        x = np.real(wo.odgd(table[i], 44100, 4096, 0.5))
        col = np.abs(wo.stft_mid_frame(x, sqrt_blackmanharris(2048), 512, 2048)) ** 2
        assert relmax(WF0[:, i], col / col.sum()) < 1e-9
    print("SeparateLeadProcess construction with the 1025 x 1092 dictionary: %.3f s" % dt)
    # second construction reads the cache the first one wrote
    names = [n for n in os.listdir(".") if n.startswith("wf0_stft_")]
    assert len(names) == 1
    p2 = SeparateLeadProcess(wav, verbose=False, outputDirSuffix="out", kernels=ck())
    assert np.array_equal(p2.SIMMParams['WF0'], WF0)
