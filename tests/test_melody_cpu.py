"""SeparateLeadProcess melody tracking and chunked second stage (runViterbi,
initiateHF0WithIndexBestPath, checkChunkSize, estimHF0, estimStereoSIMMParamsWriteSeps,
overlapAddChunks, autoMelSepAndWrite) on the NumPy kernel specification.  Golden vectors
(tests/golden/melody.npz): the reference's own runViterbi / initiateHF0WithIndexBestPath source
executed with its compiled Cython tracker (oracle/make_golden.py: run_melody).  CPU only."""
import os
import shutil

import numpy as np
import pytest
import scipy.io.wavfile as wavfile
from numpy.testing import assert_allclose

from pyfasst_b200.SeparateLeadStereo import SeparateLeadStereoTF as sls
from tests.test_lead_sep_cpu import AllFakeKernels, GOLDEN, FS, WLEN, HOP


def load():
    return np.load(os.path.join(GOLDEN, "melody.npz"))


def make_process(tmp_path, kernels, WF0, stepNotes, **kw):
    wav = os.path.join(str(tmp_path), "mix_lead.wav")
    shutil.copy(os.path.join(GOLDEN, "mix_lead.wav"), wav)
    return sls.SeparateLeadProcess(wav, windowSize=WLEN / float(FS), hopsize=HOP, nbIter=2,
                                   numCompAccomp=3, K_numFilters=2, P_numAtomFilters=6,
                                   minF0=100, stepNotes=stepNotes, WF0=WF0, verbose=False,
                                   kernels=kernels, **kw)


def check_tracking(tmp_path, kernels):
    g = load()
    NF0, N = int(g["NF0"]), int(g["N"])
    WF0 = np.abs(np.random.default_rng(0).standard_normal((WLEN // 2 + 1, NF0)))
    for tag, search in (("full", {}), ("band", dict(minF0search=130.0, maxF0search=200.0))):
        proc = make_process(tmp_path, kernels, WF0, int(g["stepNotes"]),
                            maxF0=float(g["F0Table"][-1]) + 1, **search)
        assert_allclose(proc.SIMMParams["F0Table"], g["F0Table"], rtol=1e-14)
        proc.SIMMParams["HF0"] = g["HF0"].copy()
        proc.totFrames = proc.N = N       # the golden HF0 has its own number of frames
        proc.lengthData = 0
        proc.runViterbi()
        np.testing.assert_array_equal(proc.indexBestPath, g[tag + "_path"])
        assert_allclose(proc.freqMelody, g[tag + "_freq"], rtol=1e-14)
        assert_allclose(np.loadtxt(proc.files["pitch_output_file"]), g[tag + "_pitches"], rtol=1e-12)
        proc.initiateHF0WithIndexBestPath()
        np.testing.assert_array_equal(proc.SIMMParams["HF00"], g[tag + "_HF00"])
    proc = make_process(tmp_path, kernels, WF0, 4)
    proc.SIMMParams.pop("HF0", None)
    with pytest.raises(AttributeError):
        proc.runViterbi()


def test_tracking_matches_reference(tmp_path):
    check_tracking(tmp_path, AllFakeKernels())


def test_check_chunk_size(tmp_path):
    WF0 = np.ones((WLEN // 2 + 1, 5))
    proc = make_process(tmp_path, AllFakeKernels(), WF0, 4)
    tot = proc.computeNFrames()
    assert tot == int(np.ceil(3000 / 32.0 + 1) + 1) == proc.N and proc.lengthData == 3000
    assert proc.checkChunkSize(1000) == (tot, 1, 1000)
    assert proc.checkChunkSize(40) == (tot, tot // 40 + 1, 40)
    # a last chunk shorter than a window (8 frames) evens the chunks out (ref: :1888-1897)
    t, n, m = proc.checkChunkSize(tot - 3)
    assert n * m >= t - m and m == int(np.ceil(tot / 2.0)) and n == tot // m


def check_pipeline(tmp_path, kernels):
    """autoMelSepAndWrite end to end; with one chunk the written files equal the one-shot
    separation of the same parameters without the leading half window (offsets['stft'])."""
    rng = np.random.default_rng(5)
    NF0 = 24
    WF0 = np.abs(rng.standard_normal((WLEN // 2 + 1, NF0))) ** 2
    WF0 /= WF0.sum(axis=0)
    # several chunks
    np.random.seed(1)
    proc = make_process(tmp_path, kernels, WF0, 4, freeMemory=False)
    proc.autoMelSepAndWrite(maxFrames=40)
    assert proc.indexBestPath.shape == (proc.totFrames,)
    assert proc.SIMMParams["HF00"].shape == (NF0, proc.totFrames)
    for key in ("voc_output_file", "mus_output_file"):
        fs, x = wavfile.read(proc.files[key])
        assert fs == FS and x.shape == (3000, 2) and x.dtype == np.int16 and np.abs(x).max() > 0
    left = [f for f in os.listdir(proc.files["outputDir"]) if "0000" in f]
    assert not left, "the per-chunk files are removed"
    assert os.path.exists(proc.files["pitch_output_file"])
    # one chunk
    np.random.seed(1)
    one = make_process(tmp_path, kernels, WF0, 4, freeMemory=False)
    one.estimHF0(maxFrames=10 ** 6)
    one.runViterbi()
    one.initiateHF0WithIndexBestPath()
    np.random.seed(2)
    one.estimStereoSIMMParamsWriteSeps(maxFrames=10 ** 6)
    _, voc = wavfile.read(one.files["voc_output_file"])
    one.computeStereoX()
    lead, _ = one.separated_pcm()
    np.testing.assert_array_equal(voc, lead[WLEN // 2:WLEN // 2 + 3000])


def test_pipeline(tmp_path):
    check_pipeline(tmp_path, AllFakeKernels())
