"""Host side of the drop-in API (pyfasst_b200.audioModel etc.) on NumPy stand-in kernels,
against the golden vectors produced by the reference; the reference's own known-answer
tests for the helpers; and the C-ABI library's exported symbols.  CPU only."""
import ctypes
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_almost_equal

import pyfasst_b200.audioModel as am
from pyfasst_b200.tools import utils
from tests.fake_kernels import FakeKernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ---- reference KATs: pyfasst_tests/pyfasst/tools/test_utils.py:11-63 -----------------
def test_utils_kats():
    assert utils.db(10.) == 10.
    assert utils.ident(3.5) == 3.5
    assert utils.nextpow2(2) == 2 and utils.nextpow2(2 ** 10 + 1) == 2 ** 11
    assert_array_almost_equal(utils.sinebell(5),
                              [0., 0.58778525, 0.95105652, 0.95105652, 0.58778525])
    assert_array_almost_equal(utils.hann(11), np.hanning(11))
    assert_array_almost_equal(
        utils.sqrt_blackmanharris(10),
        np.array([0.00774597, 0.12276471, 0.38345737, 0.72150884, 0.96522498,
                  0.96522498, 0.72150884, 0.38345737, 0.12276471, 0.00774597]))
    assert_array_almost_equal(
        utils.sqrt_blackmanharris(22),
        np.array([0.00774597, 0.04002509, 0.09757199, 0.18273402, 0.29549176, 0.43029881,
                  0.57691724, 0.72150884, 0.84850893, 0.94306403, 0.99353553, 0.99353553,
                  0.94306403, 0.84850893, 0.72150884, 0.57691724, 0.43029881, 0.29549176,
                  0.18273402, 0.09757199, 0.04002509, 0.00774597]))
    assert utils.db(1) == 0 and utils.nextpow2(2 ** 20 + 1) == 2 ** 21
    assert_array_almost_equal(
        utils.sinebell(10),
        np.array([0., 0.30901699, 0.58778525, 0.80901699, 0.95105652, 1., 0.95105652,
                  0.80901699, 0.58778525, 0.30901699]))
    assert_array_almost_equal(utils.hann(22), np.hanning(22))


# ---- the C ABI: library loads, every declared symbol is exported and bound ---------------
def test_cabi_exports_every_declared_symbol():
    from pyfasst_b200 import _lib
    from pyfasst_b200.build import build
    path = build()
    header = open(os.path.join(ROOT, "include", "pyfasst_b200.h")).read()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib.pf_abi_version.restype = ctypes.c_int
    assert lib.pf_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("J,F,N,dtype", [(4, 1025, 51682, 1), (3, 1025, 1122, 1), (4, 129, 51682, 1),
                                         (4, 1025, 51682, 0), (1, 1, 1, 1), (6, 513, 2586, 1)])
def test_estep_plan_covers_every_frame(J, F, N, dtype, monkeypatch):
    """pf_estep_plan is host-only: the splits of a frequency row cover all its frames, the
    workspace holds the partial moments of every (frequency, split) plus the per-frequency
    coefficients, and enough CTAs are planned to fill a 148-SM part when the problem allows."""
    from pyfasst_b200.build import build
    lib = ctypes.CDLL(build())
    i64, i32 = ctypes.c_int64, ctypes.c_int

    def plan():
        chunk, ns, nbytes = i64(), i32(), i64()
        rc = lib.pf_estep_plan(i32(J), i64(N), i32(dtype), ctypes.byref(chunk), ctypes.byref(ns),
                               ctypes.byref(nbytes), i32(F))
        assert rc == 0
        return chunk.value, ns.value, nbytes.value

    monkeypatch.delenv("PYFASST_ESTEP_PASSES", raising=False)
    chunk, ns, nbytes = plan()
    nacc = 4 * (J * (J + 1) // 2) + 13 * J + 1  # S, Z, sv, clamp corrections, ll
    ncoef = 4 * J + J * (J + 1) // 2
    assert ns >= 1 and chunk * ns >= N and chunk * (ns - 1) < N
    assert nbytes == F * ns * nacc * 8 + F * ncoef * 8
    pass_bins = 128 * (2 if dtype == 1 else 4)  # PF_F64 = 1: double2 accesses
    assert chunk % pass_bins == 0 and chunk // pass_bins <= 64
    passes = -(-N // pass_bins)
    assert F * ns >= min(148 * 8, F * passes)
    # tuning knob: passes per CTA
    monkeypatch.setenv("PYFASST_ESTEP_PASSES", "2")
    chunk2, ns2, _ = plan()
    assert chunk2 == min(chunk, 2 * pass_bins) and chunk2 * ns2 >= N


def test_cuda_kernels_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyfasst_b200._lib import CudaKernels
    with pytest.raises(RuntimeError):
        CudaKernels()
    with pytest.raises(RuntimeError):
        am.MultiChanNMFInst_FASST(os.path.join(GOLDEN, "mix_inst.wav"), wlen=256, hopsize=64)


# ---- the model classes -----------------------------------------------------------------------
CASES = [("fasst_inst_r1", "mix_inst.wav", False, 1, 3),
         ("fasst_conv_r2", "mix_conv.wav", True, 2, 2)]


def build_model(wav, conv, rank, nbcomps, dtype="float64", iters=6):
    np.random.seed(0)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    model = cls(audio=os.path.join(GOLDEN, wav), nbComps=nbcomps, nbNMFComps=4,
                spatial_rank=rank, wlen=256, hopsize=64, iter_num=iters, verbose=0,
                ann_PSD_lim=[None, None], compute_dtype=dtype, kernels=FakeKernels())
    if conv:
        model.makeItConvolutive()
    return model


@pytest.mark.parametrize("nbcomps,rank,conv", [(4, 2, False), (3, 3, True), (5, 2, False)])
def test_compute_suff_stat_more_than_six_subsources(nbcomps, rank, conv):
    """8 / 9 / 10 sub-sources (e.g. the headline model: 4 sources at rank 2): the E-step kernel takes
    6 spatial components, compute_suff_stat covers the sub-sources in several passes (sub-sources
    of equal power merged); checked against the oracle (reference algorithm, any R)."""
    from oracle import fasst_oracle as fo
    wav = "mix_conv.wav" if conv else "mix_inst.wav"
    model = build_model(wav, conv, rank, nbcomps)
    np.random.seed(0)
    ref = fo.OracleFASST(os.path.join(GOLDEN, wav), nbComps=nbcomps, nbNMFComps=4,
                         spatial_rank=rank, wlen=256, hopsize=64, iter_num=1)
    if conv:
        ref.makeItConvolutive()
    # the same parameters on both sides
    ref.spat_comps, ref.spec_comps = model.spat_comps, model.spec_comps
    ref.noise["PSD"] = model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    powers, mix, ranks = model.retrieve_subsrc_params()
    assert powers.shape[0] == nbcomps * rank > 6
    _, hRxs, hRss, hWs, ll = model.compute_suff_stat(powers, mix)
    _, rRxs, rRss, rWs, rll = ref.compute_suff_stat(powers, mix)
    assert_allclose(hRxs, rRxs, rtol=1e-8, atol=1e-13)
    assert_allclose(hRss, rRss, rtol=1e-8, atol=1e-13)
    assert_allclose(hWs, rWs, rtol=1e-7, atol=1e-300)
    assert_allclose(ll, np.real(rll), rtol=1e-11)


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_model_api_matches_reference(name, wav, conv, rank, nbcomps, tmp_path):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = build_model(wav, conv, rank, nbcomps)
    if "Cx" in g:
        assert_allclose(model.Cx, g["Cx"], atol=1e-12)
    assert_allclose(model.noise["ann_PSD_lim"][0], g["ann0"], rtol=1e-12)
    assert_allclose(model.noise["ann_PSD_lim"][1], g["ann1"], rtol=1e-12)
    for j in range(nbcomps):
        assert rel_err(model.spat_comps[j]["params"], g["init_A%d" % j]) < 1e-12
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["init_%s%d" % (nm, j)]) < 1e-12
    # compute_suff_stat on the initial parameters (per sub-source hat_Ws)
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    powers, mix, ranks = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(powers, mix)
    assert_allclose(hRxs, g["e0_hat_Rxs"], rtol=1e-8, atol=1e-13)
    assert_allclose(hRss, g["e0_hat_Rss"], rtol=1e-8, atol=1e-13)
    assert_allclose(hWs, g["e0_hat_Ws"], rtol=1e-7, atol=1e-300)
    assert_allclose(ll, g["e0_loglik"], rtol=1e-11)
    lls = model.estim_param_a_post_model()
    assert_allclose(lls, g["logliks"], rtol=1e-9)
    for j in range(nbcomps):
        assert rel_err(model.spat_comps[j]["params"], g["final_A%d" % j]) < 1e-7
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["final_%s%d" % (nm, j)]) < 1e-7
    assert_allclose(model.noise["PSD"], g["noise_PSD_final"], rtol=1e-12)
    model.separate_spat_comps(dir_results=str(tmp_path))
    import scipy.io.wavfile as wavfile
    assert len(model.files["spat_comp"]) == nbcomps
    for n, f in enumerate(model.files["spat_comp"]):
        assert os.path.basename(f) == "%s_%d-%d.wav" % (wav[:-4], n, nbcomps)
        fs, y = wavfile.read(f)
        ref = g["sep%d" % n]
        assert y.shape == ref.shape and y.dtype == ref.dtype
        diff = np.abs(y.astype(int) - ref.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


def test_gem_iteration_and_errors():
    model = build_model("mix_inst.wav", False, 1, 3)
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    g = np.load(os.path.join(GOLDEN, "fasst_inst_r1.npz"))
    ll = model.GEM_iteration()
    assert_allclose(ll, g["ll_it1"][0], rtol=1e-11)
    with pytest.raises(AttributeError):
        am.FASST(audio=3)
    with pytest.raises(NotImplementedError):
        am.FASST(audio=os.path.join(GOLDEN, "mix_inst.wav"), transf="mqt")
    with pytest.raises(AttributeError):
        model.setSpecCompFB(0, np.ones([3, 2]))
