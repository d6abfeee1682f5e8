"""Host logic of the GEM engine (pyfasst_b200/engine.py) on NumPy stand-in kernels,
against the oracle and the golden vectors produced by the reference itself.  CPU only.

This pins the orchestration (order of updates, quirks Q1-Q8, renormalisation, the
source-pair-moment form of the E-step) independently of the CUDA implementation; the
`-m gpu` tests then compare each CUDA kernel with its stand-in and run the same
end-to-end comparisons on the device.
"""
import copy
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import fasst_oracle as fo
from pyfasst_b200.engine import GemEngine, shard_bounds
from tests.fake_kernels import FakeKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = [("fasst_inst_r1", "mix_inst.wav", False, 1, 3),
         ("fasst_inst_r2", "mix_inst.wav", False, 2, 3),
         ("fasst_conv_r1", "mix_conv.wav", True, 1, 3),
         ("fasst_conv_r2", "mix_conv.wav", True, 2, 2)]


def oracle_model(wav, conv, rank, nbcomps, iters=6, K=4):
    np.random.seed(0)
    m = fo.OracleFASST(os.path.join(GOLDEN, wav), nbComps=nbcomps, nbNMFComps=K,
                       spatial_rank=rank, wlen=256, hopsize=64, iter_num=iters)
    if conv:
        m.makeItConvolutive()
    return m


def engine_for(m, dtype, kernels=None, comm=None):
    eng = GemEngine(kernels or FakeKernels(), m.nbFreqsSigRepr, m.nbFramesSigRepr, dtype=dtype,
                    comm=comm)
    eng.set_X_host(m.X)
    lim = m.noise["ann_PSD_lim"]
    eng.set_noise(m.noise["sim_ann_opt"], lim[0], lim[1], m.noise["PSD"])
    eng.set_model(m.spat_comps, m.spec_comps, m.nmfUpdateCoeff)
    return eng


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def check_params(spat, spec, g, prefix, tol):
    for j, sc in spat.items():
        assert rel_err(sc["params"], g["%s_A%d" % (prefix, j)]) < tol
    for k, sp in spec.items():
        fac = sp["factor"][0]
        for nm in ("FB", "FW", "TW"):
            assert rel_err(fac[nm], g["%s_%s%d" % (prefix, nm, k)]) < tol, (nm, k)


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_engine_f64_matches_reference(name, wav, conv, rank, nbcomps):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = oracle_model(wav, conv, rank, nbcomps)
    # E-step statistics on the initial parameters
    m.noise["PSD"] = m.noise["ann_PSD_lim"][0]
    eng = engine_for(m, "float64")
    hRxs, hRss, hW, ll = eng.suff_stat()
    assert_allclose(hRxs, g["e0_hat_Rxs"], rtol=1e-8, atol=1e-13)
    assert_allclose(hRss, g["e0_hat_Rss"], rtol=1e-8, atol=1e-13)
    assert_allclose(ll, g["e0_loglik"], rtol=1e-11)
    # hat_W = rank-mean of the reference's per-sub-source hat_Ws (audioModel.py:408-414)
    ref_hW = np.array([g["e0_hat_Ws"][eng.ranks[j]].mean(0) for j in range(eng.J)])
    assert_allclose(hW, ref_hW, rtol=1e-7, atol=1e-300)
    # one iteration, then the whole trajectory
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng = engine_for(m, "float64")
    ll1 = eng.run(1)
    eng.read_model(spat, spec)
    assert_allclose(ll1, g["ll_it1"], rtol=1e-11)
    check_params(spat, spec, g, "it1", 1e-9)
    eng = engine_for(m, "float64")
    lls = eng.run(6)
    eng.read_model(spat, spec)
    assert_allclose(lls, g["logliks"], rtol=1e-9)
    check_params(spat, spec, g, "final", 1e-7)
    assert_allclose(eng.noise_psd(), g["noise_PSD_final"], rtol=1e-12)


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_engine_f32_within_north_star_tolerances(name, wav, conv, rank, nbcomps):
    """float32 planes/factors: W/H/A <= 1e-4 relative after one iteration, log-likelihood
    trajectory <= 1e-5 relative (BASELINE.json north_star)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = oracle_model(wav, conv, rank, nbcomps)
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng = engine_for(m, "float32")
    ll1 = eng.run(1)
    eng.read_model(spat, spec)
    assert_allclose(ll1, g["ll_it1"], rtol=1e-5)
    check_params(spat, spec, g, "it1", 1e-4)
    eng = engine_for(m, "float32")
    lls = eng.run(6)
    assert_allclose(lls, g["logliks"], rtol=1e-5)


def test_engine_wiener_matches_oracle():
    m = oracle_model("mix_inst.wav", False, 2, 3)
    m.estim_param_a_post_model()
    eng = engine_for(m, "float64")
    # the oracle's noise PSD after the run is the last iteration's (Q8)
    eng.noise[:] = eng._f64(m.noise["PSD"])
    Y = eng.wiener(list(range(eng.J)), eng.J).numpy()
    WG = m.separation_gains()
    for n in range(eng.J):
        for c in range(2):
            ref = WG[n, c, 0] * m.X[0] + WG[n, c, 1] * m.X[1]
            got = Y[4 * n + 2 * c, :, :eng.N] + 1j * Y[4 * n + 2 * c + 1, :, :eng.N]
            assert_allclose(got, ref, rtol=1e-8, atol=1e-12)


def test_engine_rejects_unsupported_structures():
    m = oracle_model("mix_inst.wav", False, 1, 3)
    eng = GemEngine(FakeKernels(), m.nbFreqsSigRepr, m.nbFramesSigRepr, dtype="float64")
    spec = copy.deepcopy(m.spec_comps)
    spec[0]["factor"][0]["TB"] = np.ones([m.nbFramesSigRepr, m.nbFramesSigRepr])
    with pytest.raises(NotImplementedError):
        eng.set_model(m.spat_comps, spec)
    spec = copy.deepcopy(m.spec_comps)
    spec[0]["factor"][0]["FW_frdm_prior"] = "free"
    with pytest.raises(NotImplementedError):
        eng.set_model(m.spat_comps, spec)
    spat = copy.deepcopy(m.spat_comps)
    spat[0]["params"] = np.ones([3, 1])
    with pytest.raises(AttributeError):
        eng.set_model(spat, m.spec_comps)


def test_fixed_priors_are_respected():
    m = oracle_model("mix_inst.wav", False, 1, 3, iters=2)
    m.spec_comps[1]["factor"][0]["FB_frdm_prior"] = "fixed"
    m.spec_comps[2]["factor"][0]["TW_frdm_prior"] = "fixed"
    m.spat_comps[0]["frdm_prior"] = "fixed"
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng = engine_for(m, "float64")
    lls = eng.run(2)
    eng.read_model(spat, spec)
    ref = m.estim_param_a_post_model()
    assert_allclose(lls, ref, rtol=1e-10)
    for j in range(3):
        assert rel_err(spat[j]["params"], m.spat_comps[j]["params"]) < 1e-9
        for nm in ("FB", "FW", "TW"):
            assert rel_err(spec[j]["factor"][0][nm], m.spec_comps[j]["factor"][0][nm]) < 1e-9


def test_shard_bounds():
    b = shard_bounds(1025, 8)
    assert b[0] == (0, 129) and b[-1] == (897, 1025)
    assert all(b[i][1] == b[i + 1][0] for i in range(7))
    assert shard_bounds(5, 1) == [(0, 5)]
