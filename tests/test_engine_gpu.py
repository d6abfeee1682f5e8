"""The GEM engine on the CUDA kernels (through the C ABI) against the golden vectors
produced by the reference itself and against the oracle.  Needs a B200."""
import copy
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from tests.test_engine_cpu import CASES, GOLDEN, check_params, engine_for, oracle_model, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_gem_f64_matches_reference(ck, name, wav, conv, rank, nbcomps):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = oracle_model(wav, conv, rank, nbcomps)
    m.noise["PSD"] = m.noise["ann_PSD_lim"][0]
    eng = engine_for(m, "float64", ck)
    hRxs, hRss, hW, ll = eng.suff_stat()
    assert_allclose(hRxs, g["e0_hat_Rxs"], rtol=1e-8, atol=1e-13)
    assert_allclose(hRss, g["e0_hat_Rss"], rtol=1e-8, atol=1e-13)
    assert_allclose(ll, g["e0_loglik"], rtol=1e-11)
    ref_hW = np.array([g["e0_hat_Ws"][eng.ranks[j]].mean(0) for j in range(eng.J)])
    assert_allclose(hW, ref_hW, rtol=1e-7, atol=1e-300)
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng = engine_for(m, "float64", ck)
    ll1 = eng.run(1)
    eng.read_model(spat, spec)
    assert_allclose(ll1, g["ll_it1"], rtol=1e-11)
    check_params(spat, spec, g, "it1", 1e-9)
    eng = engine_for(m, "float64", ck)
    lls = eng.run(6)
    eng.read_model(spat, spec)
    assert_allclose(lls, g["logliks"], rtol=1e-9)
    check_params(spat, spec, g, "final", 1e-7)
    assert_allclose(eng.noise_psd(), g["noise_PSD_final"], rtol=1e-12)


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_gem_f32_within_north_star_tolerances(ck, name, wav, conv, rank, nbcomps):
    """float32 planes: W/H/A <= 1e-4 relative after 1 iteration, LL trajectory <= 1e-5."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = oracle_model(wav, conv, rank, nbcomps)
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng = engine_for(m, "float32", ck)
    ll1 = eng.run(1)
    eng.read_model(spat, spec)
    assert_allclose(ll1, g["ll_it1"], rtol=1e-5)
    check_params(spat, spec, g, "it1", 1e-4)
    eng = engine_for(m, "float32", ck)
    lls = eng.run(6)
    assert_allclose(lls, g["logliks"], rtol=1e-5)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_cuda_graph_replay_equals_eager(ck, dtype):
    m = oracle_model("mix_conv.wav", True, 2, 2)
    a = engine_for(m, dtype, ck).run(6, use_graph=False)
    b = engine_for(m, dtype, ck).run(6, use_graph=True)
    assert_allclose(b, a, rtol=1e-12 if dtype == "float64" else 1e-6)


def test_gem_50_iterations_loglik_trajectory(ck):
    """north_star: log-likelihood trajectory within 1e-5 relative over 50 iterations
    (float32 device path vs the float64 oracle, same initialisation)."""
    m = oracle_model("mix_inst.wav", False, 2, 3, iters=50)
    eng = engine_for(m, "float32", ck)
    lls = eng.run(50)
    ref = m.estim_param_a_post_model()
    assert_allclose(lls, ref, rtol=1e-5)
    eng64 = engine_for(oracle_model("mix_inst.wav", False, 2, 3, iters=50), "float64", ck)
    assert_allclose(eng64.run(50), ref, rtol=1e-8)


def test_wiener_matches_oracle(ck):
    m = oracle_model("mix_inst.wav", False, 2, 3)
    m.estim_param_a_post_model()
    for dtype, tol in (("float64", 1e-9), ("float32", 2e-5)):
        eng = engine_for(m, dtype, ck)
        eng.noise[:] = eng._f64(m.noise["PSD"])
        Y = eng.wiener(list(range(eng.J)), eng.J).cpu().numpy()
        WG = m.separation_gains()
        for n in range(eng.J):
            for c in range(2):
                ref = WG[n, c, 0] * m.X[0] + WG[n, c, 1] * m.X[1]
                got = Y[4 * n + 2 * c, :, :eng.N] + 1j * Y[4 * n + 2 * c + 1, :, :eng.N]
                assert rel_err(got, ref) < tol
