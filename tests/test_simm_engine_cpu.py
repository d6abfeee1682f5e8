"""Host orchestration of SIMM / Stereo_SIMM (pyfasst_b200/simm_engine.py and the drop-in
functions of pyfasst_b200/SeparateLeadStereo/SIMM/SIMM.py) on the NumPy specification of the
kernels, against the reference's golden vectors (tests/golden/simm.npz) and the oracle.  CPU only;
the `-m gpu` twin is tests/test_simm_gpu.py."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import simm_oracle as so
from pyfasst_b200.SeparateLeadStereo.SIMM import SIMM as simm_mod
from tests.fake_simm_kernels import FakeSimmKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# float32 planes / factors against the float64 reference after 4 full iterations
RTOL = 2e-4


def load():
    return np.load(os.path.join(GOLDEN, "simm.npz"))


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_mono_simm_matches_reference():
    g = load()
    SX = 0.5 * (g["SXR"] + g["SXL"])
    res = simm_mod.SIMM(SX, g["WF0"], g["WGAMMA"], numberOfFilters=g["HGAMMA0"].shape[1],
                        numberOfAccompanimentSpectralShapes=1, HGAMMA0=g["HGAMMA0"],
                        HPHI0=g["HPHI0"], HF00=g["HF00"], WM0=g["WM0"][:, :1], HM0=g["HM0"][:1],
                        numberOfIterations=4, verbose=False, kernels=FakeSimmKernels())
    for nm, a in zip(("HGAMMA", "HPHI", "HF0", "HM", "WM"), res):
        assert rel_err(a, g["mono_" + nm]) < RTOL, nm
    assert not res[5].any()  # recoError is never filled by the mono function


def test_stereo_simm_matches_reference():
    g = load()
    R = g["WM0"].shape[1]
    np.random.seed(5)  # the seed make_golden.py used before Stereo_SIMM drew betaR (SIMM.py:581)
    res = simm_mod.Stereo_SIMM(g["SXR"], g["SXL"], g["WF0"], g["WGAMMA"],
                               numberOfFilters=g["HGAMMA0"].shape[1],
                               numberOfAccompanimentSpectralShapes=R, HGAMMA0=g["HGAMMA0"],
                               HPHI0=g["HPHI0"], HF00=g["HF00"], WM0=g["WM0"], HM0=g["HM0"],
                               numberOfIterations=4, verbose=False, computeError=True,
                               kernels=FakeSimmKernels())
    names = ("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR", "betaL", "HM", "WM")
    for nm, a in zip(names, res):
        assert rel_err(np.asarray(a), np.asarray(g["st_" + nm])) < RTOL, nm
    reco, ref = res[9], g["st_recoError"]
    assert reco.shape == ref.shape
    assert_allclose(reco, ref, rtol=1e-4, atol=1e-3)


def test_stereo_simm_options_against_oracle():
    """updateHGAMMA=False and an update exponent != 1, against the oracle."""
    g = load()
    R = g["WM0"].shape[1]
    beta0 = np.linspace(0.2, 0.8, R)
    ref = so.stereo_simm(g["SXR"], g["SXL"], g["WF0"], g["WGAMMA"], g["HGAMMA0"], g["HPHI0"],
                         g["HF00"], g["WM0"], g["HM0"], beta0, numberOfIterations=2,
                         updateRulePower=0.7, updateHGAMMA=False)
    from pyfasst_b200.simm_engine import SimmEngine
    eng = SimmEngine(FakeSimmKernels(), [g["SXR"], g["SXL"]], g["WF0"], g["WGAMMA"], g["HGAMMA0"],
                     g["HPHI0"], g["HF00"], g["WM0"], g["HM0"], betaR=beta0, omega=0.7,
                     update_hgamma=False, n_iter=2)
    eng.iterate()
    eng.iterate()
    r = eng.results()
    got = (r["alphaR"], r["alphaL"], r["HGAMMA"], r["HPHI"], r["HF0"], np.diag(r["betaR"]),
           np.diag(r["betaL"]), r["HM"], r["WM"])
    for nm, a, b in zip("alphaR alphaL HGAMMA HPHI HF0 betaR betaL HM WM".split(), got, ref):
        assert rel_err(np.asarray(a), np.asarray(b)) < RTOL, nm


def test_error_behaviour():
    g = load()
    k = FakeSimmKernels()
    assert simm_mod.SIMM(g["SXR"], g["WF0"][:-1], g["WGAMMA"], kernels=k) is False
    assert simm_mod.Stereo_SIMM(g["SXR"], g["SXL"], g["WF0"][:-1], g["WGAMMA"], kernels=k) is False
    with pytest.raises(ValueError):
        simm_mod.Stereo_SIMM(g["SXR"], g["SXL"][:, :-1], g["WF0"], g["WGAMMA"], kernels=k)
    with pytest.raises(ValueError):  # mono `HM *= sumWM` only broadcasts for R == 1 (or R == N)
        simm_mod.SIMM(g["SXR"], g["WF0"], g["WGAMMA"], numberOfAccompanimentSpectralShapes=3,
                      numberOfIterations=1, verbose=False, kernels=k)


def test_random_initialisation_order():
    """Missing initial matrices are drawn from the global RNG in the reference's order."""
    g = load()
    F, N = g["SXR"].shape
    P, K = g["HGAMMA0"].shape
    NF0 = g["WF0"].shape[1]
    np.random.seed(11)
    init = [np.abs(np.random.randn(*s)) for s in ((P, K), (K, N), (NF0, N), (1, N), (F, 1))]
    ref = so.simm(g["SXR"], g["WF0"], g["WGAMMA"], init[0], init[1], init[2], init[4], init[3],
                  numberOfIterations=1)
    np.random.seed(11)
    res = simm_mod.SIMM(g["SXR"], g["WF0"], g["WGAMMA"], numberOfFilters=K,
                        numberOfAccompanimentSpectralShapes=1, numberOfIterations=1,
                        verbose=False, kernels=FakeSimmKernels())
    for nm, a, b in zip(("HGAMMA", "HPHI", "HF0", "HM", "WM"), res, ref):
        assert rel_err(a, b) < RTOL, nm
