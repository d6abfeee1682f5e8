"""The SIMM / Stereo_SIMM oracle against the golden vectors produced by the reference itself
(oracle/make_golden.py: run_simm).  CPU only."""
import os

import numpy as np
from numpy.testing import assert_allclose

from oracle import simm_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    return np.load(os.path.join(GOLDEN, "simm.npz"))


def stereo_beta0(R):
    np.random.seed(5)  # make_golden.py seeds the global RNG before Stereo_SIMM (SIMM.py:581)
    return np.random.rand(R)


def test_mono_simm_oracle_matches_reference():
    g = load()
    SX = 0.5 * (g["SXR"] + g["SXL"])
    res = so.simm(SX, g["WF0"], g["WGAMMA"], g["HGAMMA0"], g["HPHI0"], g["HF00"],
                  g["WM0"][:, :1], g["HM0"][:1], numberOfIterations=4)
    for nm, a in zip(("HGAMMA", "HPHI", "HF0", "HM", "WM"), res):
        assert_allclose(a, g["mono_" + nm], rtol=1e-10, atol=1e-300, err_msg=nm)


def test_stereo_simm_oracle_matches_reference():
    g = load()
    R = g["WM0"].shape[1]
    res = so.stereo_simm(g["SXR"], g["SXL"], g["WF0"], g["WGAMMA"], g["HGAMMA0"], g["HPHI0"],
                         g["HF00"], g["WM0"], g["HM0"], stereo_beta0(R), numberOfIterations=4,
                         computeError=True)
    names = ("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR", "betaL", "HM", "WM",
             "recoError")
    for nm, a in zip(names, res):
        assert_allclose(a, g["st_" + nm], rtol=1e-9, atol=1e-300, err_msg=nm)
