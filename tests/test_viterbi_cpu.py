"""Viterbi melody tracking: the oracle and the NumPy kernel specification against golden paths
decoded by the reference's own Cython module (compiled by oracle/build_ref.py), and against that
module directly when it is present (it is built from /root/reference, which exists only in the
authoring container, and travels to the GPU box as oracle/_ref/*.so).  CPU only."""
import os

import numpy as np
import pytest

from oracle import build_ref, viterbi_oracle as vo
from pyfasst_b200.SeparateLeadStereo.tracking import _tracking
from tests.fake_simm_kernels import FakeSimmKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("a", "b", "c", "d", "m")


def load():
    return np.load(os.path.join(GOLDEN, "viterbi.npz"))


def sizes(g, tag):
    dens = g[tag + "_dens"]
    return dens.shape[0] - 1, dens.shape[1]  # the arrays hold one more state than is decoded


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference(tag):
    g = load()
    S, N = sizes(g, tag)
    path = vo.viterbi_tracking(S, N, g[tag + "_dens"], g[tag + "_prior"], g[tag + "_trans"])
    np.testing.assert_array_equal(path, g[tag + "_path"])


def check_function(kernels, tag):
    g = load()
    S, N = sizes(g, tag)
    path = _tracking.viterbiTracking(S, N, g[tag + "_dens"], g[tag + "_prior"], g[tag + "_trans"],
                                     kernels=kernels)
    assert path.dtype == np.int64 and path.shape == (N,)
    np.testing.assert_array_equal(path, g[tag + "_path"])


@pytest.mark.parametrize("tag", CASES)
def test_function_on_kernel_spec(tag):
    check_function(FakeSimmKernels(), tag)


def test_oracle_against_compiled_reference():
    trk = build_ref.load()
    if trk is None:
        pytest.skip("the reference's Cython tracker is not available (no /root/reference and no "
                    "oracle/_ref build)")
    rng = np.random.default_rng(1)
    for S, N in ((5, 50), (40, 120), (97, 61)):
        dens = np.round(rng.standard_normal((S, N)) * 3) / 2   # many exact ties
        prior = np.log(rng.random(S))
        trans = np.round(np.log(rng.random((S, S))) * 2) / 2
        np.testing.assert_array_equal(vo.viterbi_tracking(S, N, dens, prior, trans),
                                      trk.viterbiTracking(S, N, dens, prior, trans))
