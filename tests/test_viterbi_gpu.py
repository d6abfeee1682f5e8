"""The CUDA Viterbi tracker: bit-identical paths to the reference's Cython module (golden vectors,
and the compiled module itself when oracle/_ref travelled to this box) and to the oracle on a
larger problem."""
import numpy as np
import pytest

from oracle import build_ref, viterbi_oracle as vo
from pyfasst_b200.SeparateLeadStereo.tracking import _tracking
from tests import test_viterbi_cpu as cpu

pytestmark = pytest.mark.gpu


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


@pytest.mark.parametrize("tag", cpu.CASES)
def test_function_matches_reference(tag):
    cpu.check_function(ck(), tag)


@pytest.mark.parametrize("S,N,ties", [(481, 700, False), (1093, 300, False), (130, 2000, True),
                                      (2100, 40, False), (3, 9, True)])
def test_larger_problems(S, N, ties):
    rng = np.random.default_rng(S + N)
    dens = rng.standard_normal((S, N)) * 3
    trans = np.log(rng.random((S, S)) + 1e-3)
    prior = np.log(rng.random(S) + 0.1)
    if ties:
        dens, trans = np.round(dens), np.round(trans)
    ref = vo.viterbi_tracking(S, N, dens, prior, trans)
    got = _tracking.viterbiTracking(S, N, dens, prior, trans, kernels=ck())
    np.testing.assert_array_equal(got, ref)
    trk = build_ref.load()
    if trk is not None and S <= 500:
        np.testing.assert_array_equal(got, trk.viterbiTracking(S, N, dens, prior, trans))
