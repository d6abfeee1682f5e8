"""GPU twin of tests/test_sourcefilter_cpu.py: multiChanSourceF0Filter on the CUDA kernels
(float32 planes, tensor-core GEMM contractions) against the golden vectors made by running the
reference's own class; plus the new elementwise kernels against their NumPy specification."""
import numpy as np
import pytest
import torch

from tests import test_sourcefilter_cpu as cpu
from tests.fake_kernels import FakeKernels

pytestmark = pytest.mark.gpu


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_plane_kernels_against_spec(dt):
    rng = np.random.default_rng(3)
    F, N, ld = 37, 203, 224
    mk = lambda: torch.tensor(np.abs(rng.standard_normal((F, ld))) + 1e-3).to(dt)
    hatW, P, O = mk(), mk(), mk()
    P[3, 5] = 0.0   # clamped at eps
    outs = []
    for k, dev in ((FakeKernels(), "cpu"), (ck(), "cuda")):
        planes = torch.full((F, 2 * ld), 7.0, dtype=dt, device=dev)
        k.gem_ratio_planes(hatW.to(dev), P.to(dev), O.to(dev), planes, N)
        prod = torch.full((F, ld), 7.0, dtype=dt, device=dev)
        k.mul_planes(hatW.to(dev), O.to(dev), prod, N)
        k.mul_planes(P.to(dev), None, prod, N, accumulate=True)
        theta = hatW.to(dev).clone()
        k.mult_update_same(theta, P.to(dev), O.to(dev), F, N, 0.7)
        outs.append([t.cpu().numpy() for t in (planes, prod, theta)])
    tol = 1e-6 if dt == torch.float32 else 1e-13
    for a, b in zip(outs[1], outs[0]):
        assert np.abs(a - b).max() / np.abs(b).max() < tol
    assert (outs[1][0][:, N:ld] == 0).all() and (outs[1][0][:, ld + N:] == 0).all()
    assert (outs[1][1][:, N:] == 0).all()


def test_model_against_reference(tmp_path, monkeypatch):
    # float32 planes: W/H/A within 1e-4 after one iteration (north_star tolerance)
    m = cpu.check_model(ck(), "float32", 5e-6, 1e-4, 2e-5, tmp_path, monkeypatch)
    cpu.check_snapshot(m, cpu.load(), "final", 5e-3)


def test_float64_is_refused():
    with pytest.raises(NotImplementedError):
        cpu.build(ck(), 1, "float64").estim_param_a_post_model()
