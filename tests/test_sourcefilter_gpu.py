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


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_renormalisation_kernels_with_large_dictionaries(dt):
    """fb_scale_colmax / fw_renorm beyond 64 columns (the 1093-comb dictionary and its square
    weight matrix), on row-strided views like GeneralGemEngine passes them."""
    rng = np.random.default_rng(11)
    F, Kb, Kw = 50, 150, 131
    FBp = torch.tensor(np.abs(rng.standard_normal((F, 152)))).to(dt)
    FWp = torch.tensor(np.abs(rng.standard_normal((152, 132)))).to(dt)
    FWp[7, :] = 0
    FBp[:, 9] = 0  # a zero column: its maximum counts as one
    sums = torch.tensor([3.0, 5.0], dtype=torch.float64)
    counts = torch.tensor([2.0, 4.0], dtype=torch.float64)
    outs = []
    for k, dev in ((FakeKernels(), "cpu"), (ck(), "cuda")):
        FB, FW = FBp.clone().to(dev), FWp.clone().to(dev)
        colmax = torch.zeros(152, dtype=torch.float64, device=dev)
        w = torch.zeros(152, dtype=torch.float64, device=dev)
        w2 = torch.zeros(132, dtype=torch.float64, device=dev)
        k.fb_scale_colmax(FB[:, :Kb], sums.to(dev), counts.to(dev), 1, colmax)
        k.fw_renorm(FW[:Kb, :Kw], colmax, w, w2)
        outs.append([t.cpu().numpy() for t in (FB, FW, colmax, w, w2)])
    tol = 1e-6 if dt == torch.float32 else 1e-13
    for a, b in zip(outs[1], outs[0]):
        assert np.abs(a - b).max() / np.abs(b).max() < tol


def test_separation_and_powers(tmp_path, monkeypatch):
    cpu.check_separation_and_powers(ck(), "float32", 2e-6, 4, tmp_path, monkeypatch)


def test_sparse_model_against_reference(tmp_path, monkeypatch):
    cpu.check_sparse_model(ck(), "float32", 5e-5, 5e-3, tmp_path, monkeypatch)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_sparsity_reweigh_kernel_against_spec(dt):
    rng = np.random.default_rng(5)
    K, N, ld = 41, 333, 352
    TW = np.zeros((44, ld))
    TW[:K, :N] = np.abs(rng.standard_normal((K, N))) ** 3
    TW[:K - 1, 7] = 0.0       # an empty frame: the barycentre falls back on the eps clamp
    outs = []
    for k, dev in ((FakeKernels(), "cpu"), (ck(), "cuda")):
        t = torch.tensor(TW).to(dt).to(dev)
        it = torch.tensor([3], dtype=torch.int32, device=dev)
        work = torch.zeros(2 * N, dtype=torch.float64, device=dev)
        k.sparsity_reweigh(t, K, N, 4, float(np.log(K ** 2)), -0.9, it, work)
        outs.append(t.cpu().numpy())
    tol = 2e-6 if dt == torch.float32 else 1e-12
    assert np.abs(outs[1] - outs[0]).max() / np.abs(outs[0]).max() < tol
    assert (outs[1][K:] == 0).all() and (outs[1][:, N:] == 0).all()
