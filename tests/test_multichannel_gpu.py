"""The general-I E-step / Wiener kernels (csrc/estep_multi.cu) against their NumPy specification,
and the engine / public API on 3- and 4-channel mixtures against the generalised oracle (itself
proven equal to the reference-pinned stereo oracle at I = 2 in tests/test_multichannel_cpu.py)."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from tests import test_multichannel_cpu as cpu
from tests.fake_kernels import FakeKernels
from tests.test_kernels_gpu import rel, tol

pytestmark = pytest.mark.gpu

DTYPES = [torch.float64, torch.float32]


@pytest.fixture(scope="module")
def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


def problem(rng, dt, I, F, N, J, rank):
    """E-step inputs that follow the model: x = sum_r a_r s_r + noise, s_r ~ CN(0, v)."""
    ld = (N + 31) // 32 * 32
    R = J * rank
    src = [j for j in range(J) for _ in range(rank)]
    Vn = np.abs(rng.standard_normal((J, F, N))) + 0.05
    A = rng.standard_normal((R, I, F)) + 1j * rng.standard_normal((R, I, F))
    noise = np.abs(rng.standard_normal(F)) * 0.01 + 2e-2
    cn = lambda shape: (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)
    Xc = np.sqrt(noise)[None, :, None] * cn((I, F, N))
    for r in range(R):
        Xc = Xc + A[r][:, :, None] * (np.sqrt(Vn[src[r]]) * cn((F, N)))[None]
    Xn = np.zeros((2 * I, F, ld))
    Xn[0::2, :, :N], Xn[1::2, :, :N] = Xc.real, Xc.imag
    Vp = np.zeros((J, F, ld))
    Vp[:, :, :N] = Vn
    return (ld, R, src, torch.tensor(Xn).to(dt), torch.tensor(Vp).to(dt),
            torch.tensor(np.ascontiguousarray(A)), torch.tensor(noise))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("I,F,N,J,rank", [(4, 17, 1000, 4, 4), (4, 9, 333, 4, 1), (3, 12, 700, 3, 2),
                                          (2, 21, 2600, 4, 2), (4, 5, 40, 1, 2), (3, 3, 4, 6, 1)])
def test_estep_multi(ck, dt, I, F, N, J, rank):
    rng = np.random.default_rng(1000 * I + N)
    ld, R, src, X, V, A, noise = problem(rng, dt, I, F, N, J, rank)
    outs = []
    for k, dev in ((FakeKernels(), "cpu"), (ck, "cuda")):
        hatW = torch.zeros((J, F, ld), dtype=dt, device=dev)
        Rss = torch.zeros((F, R, R), dtype=torch.complex128, device=dev)
        Rxs = torch.zeros((F, I, R), dtype=torch.complex128, device=dev)
        ll = torch.zeros(F, dtype=torch.float64, device=dev)
        ws = torch.zeros((k.estep_multi_workspace_bytes(I, J, F, N) + 7) // 8, dtype=torch.float64,
                         device=dev)
        k.estep_multi(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), N, hatW, Rss, Rxs, ll, ws)
        outs.append([t.cpu().numpy() for t in (hatW, Rss, Rxs, ll)])
    (hw0, rss0, rxs0, ll0), (hw1, rss1, rxs1, ll1) = outs
    assert np.isfinite(hw1).all() and np.isfinite(rss1).all()
    assert rel(hw1[:, :, :N], hw0[:, :, :N]) < tol(dt, f32=1e-6)
    assert (hw1[:, :, N:] == 0).all(), "padding frames must stay zero"
    t = tol(dt, f64=1e-9, f32=1e-9)  # the moments are float64 whatever the plane type
    assert rel(rss1, rss0) < t
    assert rel(rxs1, rxs0) < t
    # float32 planes take the float logarithm (as the stereo kernel): ~1e-6 absolute per bin
    assert_allclose(ll1, ll0, rtol=tol(dt, f64=1e-11, f32=1e-6),
                    atol=0.0 if dt == torch.float64 else 1e-6 * N)
    assert_allclose(rss1, np.conj(np.transpose(rss1, (0, 2, 1))), atol=1e-14 * np.abs(rss1).max())


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("I,J,rank", [(4, 4, 2), (3, 2, 3), (2, 3, 2)])
def test_estep_multi_determinant_clamp(ck, dt, I, J, rank):
    """Quiet rows: det Sigma < 1e-10 activates the determinant clamp (the generic form of
    signalTools.py:183-188), where the scaled Sigma^-1 is not the inverse of Sigma and the kernel's
    identity for x y^H needs its correction term; rows 0-2 are clamped in every bin, rows 3-5 in
    some, the rest in none."""
    rng = np.random.default_rng(177 + I)
    F, N = 9, 1300
    ld, R, src, X, V, A, noise = problem(rng, dt, I, F, N, J, rank)
    scale = np.ones(F)
    scale[:3], scale[3:6] = 10.0 ** (-14.0 / I), 10.0 ** (-9.0 / I)
    sc = torch.tensor(scale)
    X = (X.to(torch.float64) * torch.sqrt(sc)[None, :, None]).to(dt)
    V = (V.to(torch.float64) * sc[None, :, None]).to(dt)
    V[:, 3:6, ::3] *= 1e-3
    X[:, 3:6, ::3] *= 1e-2
    noise = noise * sc
    outs = []
    for k, dev in ((FakeKernels(), "cpu"), (ck, "cuda")):
        hatW = torch.zeros((J, F, ld), dtype=dt, device=dev)
        Rss = torch.zeros((F, R, R), dtype=torch.complex128, device=dev)
        Rxs = torch.zeros((F, I, R), dtype=torch.complex128, device=dev)
        ll = torch.zeros(F, dtype=torch.float64, device=dev)
        ws = torch.zeros((k.estep_multi_workspace_bytes(I, J, F, N) + 7) // 8, dtype=torch.float64,
                         device=dev)
        k.estep_multi(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), N, hatW, Rss, Rxs, ll, ws)
        outs.append([t.cpu().numpy() for t in (hatW, Rss, Rxs, ll)])
    (hw0, rss0, rxs0, ll0), (hw1, rss1, rxs1, ll1) = outs
    assert np.isfinite(hw1).all() and np.isfinite(rss1).all() and np.isfinite(rxs1).all()
    for f in range(F):  # per row: the rows differ by many orders of magnitude
        assert rel(hw1[:, f, :N], hw0[:, f, :N]) < tol(dt, f64=1e-7, f32=1e-5), f
        assert rel(rss1[f], rss0[f]) < 1e-7, f
        assert rel(rxs1[f], rxs0[f]) < 1e-7, f
    assert_allclose(ll1, ll0, rtol=tol(dt, f64=1e-10, f32=1e-6),
                    atol=0.0 if dt == torch.float64 else 1e-6 * N)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("I", [2, 3, 4])
def test_wiener_multi(ck, dt, I):
    rng = np.random.default_rng(5 + I)
    F, N, J, rank = 19, 1030, 3, 2
    ld, R, src, X, V, A, noise = problem(rng, dt, I, F, N, J, rank)
    groups = [1, -1, 0]
    outs = []
    for k, dev in ((FakeKernels(), "cpu"), (ck, "cuda")):
        Y = torch.zeros((2 * 2 * I, F, ld), dtype=dt, device=dev)
        ws = torch.zeros(8192, dtype=torch.float64, device=dev)
        k.wiener_multi(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), groups, 2, N, Y, ws)
        outs.append(Y.cpu().numpy())
    assert rel(outs[1][:, :, :N], outs[0][:, :, :N]) < tol(dt, f32=1e-6)
    assert (outs[1][:, :, N:] == 0).all()


@pytest.mark.parametrize("nch,conv,rank", [(4, False, 2), (4, True, 2), (3, True, 1)])
def test_engine_multichannel_float64(ck, nch, conv, rank):
    ref, model = cpu.run_engine(ck, nch, conv, rank, 3, "float64")
    cpu.compare(ref, model, 1e-8)


@pytest.mark.parametrize("nch,conv,rank", [(4, True, 2), (4, False, 2)])
def test_engine_multichannel_float32(ck, nch, conv, rank):
    """float32 planes: the north_star tolerances (parameters <= 1e-4 after one iteration,
    log-likelihoods <= 1e-5) on a 4-channel mixture with at least as many sub-sources as
    channels (with fewer, the model is rank deficient, the noise floor alone explains part of x
    and the mixing update is ill conditioned in ANY precision)."""
    ref, model = cpu.run_engine(ck, nch, conv, rank, 1, "float32")
    ll_ref, ll = ref.estim_param_a_post_model(), model.estim_param_a_post_model()
    assert_allclose(ll, ll_ref, rtol=1e-5)
    for j in ref.spat_comps:
        pa, pb = np.asarray(model.spat_comps[j]["params"]), np.asarray(ref.spat_comps[j]["params"])
        assert np.abs(pa - pb).max() <= 1e-4 * np.abs(pb).max(), ("A", j)
        for key in ("FB", "TW"):
            xa, xb = model.spec_comps[j]["factor"][0][key], ref.spec_comps[j]["factor"][0][key]
            assert np.abs(xa - xb).max() <= 1e-4 * np.abs(xb).max(), (key, j)


def test_stereo_through_the_general_kernels(ck, monkeypatch):
    monkeypatch.setenv("PYFASST_FORCE_MULTI", "1")
    ref, model = cpu.run_engine(ck, 2, True, 2, 3, "float64")
    cpu.compare(ref, model, 1e-8)
