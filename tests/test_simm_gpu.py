"""SIMM / Stereo_SIMM on the GPU: every kernel of csrc/simm.cu against its NumPy specification
(tests/fake_simm_kernels.py), the split-K tensor-core GEMM against float64 NumPy, and the public
functions against the reference's golden vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import simm_oracle as so
from tests.fake_simm_kernels import FakeSimmKernels

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 2e-4  # float32 planes / 3xTF32 products against the float64 reference, 4 iterations


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


def rel_err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def ru4(n):
    return (n + 3) // 4 * 4


class Pair(object):
    """The same float32 / float64 buffers on the CPU (for the specification) and on the GPU."""

    def __init__(self, **arrays):
        self.cpu = {k: torch.tensor(np.ascontiguousarray(v)) for k, v in arrays.items()}
        self.gpu = {k: v.clone().cuda() for k, v in self.cpu.items()}

    def run(self, name, build_args):
        getattr(FakeSimmKernels(), name)(*build_args(self.cpu))
        getattr(ck(), name)(*build_args(self.gpu))
        torch.cuda.synchronize()

    def check(self, key, rtol=2e-6, rows=None):
        a, b = self.gpu[key].cpu().numpy(), self.cpu[key].numpy()
        if rows is not None:
            a, b = a[:rows], b[:rows]
        assert np.isfinite(a).all(), key
        err = rel_err(a.astype(np.float64), b.astype(np.float64))
        assert err < rtol, (key, err)


def planes(rng, F, N, nch):
    ldn = ru4(N)

    def pl(width):
        a = np.zeros((F, width * ldn), np.float32)
        for c in range(width):
            a[:, c * ldn:c * ldn + N] = rng.random((F, N)) ** 3 + 1e-3
        return a
    return ldn, pl


@pytest.mark.parametrize("nch", [1, 2])
@pytest.mark.parametrize("F,N", [(37, 101), (129, 1030)])
def test_plane_kernels(nch, F, N):
    rng = np.random.default_rng(F + N + nch)
    ldn, pl = planes(rng, F, N, nch)
    a2 = np.array([0.3, 0.6], np.float32) if nch == 2 else np.ones(2, np.float32)
    p = Pair(SX=pl(nch), SM=pl(nch), SF0=pl(1), SPHI=pl(1), a2=a2,
             lead=np.full((F, 2 * ldn), 7, np.float32), acc=np.full((F, 2 * nch * ldn), 7, np.float32),
             newhat=np.zeros((F, nch * ldn), np.float32), isd=np.zeros(1), ws=np.zeros(148 * 4 * 4))
    for which in (0, 1):
        p.run("simm_lead_terms", lambda d: (d["SM"], d["SF0"], d["SPHI"], d["SX"], d["a2"], which,
                                            d["lead"], nch, F, N, ldn))
        p.check("lead", rtol=5e-6)
    for sq in (0, 1):
        p.run("simm_acc_terms", lambda d: (d["SM"], d["SF0"], d["SPHI"], d["SX"], d["a2"], d["acc"],
                                           nch, sq, F, N, ldn))
        p.check("acc", rtol=5e-6)
    p.run("simm_hat", lambda d: (d["SM"], d["SF0"], d["SPHI"], d["a2"], d["newhat"], nch, F, N, ldn))
    p.check("newhat")
    p.run("simm_is_divergence", lambda d: (d["SX"], d["SM"], d["SF0"], d["SPHI"], d["a2"], nch, F, N,
                                           ldn, d["ws"], d["isd"]))
    p.check("isd", rtol=1e-5)
    # the padding columns of the work planes are zero (they are contracted over n by the GEMMs)
    lead = p.gpu["lead"].cpu().numpy()
    assert (lead[:, N:ldn] == 0).all() and (lead[:, ldn + N:] == 0).all()


def test_alpha_update():
    rng = np.random.default_rng(3)
    F, N = 65, 333
    ldn, pl = planes(rng, F, N, 2)
    p = Pair(SM=pl(2), SX=pl(2), SF0=pl(1), SPHI=pl(1), ws=np.zeros(148 * 4 * 4),
             alpha=np.array([0.4, 0.6]), a2=np.array([0.16, 0.36], np.float32))
    p.run("simm_alpha_update", lambda d: (d["SX"], d["SM"], d["SF0"], d["SPHI"], F, N, ldn, 0.8,
                                          d["ws"], d["alpha"], d["a2"]))
    p.check("alpha", rtol=1e-6)
    p.check("a2", rtol=1e-6)


@pytest.mark.parametrize("nch,omega,floor", [(1, 1.0, 0.0), (1, 0.6, 1e-20), (2, 1.0, 0.0), (2, 0.9, 0.0)])
def test_update_rows(nch, omega, floor):
    rng = np.random.default_rng(nch)
    rows, N = 7, 203
    ldn, ldr = ru4(N), ru4(rows)
    C = (rng.random((ldr, 2 * nch * ldn)) + 0.1).astype(np.float32)
    w = None if nch == 1 else (rng.random((2, ldr)) + 0.1).astype(np.float32)
    arrays = dict(theta=rng.random((ldr, ldn)).astype(np.float32), C=C)
    if w is not None:
        arrays["w"] = w
    p = Pair(**arrays)
    p.run("simm_update_rows", lambda d: (d["theta"], d["C"], nch, ldn, d.get("w"), omega, floor, rows, N))
    p.check("theta", rtol=5e-6)


def test_normalise_and_scale():
    rng = np.random.default_rng(9)
    K, N, rows = 5, 301, 9
    ldn = ru4(N)
    H = np.zeros((ru4(K), ldn), np.float32)
    H[:K, :N] = rng.random((K, N))
    H[:K, 17] = 0  # a column that sums to zero is left alone (SIMM.py:325)
    P = np.zeros((rows, ldn), np.float32)
    P[:, :N] = rng.random((rows, N))
    p = Pair(H=H, H2=H.copy(), P=P, P2=P.copy(), s=np.zeros(ldn, np.float32), s2=np.zeros(ldn, np.float32),
             rs=(rng.random(ru4(K)) + 0.5).astype(np.float32), sr=(rng.random(ru4(rows)) + 0.5).astype(np.float32))
    p.run("simm_hphi_normalise", lambda d: (d["H"], K, None, N, d["s"]))
    p.check("H"); p.check("s")
    p.run("simm_hphi_normalise", lambda d: (d["H2"], K, d["rs"], N, d["s2"]))
    p.check("H2"); p.check("s2")
    p.run("simm_scale_columns", lambda d: (d["P"], rows, N, d["s2"]))
    p.check("P")
    p.run("simm_scale_rows", lambda d: (d["P2"], rows, N, d["sr"]))
    p.check("P2")


def test_small_matrix_updates():
    rng = np.random.default_rng(21)
    F, P_, K, R = 131, 30, 4, 6
    ldp, ldk, ldr = ru4(P_), ru4(K), ru4(R)

    def mat(r, c, rr, cc):
        a = np.zeros((r, c), np.float32)
        a[:rr, :cc] = rng.random((rr, cc)) + 0.05
        return a
    for nch in (1, 2):
        b2 = mat(2, ldr, 2, R)
        be = np.zeros((2, ldr))
        be[0, :R] = rng.random(R) * 0.8 + 0.1
        be[1, :R] = 1 - be[0, :R]
        p = Pair(HG=mat(ldp, ldk, P_, K), WG=mat(F, ldp, F, P_), tn=mat(F, ldk, F, K), td=mat(F, ldk, F, K),
                 sk=np.zeros(ldk, np.float32), WM=mat(F, ldr, F, R),
                 D=np.stack([mat(F, ldr, F, R) for _ in range(2 * nch)]), b2=b2,
                 sr=np.zeros(ldr, np.float32), beta=be, b2o=np.zeros((2, ldr), np.float32),
                 WMs=np.full((nch, F, ldr), 3, np.float32))
        p.run("simm_hgamma_update", lambda d: (d["HG"], d["WG"], d["tn"], d["td"], F, P_, K, 0.9, d["sk"]))
        p.check("HG", rtol=5e-6); p.check("sk", rtol=5e-6)
        p.run("simm_wm_scaled", lambda d: (d["WM"], R, d["b2"] if nch == 2 else None, nch, F, d["WMs"]))
        p.check("WMs")
        if nch == 2:
            p.run("simm_beta_update", lambda d: (d["WM"], R, d["D"], F, 1.0, d["beta"], d["b2o"]))
            p.check("beta", rtol=1e-6); p.check("b2o", rtol=1e-6)
        p.run("simm_wm_update", lambda d: (d["WM"], R, d["D"], nch, d["b2"] if nch == 2 else None,
                                           nch == 1, 0.8, F, d["sr"]))
        p.check("WM", rtol=5e-6); p.check("sr", rtol=5e-6)


@pytest.mark.parametrize("M,N,K,transB", [(1025, 40, 10340, True), (1025, 4, 5004, True),
                                          (300, 130, 4100, False), (129, 7, 96, True)])
def test_gemm_splitk(M, N, K, transB):
    k = ck()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    Bs = np.ascontiguousarray(B.T) if transB else B
    ldb = ru4(Bs.shape[1])
    Bp = np.zeros((Bs.shape[0], ldb), np.float32)
    Bp[:, :Bs.shape[1]] = Bs
    Ad, Bd = torch.tensor(A).cuda(), torch.tensor(Bp).cuda()
    Cd = torch.full((M, ru4(N)), -3.0, dtype=torch.float32, device="cuda")
    ws = torch.empty(max(k.gemm_splitk_workspace_bytes(M, N, K) // 4, 4), dtype=torch.float32, device="cuda")
    ws.fill_(float("nan"))
    k.gemm_view(Ad, Bd, Cd, M, N, K, transB=transB, workspace=ws)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = Cd.cpu().numpy()
    assert rel_err(got[:, :N], ref) < 2e-5
    if k.gemm_splitk_workspace_bytes(M, N, K) > 0:
        assert (got[:, N:] == 0).all()


def load():
    return np.load(os.path.join(GOLDEN, "simm.npz"))


def test_mono_simm_matches_reference():
    from pyfasst_b200.SeparateLeadStereo.SIMM import SIMM as simm_mod
    g = load()
    SX = 0.5 * (g["SXR"] + g["SXL"])
    res = simm_mod.SIMM(SX, g["WF0"], g["WGAMMA"], numberOfFilters=g["HGAMMA0"].shape[1],
                        numberOfAccompanimentSpectralShapes=1, HGAMMA0=g["HGAMMA0"],
                        HPHI0=g["HPHI0"], HF00=g["HF00"], WM0=g["WM0"][:, :1], HM0=g["HM0"][:1],
                        numberOfIterations=4, verbose=False)
    for nm, a in zip(("HGAMMA", "HPHI", "HF0", "HM", "WM"), res):
        assert rel_err(a, g["mono_" + nm]) < RTOL, nm


def test_stereo_simm_matches_reference():
    from pyfasst_b200.SeparateLeadStereo.SIMM import SIMM as simm_mod
    g = load()
    R = g["WM0"].shape[1]
    np.random.seed(5)
    res = simm_mod.Stereo_SIMM(g["SXR"], g["SXL"], g["WF0"], g["WGAMMA"],
                               numberOfFilters=g["HGAMMA0"].shape[1],
                               numberOfAccompanimentSpectralShapes=R, HGAMMA0=g["HGAMMA0"],
                               HPHI0=g["HPHI0"], HF00=g["HF00"], WM0=g["WM0"], HM0=g["HM0"],
                               numberOfIterations=4, verbose=False, computeError=True)
    names = ("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR", "betaL", "HM", "WM")
    for nm, a in zip(names, res):
        assert rel_err(np.asarray(a), np.asarray(g["st_" + nm])) < RTOL, nm
    np.testing.assert_allclose(res[9], g["st_recoError"], rtol=1e-4, atol=1e-3)


def test_stereo_simm_larger_against_oracle():
    """A problem with the F0-dictionary width of config 3 (NF0 = 480: the tensor-core path with
    several M tiles, split-K contractions over > 1000 frames), one iteration, against the oracle
    -- the north_star bar: relative error of every factor <= 1e-4 after one iteration."""
    from pyfasst_b200.simm_engine import SimmEngine
    rng = np.random.default_rng(77)
    F, N, NF0, P_, K, R = 513, 1203, 480, 30, 4, 40
    WF0 = np.abs(rng.standard_normal((F, NF0)))
    WF0 /= WF0.sum(axis=0)
    WG = np.abs(rng.standard_normal((F, P_)))
    HG0, HPHI0 = np.abs(rng.standard_normal((P_, K))), np.abs(rng.standard_normal((K, N)))
    HF00, HM0 = np.abs(rng.standard_normal((NF0, N))), np.abs(rng.standard_normal((R, N)))
    WM0 = np.abs(rng.standard_normal((F, R)))
    SXR = np.abs(rng.standard_normal((F, N))) ** 2 * np.linspace(3, .1, F)[:, None]
    SXL = np.abs(rng.standard_normal((F, N))) ** 2 * np.linspace(2, .2, F)[:, None]
    beta0 = rng.random(R)
    ref = so.stereo_simm(SXR, SXL, WF0, WG, HG0, HPHI0, HF00, WM0, HM0, beta0, numberOfIterations=1)
    eng = SimmEngine(ck(), [SXR, SXL], WF0, WG, HG0, HPHI0, HF00, WM0, HM0, betaR=beta0, n_iter=1)
    eng.iterate()
    r = eng.results()
    got = (r["alphaR"], r["alphaL"], r["HGAMMA"], r["HPHI"], r["HF0"], np.diag(r["betaR"]),
           np.diag(r["betaL"]), r["HM"], r["WM"])
    for nm, a, b in zip("alphaR alphaL HGAMMA HPHI HF0 betaR betaL HM WM".split(), got, ref):
        assert rel_err(np.asarray(a), np.asarray(b)) < 1e-4, nm
