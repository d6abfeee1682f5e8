"""Every CUDA kernel of the C ABI against its NumPy specification (tests/fake_kernels.py),
on a B200.  Called through the C ABI (ctypes) exactly as the product does."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from tests.fake_kernels import FakeKernels

pytestmark = pytest.mark.gpu

DTYPES = [torch.float64, torch.float32]


@pytest.fixture(scope="module")
def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


@pytest.fixture(scope="module")
def fk():
    return FakeKernels()


def tol(dt, f64=1e-10, f32=2e-5):
    return f64 if dt == torch.float64 else f32


def both(t):
    """(cpu tensor, cuda copy)"""
    return t, t.clone().cuda()


def rnd(rng, shape, dt, positive=False, pad_to=None):
    a = rng.standard_normal(shape)
    if positive:
        a = np.abs(a) + 0.05
    if pad_to is not None:
        full = np.zeros(tuple(shape[:-1]) + (pad_to,))
        full[..., :shape[-1]] = a
        a = full
    return torch.tensor(a).to(dt)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.complex128), np.asarray(b, dtype=np.complex128)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def problem(rng, dt, F, N, J, rank, conv=True, consistent=True):
    """Random E-step inputs.  With `consistent` the observations follow the model
    (x = sum_r a_r s_r + noise, s_r ~ CN(0, v)), like real data after a few iterations;
    otherwise X is unrelated to Sigma (adversarial: |Sigma^-1 x| ~ |x| / noise)."""
    ld = (N + 31) // 32 * 32
    R = J * rank
    src = [j for j in range(J) for _ in range(rank)]
    Vn = np.abs(rng.standard_normal((J, F, N))) + 0.05
    A = rng.standard_normal((R, 2, F)) + 1j * rng.standard_normal((R, 2, F))
    if not conv:
        A = np.broadcast_to(rng.standard_normal((R, 2, 1)), (R, 2, F)) + 0j
    noise = np.abs(rng.standard_normal(F)) * 0.01 + 1e-3
    if consistent:
        cn = lambda shape: (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)
        Xc = np.sqrt(noise)[None, :, None] * cn((2, F, N))
        for r in range(R):
            s = np.sqrt(Vn[src[r]]) * cn((F, N))
            Xc = Xc + A[r][:, :, None] * s[None]
        Xn = np.array([Xc[0].real, Xc[0].imag, Xc[1].real, Xc[1].imag])
    else:
        Xn = rng.standard_normal((4, F, N))
    pad = lambda a: np.concatenate([a, np.zeros(a.shape[:-1] + (ld - N,))], axis=-1)
    X = torch.tensor(pad(Xn)).to(dt)
    V = torch.tensor(pad(Vn)).to(dt)
    A = torch.tensor(np.ascontiguousarray(A))
    noise = torch.tensor(noise)
    return ld, R, src, X, V, A, noise


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("consistent", [True, False])
@pytest.mark.parametrize("F,N,J,rank", [(5, 77, 1, 1), (33, 1000, 3, 1), (17, 2600, 4, 2),
                                        (9, 515, 2, 3), (3, 4, 6, 1), (6, 1531, 4, 1),
                                        (3, 5, 4, 2)])
def test_estep_stereo(ck, fk, dt, F, N, J, rank, consistent):
    rng = np.random.default_rng(F * 1000 + N)
    ld, R, src, X, V, A, noise = problem(rng, dt, F, N, J, rank, consistent=consistent)
    outs = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        hatW = torch.zeros((J, F, ld), dtype=dt, device=dev)
        Rss = torch.zeros((F, R, R), dtype=torch.complex128, device=dev)
        Rxs = torch.zeros((F, 2, R), dtype=torch.complex128, device=dev)
        ll = torch.zeros(F, dtype=torch.float64, device=dev)
        ws = torch.zeros((k.estep_workspace_bytes(J, F, N, k.dtype_code(V)) + 7) // 8,
                         dtype=torch.float64, device=dev)
        k.estep_stereo(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), N, hatW, Rss, Rxs, ll, ws)
        outs.append([t.cpu().numpy() for t in (hatW, Rss, Rxs, ll)])
    (hw0, rss0, rxs0, ll0), (hw1, rss1, rxs1, ll1) = outs
    # storage may be float32 but the per-bin algebra AND the moment sums are float64: only the
    # final rounding of hat_W differs
    assert rel(hw1[:, :, :N], hw0[:, :, :N]) < tol(dt, f32=1e-6)
    assert (hw1[:, :, N:] == 0).all(), "padding frames must stay zero"
    # hat_Rxs comes from the identity x y^H = Sigma M + I (T_j = sv_j I + s2 Z_j + sum_l R_l S_lj),
    # exact to 1e-16 cond(Sigma); the adversarial inputs (|y| ~ |x| / noise) stress that
    t = 1e-9 if consistent else 1e-7
    assert rel(rss1, rss0) < t
    assert rel(rxs1, rxs0) < t
    assert_allclose(ll1, ll0, rtol=tol(dt, f64=1e-12, f32=1e-6),
                    atol=0.0 if dt == torch.float64 else 1e-6 * N)
    assert_allclose(rss1, np.conj(np.transpose(rss1, (0, 2, 1))), atol=1e-14 * np.abs(rss1).max())


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("F,N,J,rank,quiet", [(17, 2600, 4, 2, False), (9, 1300, 3, 1, True),
                                              (5, 77, 1, 2, False), (4, 700, 6, 1, True)])
def test_estep_stereo_inst(ck, fk, dt, F, N, J, rank, quiet):
    """Real mixing vectors (instantaneous mixing): hat_W, the log-likelihood and the REAL parts of
    the statistics equal those of the general kernel; the imaginary parts are returned as zero.
    quiet: some rows inside the determinant clamp."""
    rng = np.random.default_rng(F * 77 + N)
    ld, R, src, X, V, A, noise = problem(rng, dt, F, N, J, rank)
    A = torch.complex(A.real.contiguous(), torch.zeros_like(A.real))
    if quiet:
        scale = np.ones(F)
        scale[:2], scale[2:4] = 1e-7, 3e-5
        sc = torch.tensor(scale)
        X = (X.to(torch.float64) * torch.sqrt(sc)[None, :, None]).to(dt)
        V = (V.to(torch.float64) * sc[None, :, None]).to(dt)
        noise = noise * sc
    outs = []
    for k, dev, fn in ((fk, "cpu", "estep_stereo"), (ck, "cuda", "estep_stereo_inst")):
        hatW = torch.zeros((J, F, ld), dtype=dt, device=dev)
        Rss = torch.zeros((F, R, R), dtype=torch.complex128, device=dev)
        Rxs = torch.zeros((F, 2, R), dtype=torch.complex128, device=dev)
        ll = torch.zeros(F, dtype=torch.float64, device=dev)
        ws = torch.zeros((k.estep_workspace_bytes(J, F, N, k.dtype_code(V)) + 7) // 8,
                         dtype=torch.float64, device=dev)
        getattr(k, fn)(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), N, hatW, Rss, Rxs, ll, ws)
        outs.append([t.cpu().numpy() for t in (hatW, Rss, Rxs, ll)])
    (hw0, rss0, rxs0, ll0), (hw1, rss1, rxs1, ll1) = outs
    for f in range(F):
        assert rel(hw1[:, f, :N], hw0[:, f, :N]) < tol(dt, f32=1e-6), f
        assert rel(rss1[f].real, rss0[f].real) < 1e-8, f
        assert rel(rxs1[f].real, rxs0[f].real) < 1e-8, f
    assert (rss1.imag == 0).all() and (rxs1.imag == 0).all()
    assert (hw1[:, :, N:] == 0).all()
    assert_allclose(ll1, ll0, rtol=tol(dt, f64=1e-12, f32=1e-6),
                    atol=0.0 if dt == torch.float64 else 1e-6 * N)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("J", [3, 4])
def test_estep_stereo_determinant_clamp(ck, fk, dt, J):
    """Quiet rows: det Sigma < 1e-10 activates the reference's clamp (signalTools.py:183-188), where
    Sigma_c^-1 is not the inverse of Sigma and the kernel's identity for x y^H needs its correction
    term; rows 0-2 are clamped in every bin, rows 3-5 in some, the rest in none."""
    rng = np.random.default_rng(77)
    F, N, rank = 9, 1300, 2
    ld, R, src, X, V, A, noise = problem(rng, dt, F, N, J, rank)
    scale = np.ones(F)
    scale[:3], scale[3:6] = 1e-7, 3e-5
    sc = torch.tensor(scale)
    X = (X.to(torch.float64) * torch.sqrt(sc)[None, :, None]).to(dt)
    V = (V.to(torch.float64) * sc[None, :, None]).to(dt)
    V[:, 3:6, ::3] *= 1e-3
    X[:, 3:6, ::3] *= 1e-2
    noise = noise * sc
    outs = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        hatW = torch.zeros((J, F, ld), dtype=dt, device=dev)
        Rss = torch.zeros((F, R, R), dtype=torch.complex128, device=dev)
        Rxs = torch.zeros((F, 2, R), dtype=torch.complex128, device=dev)
        ll = torch.zeros(F, dtype=torch.float64, device=dev)
        ws = torch.zeros((k.estep_workspace_bytes(J, F, N, k.dtype_code(V)) + 7) // 8,
                         dtype=torch.float64, device=dev)
        k.estep_stereo(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), N, hatW, Rss, Rxs, ll, ws)
        outs.append([t.cpu().numpy() for t in (hatW, Rss, Rxs, ll)])
    (hw0, rss0, rxs0, ll0), (hw1, rss1, rxs1, ll1) = outs
    for f in range(F):  # per row: the rows differ by many orders of magnitude
        assert rel(hw1[:, f, :N], hw0[:, f, :N]) < tol(dt, f32=1e-6), f
        assert rel(rss1[f], rss0[f]) < 1e-8, f
        assert rel(rxs1[f], rxs0[f]) < 1e-8, f
    assert_allclose(ll1, ll0, rtol=tol(dt, f64=1e-12, f32=1e-6), atol=0.0 if dt == torch.float64 else 1e-6 * N)


@pytest.mark.parametrize("dt", DTYPES)
def test_wiener_stereo(ck, fk, dt):
    rng = np.random.default_rng(5)
    F, N, J, rank = 19, 1030, 3, 2
    ld, R, src, X, V, A, noise = problem(rng, dt, F, N, J, rank)
    groups = [1, -1, 0]
    outs = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        Y = torch.zeros((2 * 4, F, ld), dtype=dt, device=dev)
        ws = torch.zeros(4096, dtype=torch.float64, device=dev)
        k.wiener_stereo(X.to(dev), V.to(dev), A.to(dev), src, noise.to(dev), groups, 2, N, Y, ws)
        outs.append(Y.cpu().numpy())
    assert rel(outs[1][:, :, :N], outs[0][:, :, :N]) < tol(dt, f32=1e-6)
    assert (outs[1][:, :, N:] == 0).all()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("F,K,N", [(7, 3, 50), (130, 4, 1001), (65, 32, 700), (40, 40, 333),
                                   (300, 32, 2500), (129, 5, 1025), (64, 17, 4000)])
def test_spec_power(ck, fk, dt, F, K, N):
    rng = np.random.default_rng(K)
    ld = (N + 31) // 32 * 32
    W = rnd(rng, (F, K), dt, positive=True)
    H = rnd(rng, (K, N), dt, positive=True, pad_to=ld)
    for acc in (False, True):
        outs = []
        for k, dev in ((fk, "cpu"), (ck, "cuda")):
            V = torch.full((F, ld), 2.0, dtype=dt, device=dev)
            V[:, N:] = 0
            k.spec_power(W.to(dev), H.to(dev), V, N, acc)
            outs.append(V.cpu().numpy())
        assert rel(outs[1], outs[0]) < tol(dt, f32=1e-6)
        assert (outs[1][:, N:] == 0).all()


@pytest.mark.parametrize("dt", DTYPES)
def test_small_matmul(ck, fk, dt):
    rng = np.random.default_rng(1)
    A, B = rnd(rng, (37, 5), dt), rnd(rng, (5, 9), dt)
    outs = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        C = torch.zeros((37, 9), dtype=dt, device=dev)
        k.small_matmul(A.to(dev), B.to(dev), C)
        outs.append(C.cpu().numpy())
    assert rel(outs[1], outs[0]) < tol(dt, f32=1e-6)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("F,K,N", [(9, 4, 300), (130, 4, 5000), (33, 32, 2100), (20, 7, 130),
                                   (18, 16, 600), (300, 32, 2500), (257, 20, 1100)])
def test_fb_contract(ck, fk, dt, F, K, N):
    rng = np.random.default_rng(N)
    ld = (N + 31) // 32 * 32
    hatW = rnd(rng, (F, N), dt, positive=True, pad_to=ld)
    P = rnd(rng, (F, N), dt, positive=True, pad_to=ld)
    O = rnd(rng, (F, N), dt, positive=True, pad_to=ld)
    P[0, :5] = 0  # exercises the eps clamp
    G = rnd(rng, (K, N), dt, positive=True, pad_to=ld)
    for same in (False, True):
        res = []
        for k, dev in ((fk, "cpu"), (ck, "cuda")):
            chunk, nsplit = k.fb_plan(F, K, N, k.dtype_code(hatW))
            pn = torch.zeros((nsplit, F, K), dtype=torch.float64, device=dev)
            pd = torch.zeros((nsplit, F, K), dtype=torch.float64, device=dev)
            Pd = P.to(dev)
            Od = Pd if same else O.to(dev)
            k.fb_contract(hatW.to(dev), Pd, Od, G.to(dev), N, pn, pd, chunk, nsplit)
            num = torch.zeros((F, K), dtype=torch.float64, device=dev)
            den = torch.zeros((F, K), dtype=torch.float64, device=dev)
            k.sum_splits(pn, num)
            k.sum_splits(pd, den)
            res.append((num.cpu().numpy(), den.cpu().numpy()))
        assert rel(res[1][0], res[0][0]) < tol(dt)
        assert rel(res[1][1], res[0][1]) < tol(dt)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("scratch", [False, True])
@pytest.mark.parametrize("F,K,N", [(9, 4, 300), (130, 4, 5000), (257, 32, 700), (70, 7, 130),
                                   (300, 32, 2500), (129, 20, 1100)])
def test_tw_contract(ck, fk, dt, F, K, N, scratch):
    """With a scratch plane, float32 planes and a large enough problem the tensor-core
    (tcgen05, 3xTF32) path runs; the CUDA-core path otherwise."""
    rng = np.random.default_rng(N + 1)
    ld = (N + 31) // 32 * 32
    hatW = rnd(rng, (F, N), dt, positive=True, pad_to=ld)
    O = rnd(rng, (F, N), dt, positive=True, pad_to=ld)
    W = rnd(rng, (F, K), dt, positive=True)
    H = rnd(rng, (K, N), dt, positive=True, pad_to=ld)
    res = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        fchunk, fsplit = k.tw_plan(F, K, N, k.dtype_code(hatW))
        pn = torch.zeros((fsplit, K, ld), dtype=torch.float64, device=dev)
        pd = torch.zeros((fsplit, K, ld), dtype=torch.float64, device=dev)
        sc = torch.zeros((F, ld), dtype=dt, device=dev) if scratch else None
        k.tw_contract(hatW.to(dev), O.to(dev), W.to(dev), H.to(dev), N, pn, pd, fchunk, fsplit, sc)
        num = torch.zeros((K, ld), dtype=torch.float64, device=dev)
        den = torch.zeros((K, ld), dtype=torch.float64, device=dev)
        k.sum_splits(pn, num)
        k.sum_splits(pd, den)
        res.append((num.cpu().numpy(), den.cpu().numpy()))
    assert rel(res[1][0][:, :N], res[0][0][:, :N]) < tol(dt)
    assert rel(res[1][1][:, :N], res[0][1][:, :N]) < tol(dt)
    assert (res[1][0][:, N:] == 0).all() and (res[1][1][:, N:] == 0).all()


def test_hot_kernels_are_deterministic(ck):
    """Run-to-run bit identity of the kernels that synchronise by hand (cp.async ring without
    barriers, mbarriers / named barriers around tcgen05 MMAs, TMA stores, DMMA tiles): a data race
    shows up as a result that changes between launches.  (compute-sanitizer is refused on this
    GPU pool, profiles/r02/sanitizer_closed_on_this_pool.log.)  Sizes: several CTAs per SM and
    several waves."""
    dt = torch.float32
    rng = np.random.default_rng(2024)
    F, N, K = 300, 40000, 32
    ld = (N + 31) // 32 * 32
    hatW = rnd(rng, (F, N), dt, positive=True, pad_to=ld).cuda()
    O = rnd(rng, (F, N), dt, positive=True, pad_to=ld).cuda()
    W = rnd(rng, (F, K), dt, positive=True).cuda()
    H = rnd(rng, (K, N), dt, positive=True, pad_to=ld).cuda()

    def tw():
        fchunk, fsplit = ck.tw_plan(F, K, N, ck.dtype_code(hatW))
        pn = torch.zeros((fsplit, K, ld), dtype=torch.float64, device="cuda")
        pd = torch.zeros((fsplit, K, ld), dtype=torch.float64, device="cuda")
        ck.tw_contract(hatW, O, W, H, N, pn, pd, fchunk, fsplit, torch.zeros((F, ld), dtype=dt, device="cuda"))
        return pn, pd

    def fb():
        chunk, nsplit = ck.fb_plan(F, K, N, ck.dtype_code(hatW))
        pn = torch.zeros((nsplit, F, K), dtype=torch.float64, device="cuda")
        pd = torch.zeros((nsplit, F, K), dtype=torch.float64, device="cuda")
        ck.fb_contract(hatW, O, O, H, N, pn, pd, chunk, nsplit)
        return pn, pd

    def power():
        V = torch.zeros((F, ld), dtype=dt, device="cuda")
        ck.spec_power(W, H, V, N, False)
        return (V,)

    def estep(which, J, rank, I):
        if I == 2:
            ldx, R, src, X, V, A, noise = problem(np.random.default_rng(5), dt, 40, 30000, J, rank)
            if which == "estep_stereo_inst":
                A = torch.complex(A.real.contiguous(), torch.zeros_like(A.real))
        else:
            from tests.test_multichannel_gpu import problem as problem_multi
            ldx, R, src, X, V, A, noise = problem_multi(np.random.default_rng(5), dt, I, 40, 30000, J, rank)
        X, V, A, noise = X.cuda(), V.cuda(), A.cuda(), noise.cuda()
        Nn, Fn = 30000, 40

        def run():
            hw = torch.zeros((J, Fn, ldx), dtype=dt, device="cuda")
            Rss = torch.zeros((Fn, R, R), dtype=torch.complex128, device="cuda")
            Rxs = torch.zeros((Fn, I, R), dtype=torch.complex128, device="cuda")
            ll = torch.zeros(Fn, dtype=torch.float64, device="cuda")
            if I == 2:
                ws = torch.zeros((ck.estep_workspace_bytes(J, Fn, Nn, ck.dtype_code(V)) + 7) // 8,
                                 dtype=torch.float64, device="cuda")
            else:
                ws = torch.zeros((ck.estep_multi_workspace_bytes(I, J, Fn, Nn) + 7) // 8,
                                 dtype=torch.float64, device="cuda")
            getattr(ck, which)(X, V, A, src, noise, Nn, hw, Rss, Rxs, ll, ws)
            return hw, torch.view_as_real(Rss), torch.view_as_real(Rxs), ll
        return run

    cases = {"tw_contract": tw, "fb_contract": fb, "spec_power": power,
             "estep_stereo": estep("estep_stereo", 4, 2, 2),
             "estep_stereo_inst": estep("estep_stereo_inst", 4, 2, 2),
             "estep_multi": estep("estep_multi", 4, 2, 4)}
    for name, fn in cases.items():
        first = [t.clone() for t in fn()]
        for _ in range(3):
            again = fn()
            for a, b in zip(first, again):
                assert torch.equal(a, b), name


@pytest.mark.parametrize("dt", DTYPES)
def test_mult_update_and_scaling(ck, fk, dt):
    rng = np.random.default_rng(3)
    rows, cols, ld = 5, 70, 96
    theta = rnd(rng, (rows, cols), dt, positive=True, pad_to=ld)
    num = rnd(rng, (rows, cols), torch.float64, positive=True, pad_to=ld)
    den = rnd(rng, (rows, cols), torch.float64, positive=True, pad_to=ld)
    den[0, 0] = 0.0
    s_row = rnd(rng, (rows,), torch.float64, positive=True)
    s_col = rnd(rng, (cols,), torch.float64, positive=True)
    for omega in (1.0, 0.7):
        res = []
        for k, dev in ((fk, "cpu"), (ck, "cuda")):
            th = theta.clone().to(dev)
            k.mult_update(th, num.to(dev), den.to(dev), rows, cols, omega)
            tot = torch.zeros(1, dtype=torch.float64, device=dev)
            k.scale_matrix(th, rows, cols, s_row.to(dev), True, False, tot)
            k.scale_matrix(th, rows, cols, s_col.to(dev), False, True)
            res.append((th.cpu().numpy(), tot.cpu().item()))
        assert rel(res[1][0], res[0][0]) < tol(dt, f32=1e-6)
        assert abs(res[1][1] - res[0][1]) < 1e-5 * abs(res[0][1])


@pytest.mark.parametrize("conv", [False, True])
def test_spatial_updates(ck, fk, conv):
    rng = np.random.default_rng(9)
    F, R = 41, 4
    Z = rng.standard_normal((F, R, 6)) + 1j * rng.standard_normal((F, R, 6))
    Rss = torch.tensor(np.einsum("fri,fsi->frs", Z, np.conj(Z)) / 6 + 0.1 * np.eye(R))
    Rxs = torch.tensor(rng.standard_normal((F, 2, R)) + 1j * rng.standard_normal((F, 2, R)))
    A0 = rng.standard_normal((R, 2, F)) + 1j * rng.standard_normal((R, 2, F))
    res = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        A = torch.tensor(A0).to(dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        if conv:
            k.mix_conv_solve(Rss.to(dev), Rxs.to(dev), A, flags)
        else:
            upd, oth = [0, 1, 3], [2]
            stats = torch.zeros(2 * 3 + 9, dtype=torch.float64, device=dev)
            k.mix_inst_stats(Rss.to(dev), Rxs.to(dev), A, upd, oth, stats)
            k.mix_inst_solve(stats, F, upd, A, flags)
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        src = [0, 0, 1, 1]
        k.spat_energy(A, src, 2, sums)
        counts = torch.tensor([2.0 * 2 * F, 2.0 * 2 * F], dtype=torch.float64, device=dev)
        k.spat_scale(A, src, sums, counts)
        res.append((A.cpu().numpy(), sums.cpu().numpy(), int(flags.cpu().item())))
    assert res[1][2] == 0
    assert rel(res[1][0], res[0][0]) < 1e-10
    assert_allclose(res[1][1], res[0][1], rtol=1e-10)


def test_singular_flag(ck):
    F, R = 8, 2
    Rss = torch.zeros((F, R, R), dtype=torch.complex128, device="cuda")
    Rxs = torch.ones((F, 2, R), dtype=torch.complex128, device="cuda")
    A = torch.ones((R, 2, F), dtype=torch.complex128, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    ck.mix_conv_solve(Rss, Rxs, A, flags)
    assert int(flags.cpu().item()) & 1


@pytest.mark.parametrize("dt", DTYPES)
def test_renorm_kernels(ck, fk, dt):
    rng = np.random.default_rng(4)
    F, Kb, Kw = 50, 6, 5
    FB0 = rnd(rng, (F, Kb), dt, positive=True)
    FW0 = rnd(rng, (Kb, Kw), dt, positive=True)
    sums = torch.tensor([3.0, 8.0], dtype=torch.float64)
    counts = torch.tensor([4.0, 2.0], dtype=torch.float64)
    res = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        FB, FW = FB0.clone().to(dev), FW0.clone().to(dev)
        colmax = torch.zeros(8, dtype=torch.float64, device=dev)
        w = torch.zeros(8, dtype=torch.float64, device=dev)
        w2 = torch.zeros(8, dtype=torch.float64, device=dev)
        k.fb_scale_colmax(FB, sums.to(dev), counts.to(dev), 1, colmax)
        k.fw_renorm(FW, colmax, w, w2)
        k.scale_matrix(FB, F, Kb, w, False, True)
        totals = torch.tensor([0.0, 5.0], dtype=torch.float64, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        k.check_totals(totals, 1e-10, flags)
        res.append([t.cpu().numpy() for t in (FB, FW, colmax, w, w2, totals, flags)])
    for a, b in zip(res[1], res[0]):
        assert rel(a, b) < tol(dt, f32=1e-6)
    assert res[1][6][0] == 2


def test_glue_kernels(ck, fk):
    rng = np.random.default_rng(6)
    F = 300
    s0 = torch.tensor(np.abs(rng.standard_normal(F)))
    s1 = torch.tensor(np.abs(rng.standard_normal(F)) * 0.1)
    llf = torch.tensor(rng.standard_normal(F))
    res = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        it = torch.tensor([2], dtype=torch.int32, device=dev)
        noise = torch.zeros(F, dtype=torch.float64, device=dev)
        k.noise_anneal(s0.to(dev), s1.to(dev), it, 7, noise)
        lls = torch.zeros(1, dtype=torch.float64, device=dev)
        k.ll_reduce(llf.to(dev), lls)
        logl = torch.ones(7, dtype=torch.float64, device=dev)
        k.ll_store(lls, 123.0, logl, it, True)
        res.append([t.cpu().numpy() for t in (noise, lls, logl, it)])
    for a, b in zip(res[1], res[0]):
        assert_allclose(a, b, rtol=1e-13)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("L,wlen,hop", [(3001, 256, 64), (5000, 2048, 512), (700, 64, 16),
                                        (9000, 1024, 256)])
def test_stft_istft(ck, fk, dt, L, wlen, hop):
    rng = np.random.default_rng(L)
    nch = 2
    pcm = torch.tensor(rng.standard_normal((nch, L)))
    window = torch.tensor(np.hanning(wlen))
    nfft = wlen
    N = int(np.ceil(L / float(hop)) + 2)
    ld = (N + 31) // 32 * 32
    F = nfft // 2 + 1
    total = (N - 1) * hop + wlen
    norm = np.zeros(total)
    for n in range(N):
        norm[n * hop:n * hop + wlen] += np.hanning(wlen) ** 2
    norm[norm == 0] = 1.0
    norm = torch.tensor(norm)
    res = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        X = torch.zeros((2 * nch, F, ld), dtype=dt, device=dev)
        psd = torch.zeros(F, dtype=torch.float64, device=dev)
        k.stft(pcm.to(dev), window.to(dev), hop, nfft, X, N, psd)
        out = torch.zeros((nch, L), dtype=torch.float64, device=dev)
        pcm16 = torch.zeros((L, nch), dtype=torch.int16, device=dev)
        k.istft(X, N, window.to(dev), norm.to(dev), hop, nfft, out, pcm16, 1000.0)
        res.append([t.cpu().numpy() for t in (X, psd, out, pcm16)])
    (X0, p0, o0, q0), (X1, p1, o1, q1) = res
    t = tol(dt, f64=1e-12, f32=2e-7)
    scale = np.abs(X0).max()
    assert np.abs(X1[:, :, :N] - X0[:, :, :N]).max() < t * scale
    assert (X1[:, :, N:] == 0).all()
    assert_allclose(p1, p0, rtol=1e-6 if dt == torch.float32 else 1e-12)
    assert np.abs(o1 - o0).max() < 10 * t * np.abs(o0).max()
    # the round trip reconstructs the input (stft.py:71-131 normalisation)
    assert np.abs(o1 - pcm.numpy()).max() < (1e-4 if dt == torch.float32 else 1e-10)
    assert (np.abs(q1.astype(int) - q0.astype(int)) <= 1).all()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("world,nsplit,K,Kmax", [(2, 3, 5, 8), (8, 1, 32, 32), (4, 2, 7, 7)])
def test_tw_pack_chunks(ck, fk, dt, world, nsplit, K, Kmax):
    """Split sums of the TW numerators / denominators, chunk-major for the reduce-scatter over the
    frames (frequency partition)."""
    rng = np.random.default_rng(world * 10 + K)
    ld = 32 * world * 3
    pn, pd = torch.tensor(rng.standard_normal((nsplit, K, ld))), torch.tensor(rng.random((nsplit, K, ld)))
    outs = []
    for k, dev in ((fk, "cpu"), (ck, "cuda")):
        out = torch.full((world, 2, Kmax, ld // world), 7.0, dtype=dt, device=dev)
        k.tw_pack_chunks(pn.to(dev), pd.to(dev), out, world)
        outs.append(out.cpu().numpy())
    assert_allclose(outs[1][:, :, :K], outs[0][:, :, :K], rtol=0, atol=tol(dt, f64=1e-14, f32=1e-6))
    assert (outs[1][:, :, K:] == 7.0).all(), "rows beyond K are not touched"


def test_library_errors(ck):
    from pyfasst_b200 import _lib
    lib = _lib.load_library()
    # invalid arguments are reported through the return code, never a crash
    rc = lib.pf_spec_power(None, 1, None, 3, None, 3, 1, 1, 1, 0, 0, None)
    assert rc == -1 and b"multiples of 4" in lib.pf_last_error()
    with pytest.raises(NotImplementedError):
        x = torch.zeros((4, 2, 32), device="cuda")
        v = torch.zeros((7, 2, 32), device="cuda")
        a = torch.zeros((7, 2, 2), dtype=torch.complex128, device="cuda")
        _lib._check(lib.pf_estep_stereo(x.data_ptr(), v.data_ptr(), a.data_ptr(),
                                        _lib._iarr(range(7)), 7, 7, None, 2, 32, 32, None, None,
                                        None, None, None, 0, 0, 0, None), lib)
