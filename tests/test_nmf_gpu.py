"""GPU twin of tests/test_nmf_cpu.py, plus the NMF kernels against their NumPy specification."""
import numpy as np
import pytest
import torch

from tests import test_nmf_cpu as cpu
from tests.fake_simm_kernels import FakeSimmKernels

pytestmark = pytest.mark.gpu


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


def test_functions_match_reference():
    cpu.check_functions(ck())


@pytest.mark.parametrize("same", [False, True])
def test_model_initialisation(same):
    cpu.check_model_init(ck(), same)


def test_nmf_kernels():
    rng = np.random.default_rng(4)
    F, N, K = 67, 203, 5
    ldn, ldk, ldx = (N + 3) // 4 * 4, (K + 3) // 4 * 4, (N + 31) // 32 * 32

    def plane(w=1):
        a = np.zeros((F, w * ldn), np.float32)
        for c in range(w):
            a[:, c * ldn:c * ldn + N] = rng.random((F, N)) ** 2 + 1e-3
        return a
    hat = plane()
    hat[3, 5] = 1e-7  # below sqrt(eps): the clamp of hat^2
    X = np.zeros((4, F, ldx), np.float32)
    X[:, :, :N] = rng.standard_normal((4, F, N))
    W = np.zeros((F, ldk), np.float32)
    W[:, :K] = rng.random((F, K)) + 0.1
    W[:, 2] = 0  # a column that sums to zero counts as one (nmf.py:47-48)
    arrs = dict(hat=hat, SX=plane(), out=np.full((F, 2 * ldn), 7, np.float32),
                H=np.pad(rng.random((K, N)).astype(np.float32), ((0, ldk - K), (0, ldn - N))),
                C=(rng.random((ldk, 2 * ldn)) + 0.1).astype(np.float32), W=W,
                D=(rng.random((2, F, ldk)) + 0.1).astype(np.float32), s=np.zeros(ldk, np.float32),
                X=X, mono=np.full((F, ldn), 3, np.float32))
    c = {k: torch.tensor(v) for k, v in arrs.items()}
    g = {k: v.clone().cuda() for k, v in c.items()}
    for d, k in ((c, FakeSimmKernels()), (g, ck())):
        k.nmf_is_terms(d["hat"], d["SX"], d["out"], 1e-10, F, N, ldn)
        k.nmf_update_rows(d["H"], d["C"], ldn, 1e-10, K, N)
        k.nmf_w_update(d["W"], K, d["D"], 1e-10, F, d["s"])
        k.mono_power(d["X"], d["mono"], F, N, ldn)
    torch.cuda.synchronize()
    for key in ("out", "H", "W", "s", "mono"):
        a, b = g[key].cpu().numpy(), c[key].numpy()
        assert np.isfinite(a).all(), key
        assert np.abs(a - b).max() <= 5e-6 * np.abs(b).max(), key
    assert g["s"].cpu().numpy()[2] == 1.0
