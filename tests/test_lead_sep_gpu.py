"""GPU twin of tests/test_lead_sep_cpu.py: the SIMM front / back end on the CUDA kernels against
the golden vectors made by executing the reference, plus the mask / power kernels against their
NumPy specification."""
import numpy as np
import pytest
import torch

from tests import test_lead_sep_cpu as cpu
from tests.fake_simm_kernels import FakeSimmKernels

pytestmark = pytest.mark.gpu


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


def test_stft_istft_match_reference():
    cpu.check_stft_istft(ck())


def test_write_separated_signals_matches_reference(tmp_path):
    cpu.check_separation(tmp_path, ck())


def test_estim_stereo_simm_params(tmp_path):
    cpu.check_estimation(tmp_path, ck())


@pytest.mark.parametrize("nch", [1, 2])
def test_power_and_mask_kernels(nch):
    rng = np.random.default_rng(nch)
    F, N = 67, 205
    ldn, ldx = (N + 3) // 4 * 4, (N + 31) // 32 * 32
    X = np.zeros((2 * nch, F, ldx), np.float32)
    X[:, :, :N] = rng.standard_normal((2 * nch, F, N))

    def plane(w):
        a = np.zeros((F, w * ldn), np.float32)
        for c in range(w):
            a[:, c * ldn:c * ldn + N] = rng.random((F, N)) ** 2 + 1e-3
        return a
    arrs = dict(X=X, SM=plane(nch), SF0=plane(1), SPHI=plane(1),
                a2=np.array([0.3, 0.5], np.float32), SX=np.full((F, nch * ldn), 9, np.float32),
                Y=np.zeros((4 * nch, F, ldx), np.float32))
    c = {k: torch.tensor(v) for k, v in arrs.items()}
    g = {k: v.clone().cuda() for k, v in c.items()}
    for d, k in ((c, FakeSimmKernels()), (g, ck())):
        k.simm_power(d["X"], d["SX"], nch, F, N, ldn)
        k.simm_masks(d["SM"], d["SF0"], d["SPHI"], d["a2"], d["X"], d["Y"], 1e-9, nch, F, N, ldn)
    torch.cuda.synchronize()
    for key in ("SX", "Y"):
        a, b = g[key].cpu().numpy(), c[key].numpy()
        assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max(), key
