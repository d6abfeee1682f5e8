"""tcgen05 building blocks (csrc/tc.cuh): descriptor and operand-layout conventions checked
with a one-CTA GEMM against NumPy for every K-major / MN-major combination."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("N,K", [(32, 32), (64, 96), (256, 64)])
@pytest.mark.parametrize("split3", [0, 1])
def test_tc_selftest(a_mn, b_mn, N, K, split3):
    from pyfasst_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(N + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    Ad = torch.tensor(np.ascontiguousarray(A.T if a_mn else A)).cuda()
    Bd = torch.tensor(np.ascontiguousarray(B.T if b_mn else B)).cuda()
    D = torch.zeros((128, N), dtype=torch.float32, device="cuda")
    _lib._check(lib.pf_tc_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, a_mn, b_mn,
                                   split3, None), lib)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(D.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert err < (2e-6 if split3 else 2e-3), err
    if not split3:
        assert err > 1e-6  # really the tf32 path
