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


@pytest.mark.parametrize("transA", [False, True])
@pytest.mark.parametrize("transB", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (1025, 700, 480), (480, 1000, 1028),
                                   (4, 515, 1025), (300, 40, 2048), (130, 5, 64),
                                   (40, 1300, 1025)])
def test_gemm_tf32x3(transA, transB, M, N, K):
    """pf_gemm_tf32x3 against float64 NumPy for all operand layouts, ragged sizes included."""
    from pyfasst_b200._lib import CudaKernels
    ck = CudaKernels()
    if (not transA or transB) and K % 4:
        K += 4 - K % 4  # operands contiguous along K need a zero-padded K % 4 == 0
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)

    def padded(a):  # leading dimension rounded up to 4, zero padding
        ld = (a.shape[1] + 3) // 4 * 4
        out = np.zeros((a.shape[0], ld), dtype=np.float32)
        out[:, :a.shape[1]] = a
        return torch.tensor(out).cuda()

    Ad = padded(A.T if transA else A)
    Bd = padded(B.T if transB else B)
    Cd = torch.full((M, (N + 3) // 4 * 4), -3.0, dtype=torch.float32, device="cuda")
    ck.gemm(Ad, Bd, Cd, M, N, K, transA, transB)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = Cd.cpu().numpy()
    err = np.abs(got[:, :N] - ref).max() / np.abs(ref).max()
    # 3xTF32 products are float32-class (~2^-21); the tensor core's float32 accumulator adds a
    # rounding error that grows with the number of K steps (measured: 4e-6 at K=480, 9e-6 at
    # K=2048 on sign-mixed data) -- still two orders of magnitude better than plain tf32
    assert err < 2e-5, err
    assert (got[:, N:] == -3.0).all(), "columns beyond N must not be written"
