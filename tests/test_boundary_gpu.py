"""tests/test_boundary_cpu.py's checks on the CUDA kernels (needs a B200): the drop-in surface of
SURVEY.md 8(b) and row a9 against golden vectors made by executing the reference."""
import pytest

from tests import test_boundary_cpu as b

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 1e-5)])
def test_update_mix_matrix(ck, dtype, tol):
    b.check_update_mix_matrix(ck, dtype, tol)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-5)])
def test_update_spectral_components(ck, dtype, tol):
    b.check_update_spectral_components(ck, dtype, tol)


def test_filter_stft(ck):
    b.check_filter_stft(ck, 1e-10)


def test_time_blobs(ck):
    # (general factor structures run on float32 planes: tf32x3 tensor-core contractions)
    b.check_time_blobs(ck, "float32", 1e-5, 2e-5, 2e-3)


def test_lambda_corr(ck):
    b.check_lambda_corr(ck, "float32", 2e-5, 2e-3)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 1e-5)])
def test_tw_redraw(ck, dtype, tol):
    b.check_redraw(ck, dtype, tol)


@pytest.mark.parametrize("dtype,lsb", [("float64", 1), ("float32", 2)])
def test_subset_separation(ck, dtype, lsb):
    b.check_subset_separation(ck, dtype, lsb)
