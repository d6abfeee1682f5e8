"""GPU twin of tests/test_melody_cpu.py: melody tracking against the reference's own code
(golden vectors) and the whole autoMelSepAndWrite pipeline on the CUDA kernels."""
import pytest

from tests import test_melody_cpu as cpu

pytestmark = pytest.mark.gpu


def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


def test_tracking_matches_reference(tmp_path):
    cpu.check_tracking(tmp_path, ck())


def test_pipeline(tmp_path):
    cpu.check_pipeline(tmp_path, ck())
