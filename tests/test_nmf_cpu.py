"""IS-NMF initialisers (pyfasst_b200/tools/nmf.py, FASST.initialize_all_spec_comps_with_NMF*):
the oracle against golden vectors produced by the reference's tools/nmf.py, and the host
orchestration on the NumPy specification of the kernels against both.  CPU only; GPU twin:
tests/test_nmf_gpu.py."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import nmf_oracle as no
from pyfasst_b200.tools import nmf as pnmf
from tests.fake_kernels import FakeKernels
from tests.fake_simm_kernels import FakeSimmKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 2e-4  # float32 factors / 3xTF32 products after 5 iterations, against float64


class AllFakeKernels(FakeSimmKernels, FakeKernels):
    pass


def load():
    return np.load(os.path.join(GOLDEN, "nmf.npz"))


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_oracle_matches_reference():
    g = load()
    W, H = no.nmf_decomposition(g["SX"], g["W0"], g["H0"], niter=5)
    assert_allclose(W, g["W1"], rtol=1e-10)
    assert_allclose(H, g["H1"], rtol=1e-10)
    W, H = no.nmf_decomp_init(g["SX"], g["Wi"], g["Hi"], niter=5)
    assert_allclose(W, g["W2"], rtol=1e-10)
    assert_allclose(H, g["H2"], rtol=1e-10)
    W, H = no.nmf_decomp_init(g["SX"], g["Wi"], g["Hi"], niter=5, updateW=False)
    assert_allclose(W, g["W3"], rtol=1e-10)
    assert_allclose(H, g["H3"], rtol=1e-10)


def check_functions(kernels):
    g = load()
    K = g["W0"].shape[1]
    np.random.seed(12)  # the seed make_golden.py used: same random initial W, H
    W, H = pnmf.NMF_decomposition(g["SX"], nbComps=K, niter=5, kernels=kernels)
    assert rel_err(W, g["W1"]) < RTOL and rel_err(H, g["H1"]) < RTOL
    W, H = pnmf.NMF_decomp_init(g["SX"], nbComps=K, niter=5, Winit=g["Wi"], Hinit=g["Hi"],
                                kernels=kernels)
    assert rel_err(W, g["W2"]) < RTOL and rel_err(H, g["H2"]) < RTOL
    W, H = pnmf.NMF_decomp_init(g["SX"], nbComps=K, niter=5, Winit=g["Wi"], Hinit=g["Hi"].T,
                                updateW=False, kernels=kernels)
    assert rel_err(W, g["W3"]) < 1e-7 and rel_err(H, g["H3"]) < RTOL
    with pytest.raises(AttributeError):
        pnmf.NMF_decomp_init(g["SX"], nbComps=K, niter=1, Hinit=g["Hi"][:, :-1], kernels=kernels)


def check_model_init(kernels, same):
    """FASST.initialize_all_spec_comps_with_NMF against the oracle run on the model's own mono
    power spectrum and initial factors."""
    import pyfasst_b200.audioModel as am
    wav = os.path.join(GOLDEN, "mix_inst.wav")
    np.random.seed(2)
    m = am.MultiChanNMFInst_FASST(audio=wav, nbComps=3, nbNMFComps=4, spatial_rank=1, wlen=256,
                                  hopsize=64, iter_num=2, compute_dtype="float64",
                                  ann_PSD_lim=[None, None], kernels=kernels)
    X = m._X[:, :, :m.nbFramesSigRepr].cpu().numpy().astype(np.float64)
    SX = (X ** 2).sum(axis=0) / 2.0
    FB0 = np.hstack([m.spec_comps[j]["factor"][0]["FB"] for j in range(3)])
    TW0 = np.vstack([m.spec_comps[j]["factor"][0]["TW"] for j in range(3)])
    ref = am.MultiChanNMFInst_FASST.__new__(am.MultiChanNMFInst_FASST)  # an untouched copy
    ref.__dict__.update({k: v for k, v in m.__dict__.items()})
    import copy
    ref.spat_comps, ref.spec_comps = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    if same:
        np.random.seed(7)
        m.initialize_all_spec_comps_with_NMF(sameInitAll=True, niter=4)
        np.random.seed(7)
        W0 = np.random.randn(SX.shape[0], 4) ** 2
        H0 = np.random.randn(4, SX.shape[1]) ** 2
        W, H = no.nmf_decomposition(SX, W0, H0, niter=4)
        order = np.argsort(H.sum(axis=1))[::-1]
        for j in range(3):
            ref.spec_comps[j]["factor"][0]["FB"][:] = W[:, order]
            ref.spec_comps[j]["factor"][0]["TW"][:] = H[order]
    else:
        m.initialize_all_spec_comps_with_NMF(niter=4)
        W, H = no.nmf_decomp_init(SX, FB0, TW0, niter=4)
        for j in range(3):
            ref.spec_comps[j]["factor"][0]["FB"] = np.maximum(W[:, 4 * j:4 * j + 4], 1e-10)
            ref.spec_comps[j]["factor"][0]["TW"] = np.maximum(H[4 * j:4 * j + 4], 1e-10)
    ref.renormalize_parameters()
    for j in range(3):
        for key in ("FB", "TW"):
            a, b = m.spec_comps[j]["factor"][0][key], ref.spec_comps[j]["factor"][0][key]
            assert rel_err(a, b) < RTOL, (j, key)
        assert_allclose(m.spat_comps[j]["params"], ref.spat_comps[j]["params"], rtol=1e-6)


def test_functions_match_reference():
    check_functions(AllFakeKernels())


@pytest.mark.parametrize("same", [False, True])
def test_model_initialisation(same):
    check_model_init(AllFakeKernels(), same)
