"""BASELINE configs[0] at its real size on the CPU: the host logic of the drop-in classes and the
float64 NumPy kernel specification (tests/fake_kernels.py), against the golden vectors made by
executing the reference on data/tamy.wav (oracle/make_golden_fullsize.py).  F = 1025,
N = 1122, 3 sources, rank 1, K = 4, STFT 2048 / 512.  The 50-iteration trajectory on the
CUDA kernels is tests/test_tamy_gpu.py; here: the first E-step, 3 iterations of the
50-iteration schedule (the annealing depends on iter_num, so the engine is driven directly)
and the oracle on the same."""
import os

import numpy as np
from numpy.testing import assert_allclose

import pyfasst_b200.audioModel as am
from tests.fake_kernels import FakeKernels
from tests.test_api_cpu import GOLDEN, rel_err


def build(iters):
    np.random.seed(0)
    return am.MultiChanNMFInst_FASST(audio=os.path.join(GOLDEN, "tamy.wav"), nbComps=3,
                                     spatial_rank=1, iter_num=iters, verbose=0,
                                     compute_dtype="float64", kernels=FakeKernels())


def test_tamy_first_iteration_matches_reference():
    g = np.load(os.path.join(GOLDEN, "tamy_inst_r1.npz"))
    model = build(1)
    assert (model.nbFreqsSigRepr, model.nbFramesSigRepr) == (1025, 1122)
    assert_allclose(model.noise["ann_PSD_lim"][0], g["ann0"], rtol=1e-10)
    assert_allclose(model.noise["ann_PSD_lim"][1], g["ann1"], rtol=1e-10)
    for j in range(3):
        assert rel_err(model.spat_comps[j]["params"], g["init_A%d" % j]) < 1e-12
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["init_%s%d" % (nm, j)]) < 1e-12
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    powers, mix, ranks = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(powers, mix)
    assert_allclose(hRss, g["e0_hat_Rss"], rtol=1e-7, atol=1e-12 * np.abs(g["e0_hat_Rss"]).max())
    assert_allclose(hRxs, g["e0_hat_Rxs"], rtol=1e-7, atol=1e-12 * np.abs(g["e0_hat_Rxs"]).max())
    assert_allclose(hWs[:, list(g["e0_rows"]), :], g["e0_hat_Ws_rows"], rtol=1e-6, atol=1e-300)
    assert_allclose(ll, g["e0_loglik"], rtol=1e-11)
    lls = model.estim_param_a_post_model()
    assert_allclose(lls, g["ll_it1"], rtol=1e-10)
    for j in range(3):
        assert rel_err(model.spat_comps[j]["params"], g["it1_A%d" % j]) < 1e-8
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["it1_%s%d" % (nm, j)]) < 1e-8


def test_tamy_first_iterations_of_the_50_schedule():
    """Iterations 1-3 of the 50-iteration run (annealed noise PSD of a 50-step schedule)."""
    g = np.load(os.path.join(GOLDEN, "tamy_inst_r1.npz"))
    model = build(50)
    eng = model._engine()
    torch = eng.torch
    logliks = torch.ones([50], dtype=torch.float64)
    eng.iter_dev.zero_()
    eng.flags.zero_()
    for _ in range(3):
        eng.totals.zero_()
        eng.gem_iteration(50, logliks)
    assert_allclose(logliks[:3].numpy(), g["logliks"][:3], rtol=1e-9)


def test_tamy_fifty_iterations_and_separation():
    """All of BASELINE configs[0] on the float64 kernel specification: logliks[50], the final
    parameters and the separated signals against the executed reference."""
    g = np.load(os.path.join(GOLDEN, "tamy_inst_r1.npz"))
    model = build(50)
    lls = model.estim_param_a_post_model()
    assert_allclose(lls, g["logliks"], rtol=1e-8)
    for j in range(3):
        assert rel_err(model.spat_comps[j]["params"], g["final_A%d" % j]) < 1e-6
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["final_%s%d" % (nm, j)]) < 1e-6
    assert_allclose(model.noise["PSD"], g["noise_PSD_final"], rtol=1e-12)
    pcm = model.separate_comps_pcm({j: [j] for j in range(3)})
    for n in range(3):
        ref = g["sep%d" % n]
        diff = np.abs(pcm[n].astype(int) - ref.astype(int))
        assert pcm[n].shape == ref.shape and diff.max() <= 1 and (diff > 0).mean() < 1e-3
