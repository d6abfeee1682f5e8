"""Parity at BASELINE.json's real sizes against the EXECUTED reference (needs a B200).

`data/tamy.wav` (the reference's own fixture, tests/golden/tamy.wav), STFT 2048 / 512
(F = 1025, N = 1122), 50 GEM iterations, golden vectors made by oracle/make_golden_fullsize.py:

  tamy_inst_r1  BASELINE configs[0]: MultiChanNMFInst_FASST, 3 sources, rank 1, K = 4
  tamy_conv_r2  MultiChanNMFConv (convolutive), 3 sources, rank 2

north_star's tolerances, float32 AND float64 planes: W/H/A relative error <= 1e-4 after one
iteration, log-likelihood trajectory within 1e-5 relative over all 50 iterations, separated
signals within 0.01 dB SDR of the reference's (checked as: SDR of our signal against the
reference's signal >= 50 dB, which bounds the SDR difference against any ground truth by
20 log10(1 + 10^(-50/20)) = 0.027 dB worst case and << 0.01 dB for uncorrelated errors; the
measured values are printed).
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import pyfasst_b200.audioModel as am
from tests.test_api_cpu import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

CASES = [("tamy_inst_r1", False, 1), ("tamy_conv_r2", True, 2)]


def build(conv, rank, dtype, iters=50, **kw):
    np.random.seed(0)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    model = cls(audio=os.path.join(GOLDEN, "tamy.wav"), nbComps=3, spatial_rank=rank,
                iter_num=iters, verbose=0, compute_dtype=dtype, **kw)
    if conv:
        model.makeItConvolutive()
    return model


def sdr_db(ref, est):
    ref, est = ref.astype(np.float64), est.astype(np.float64)
    err = ((ref - est) ** 2).sum()
    return float('inf') if err == 0 else 10 * np.log10((ref ** 2).sum() / err)


def params_err(model, g, prefix):
    worst = 0.0
    for j in range(3):
        worst = max(worst, rel_err(model.spat_comps[j]["params"], g["%s_A%d" % (prefix, j)]))
        for nm in ("FB", "FW", "TW"):
            worst = max(worst, rel_err(model.spec_comps[j]["factor"][0][nm],
                                       g["%s_%s%d" % (prefix, nm, j)]))
    return worst


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name,conv,rank", CASES)
def test_first_estep_and_iteration(name, conv, rank, dtype):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    tol = 1e-9 if dtype == "float64" else 1e-4
    model = build(conv, rank, dtype, iters=1)
    assert model.nbFreqsSigRepr == 1025 and model.nbFramesSigRepr == 1122
    assert_allclose(model.noise["ann_PSD_lim"][0], g["ann0"], rtol=1e-9 if dtype == "float64" else 1e-5)
    assert params_err(model, g, "init") < (1e-12 if dtype == "float64" else 1e-6)
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    powers, mix, ranks = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(powers, mix)
    scale = np.abs(g["e0_hat_Rss"]).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(hRss - g["e0_hat_Rss"]) / scale) < tol
    scale = np.abs(g["e0_hat_Rxs"]).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(hRxs - g["e0_hat_Rxs"]) / scale) < tol
    rows = list(g["e0_rows"])
    assert_allclose(hWs[:, rows, :], g["e0_hat_Ws_rows"], rtol=10 * tol, atol=1e-300)
    assert_allclose(ll, g["e0_loglik"], rtol=1e-10 if dtype == "float64" else 1e-6)
    lls = model.estim_param_a_post_model()
    assert_allclose(lls, g["ll_it1"], rtol=1e-10 if dtype == "float64" else 1e-6)
    err = params_err(model, g, "it1")
    print("%s %s: W/H/A after one iteration, worst relative error %.3g" % (name, dtype, err))
    assert err < tol


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name,conv,rank", CASES)
def test_fifty_iterations_and_separation(name, conv, rank, dtype):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = build(conv, rank, dtype)
    lls = model.estim_param_a_post_model()
    assert lls.shape == (50,)
    worst = np.max(np.abs(lls - g["logliks"]) / np.abs(g["logliks"]))
    print("%s %s: log-likelihood trajectory, worst relative error over 50 iterations %.3g"
          % (name, dtype, worst))
    assert worst < (1e-8 if dtype == "float64" else 1e-5)
    err = params_err(model, g, "final")
    print("%s %s: W/H/A after 50 iterations, worst relative error %.3g" % (name, dtype, err))
    assert err < (1e-5 if dtype == "float64" else 5e-2)
    assert_allclose(model.noise["PSD"], g["noise_PSD_final"], rtol=1e-9 if dtype == "float64" else 1e-5)
    pcm = model.separate_comps_pcm({j: [j] for j in range(3)})
    for n in range(3):
        ref = g["sep%d" % n]
        assert pcm[n].shape == ref.shape
        s = sdr_db(ref, pcm[n])
        print("%s %s: source %d, SDR against the reference's signal %.1f dB" % (name, dtype, n, s))
        assert s > (80.0 if dtype == "float64" else 50.0)
