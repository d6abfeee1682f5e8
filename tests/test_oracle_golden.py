"""The oracle (oracle/fasst_oracle.py) against vectors produced by executing the
reference itself (oracle/make_golden.py) and against the reference's own
known-answer tests.  CPU only."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_almost_equal

from oracle import fasst_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


# ---- reference KATs (pyfasst_tests/pyfasst/tools/test_utils.py:23-63) -------
def test_nextpow2_kat():
    assert fo.nextpow2(2) == 2
    assert fo.nextpow2(2 ** 10 + 1) == 2 ** 11
    assert fo.nextpow2(2 ** 20 + 1) == 2 ** 21


def test_sinebell_kat():
    assert_array_almost_equal(
        fo.sinebell(5),
        np.array([0., 0.58778525, 0.95105652, 0.95105652, 0.58778525]))
    assert_array_almost_equal(
        fo.sinebell(10),
        np.array([0., 0.30901699, 0.58778525, 0.80901699, 0.95105652, 1.,
                  0.95105652, 0.80901699, 0.58778525, 0.30901699]))


def test_hann_kat():
    assert_array_almost_equal(fo.hann(11), np.hanning(11))
    assert_array_almost_equal(fo.hann(22), np.hanning(22))


# ---- reference KAT (pyfasst_tests/pyfasst/tools/test_signalTools.py:28-63) --
SIGMA_X_DIAG = np.array(
    [[0.00977917, 0.01021195, 0.00949931, 0.01081156, 0.00982221,
      0.00927985, 0.01090643, 0.01078789, 0.00941831, 0.01113587],
     [0.00785231, 0.00819886, 0.00762822, 0.00867899, 0.00788678,
      0.00745249, 0.00875495, 0.00866003, 0.00756336, 0.00893867]])
SIGMA_X_OFF = np.array(
    [0.00865282, 0.00904009, 0.00840240, 0.00957665, 0.00869134,
     0.00820601, 0.00966154, 0.00955547, 0.00832991, 0.00986685]) + 0j
INV_DIAG_REF = np.array(
    [[4094.58407492, 4093.23448666, 4095.52259353, 4091.54400714,
      4094.44450083, 4096.2983896, 4091.29365679, 4091.60715192,
      4095.80470268, 4090.70587309],
     [5099.34077392, 5098.26012193, 5100.09227127, 5096.90650902,
      5099.22901315, 5100.71347227, 5096.7060467, 5096.95707075,
      5100.31816375, 5096.2353923]])


def test_inv_herm_mat_2d_kat():
    idg, iof, det = fo.inv_herm_mat_2d(SIGMA_X_DIAG, SIGMA_X_OFF)
    assert_array_almost_equal(
        idg[0] * SIGMA_X_DIAG[0] + SIGMA_X_OFF * np.conj(iof),
        np.ones_like(iof))
    assert_array_almost_equal(
        idg[0] * np.conj(SIGMA_X_OFF) + SIGMA_X_DIAG[1] * np.conj(iof),
        np.zeros_like(iof))
    # listed inverses: inputs are printed to 8 digits and ill-conditioned
    assert_allclose(idg, INV_DIAG_REF, rtol=1e-4)


def test_inv_herm_mat_2d_golden():
    g = load("inv2d")
    idg, iof, det = fo.inv_herm_mat_2d(g["d"], g["o"])
    assert_allclose(idg, g["inv_d"], rtol=1e-14)
    assert_allclose(iof, g["inv_o"], rtol=1e-14)
    assert_allclose(det, g["det"], rtol=1e-14)


# ---- STFT / iSTFT -----------------------------------------------------------
def test_stft_golden():
    g = load("stft")
    tf = fo.STFT(linFTLen=256, atomHopFactor=0.25, fs=8000)
    tf.computeTransform(g["x"])
    assert tf.transfo.shape == g["X"].shape
    assert_allclose(tf.transfo, g["X"], atol=1e-12)
    assert_allclose(tf.freq_stamps, g["freqs"])
    assert_allclose(tf.time_stamps, g["times"])
    assert_allclose(tf.invertTransform(), g["y"], atol=1e-13)
    tf2 = fo.STFT(linFTLen=2048, atomHopFactor=0.25, fs=44100)
    tf2.computeTransform(g["x2"])
    assert_allclose(tf2.transfo, g["X2"], atol=1e-11)
    assert_allclose(tf2.invertTransform(), g["y2"], atol=1e-13)


# ---- FASST GEM --------------------------------------------------------------
CASES = [("fasst_inst_r1", "mix_inst.wav", False, 1, 3),
         ("fasst_inst_r2", "mix_inst.wav", False, 2, 3),
         ("fasst_conv_r1", "mix_conv.wav", True, 1, 3),
         ("fasst_conv_r2", "mix_conv.wav", True, 2, 2)]


def build(wav, conv, rank, nbcomps, iters=6):
    np.random.seed(0)
    m = fo.OracleFASST(os.path.join(GOLDEN, wav), nbComps=nbcomps,
                       nbNMFComps=4, spatial_rank=rank, wlen=256, hopsize=64,
                       iter_num=iters)
    if conv:
        m.makeItConvolutive()
    return m


def check_state(m, g, prefix, rtol):
    for j, sc in m.spat_comps.items():
        assert_allclose(sc["params"], g["%s_A%d" % (prefix, j)], rtol=rtol,
                        atol=1e-13)
    for k, sp in m.spec_comps.items():
        fac = sp["factor"][0]
        for nm in ("FB", "FW", "TW"):
            assert_allclose(fac[nm], g["%s_%s%d" % (prefix, nm, k)], rtol=rtol,
                            atol=1e-300)


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_fasst_golden(name, wav, conv, rank, nbcomps):
    g = load(name)
    m = build(wav, conv, rank, nbcomps)
    if "Cx" in g:
        assert_allclose(m.Cx, g["Cx"], atol=1e-12)
    assert_allclose(m.noise["ann_PSD_lim"][0], g["ann0"], rtol=1e-12)
    assert_allclose(m.noise["ann_PSD_lim"][1], g["ann1"], rtol=1e-12)
    check_state(m, g, "init", 1e-13)
    # E-step on the initial parameters
    m.noise["PSD"] = m.noise["ann_PSD_lim"][0]
    powers, mix, ranks = m.retrieve_subsrc_params()
    _, hRxs, hRss, hWs, ll = m.compute_suff_stat(powers, mix)
    assert_allclose(hRxs, g["e0_hat_Rxs"], rtol=1e-9, atol=1e-14)
    assert_allclose(hRss, g["e0_hat_Rss"], rtol=1e-9, atol=1e-14)
    assert_allclose(hWs, g["e0_hat_Ws"], rtol=1e-9, atol=1e-300)
    assert_allclose(np.real(ll), g["e0_loglik"], rtol=1e-12)
    # generalised-I E-step == stereo closed form at I = 2
    gRxs, gRss, gWs, gll = fo.estep_general(m.X, powers, mix, m.noise["PSD"])
    assert_allclose(gRxs, hRxs, rtol=1e-8, atol=1e-13)
    assert_allclose(gRss, hRss, rtol=1e-8, atol=1e-13)
    assert_allclose(gWs, hWs, rtol=1e-8, atol=1e-300)
    assert_allclose(gll, np.real(ll), rtol=1e-11)
    # one iteration
    m.iter_num = 1
    ll1 = m.estim_param_a_post_model()
    assert_allclose(ll1, g["ll_it1"], rtol=1e-12)
    check_state(m, g, "it1", 1e-9)
    # full trajectory from the same init
    m = build(wav, conv, rank, nbcomps)
    lls = m.estim_param_a_post_model()
    assert_allclose(lls, g["logliks"], rtol=1e-9)
    check_state(m, g, "final", 1e-7)
    assert_allclose(m.noise["PSD"], g["noise_PSD_final"], rtol=1e-12)
    # separation: int16 truncation may flip an LSB on rounding ties only
    pcm = m.separate_spat_comps(dir_results="/tmp")
    for n, y in enumerate(pcm):
        ref = g["sep%d" % n]
        assert y.shape == ref.shape
        diff = np.abs(y.astype(int) - ref.astype(int))
        assert diff.max() <= 1
        assert (diff > 0).mean() < 1e-3
