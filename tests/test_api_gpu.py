"""The drop-in API on the CUDA kernels against the reference's golden vectors. Needs a B200."""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import pyfasst_b200.audioModel as am
from tests.test_api_cpu import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

CASES = [("fasst_inst_r1", "mix_inst.wav", False, 1, 3),
         ("fasst_inst_r2", "mix_inst.wav", False, 2, 3),
         ("fasst_conv_r1", "mix_conv.wav", True, 1, 3),
         ("fasst_conv_r2", "mix_conv.wav", True, 2, 2)]


def build_model(wav, conv, rank, nbcomps, dtype, iters=6, **kw):
    np.random.seed(0)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    model = cls(audio=os.path.join(GOLDEN, wav), nbComps=nbcomps, nbNMFComps=4,
                spatial_rank=rank, wlen=256, hopsize=64, iter_num=iters, verbose=0,
                ann_PSD_lim=[None, None], compute_dtype=dtype, **kw)
    if conv:
        model.makeItConvolutive()
    return model


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 2e-6)])
def test_compute_suff_stat_more_than_six_subsources(dtype, tol):
    """4 sources at rank 2 = 8 sub-sources (the headline model): several kernel passes; against the
    oracle (the reference algorithm takes any number of sub-sources)."""
    from oracle import fasst_oracle as fo
    model = build_model("mix_inst.wav", False, 2, 4, dtype)
    ref = fo.OracleFASST(os.path.join(GOLDEN, "mix_inst.wav"), nbComps=4, nbNMFComps=4,
                         spatial_rank=2, wlen=256, hopsize=64, iter_num=1)
    ref.spat_comps, ref.spec_comps = model.spat_comps, model.spec_comps
    ref.noise["PSD"] = model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    powers, mix, ranks = model.retrieve_subsrc_params()
    assert powers.shape[0] == 8
    _, hRxs, hRss, hWs, ll = model.compute_suff_stat(powers, mix)
    _, rRxs, rRss, rWs, rll = ref.compute_suff_stat(powers, mix)
    # (float32 planes: the statistics are float64 sums, the inputs are rounded to float32)
    assert rel_err(hRxs, rRxs) < tol and rel_err(hRss, rRss) < tol
    assert rel_err(hWs, rWs) < max(tol, 1e-6 if dtype == "float32" else 0)
    assert_allclose(ll, np.real(rll), rtol=1e-10 if dtype == "float64" else 1e-6)


def sdr_db(ref, est):
    ref, est = ref.astype(np.float64), est.astype(np.float64)
    err = ((ref - est) ** 2).sum()
    return float('inf') if err == 0 else 10 * np.log10((ref ** 2).sum() / err)


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_model_f64_matches_reference(name, wav, conv, rank, nbcomps, tmp_path):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = build_model(wav, conv, rank, nbcomps, "float64")
    if "Cx" in g:
        assert_allclose(model.Cx, g["Cx"], atol=1e-11)
    assert_allclose(model.noise["ann_PSD_lim"][0], g["ann0"], rtol=1e-11)
    for j in range(nbcomps):
        assert rel_err(model.spat_comps[j]["params"], g["init_A%d" % j]) < 1e-12
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["init_%s%d" % (nm, j)]) < 1e-12
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    powers, mix, ranks = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(powers, mix)
    assert_allclose(hRxs, g["e0_hat_Rxs"], rtol=1e-8, atol=1e-13)
    assert_allclose(hRss, g["e0_hat_Rss"], rtol=1e-8, atol=1e-13)
    assert_allclose(hWs, g["e0_hat_Ws"], rtol=1e-7, atol=1e-300)
    assert_allclose(ll, g["e0_loglik"], rtol=1e-10)
    lls = model.estim_param_a_post_model()
    assert_allclose(lls, g["logliks"], rtol=1e-8)
    for j in range(nbcomps):
        assert rel_err(model.spat_comps[j]["params"], g["final_A%d" % j]) < 1e-6
        for nm in ("FB", "FW", "TW"):
            assert rel_err(model.spec_comps[j]["factor"][0][nm], g["final_%s%d" % (nm, j)]) < 1e-6
    model.separate_spat_comps(dir_results=str(tmp_path))
    import scipy.io.wavfile as wavfile
    for n, f in enumerate(model.files["spat_comp"]):
        fs, y = wavfile.read(f)
        ref = g["sep%d" % n]
        assert y.shape == ref.shape and y.dtype == ref.dtype
        diff = np.abs(y.astype(int) - ref.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


@pytest.mark.parametrize("name,wav,conv,rank,nbcomps", CASES)
def test_model_f32_sdr_parity(name, wav, conv, rank, nbcomps):
    """float32 device path: log-likelihoods within 1e-5, separated signals within 0.01 dB
    SDR of the reference's (north_star) -- measured as the SDR of our output against the
    reference's output being > 50 dB, which bounds any SDR difference well below 0.01 dB."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = build_model(wav, conv, rank, nbcomps, "float32")
    lls = model.estim_param_a_post_model()
    assert_allclose(lls, g["logliks"], rtol=1e-5)
    pcm = model.separate_comps_pcm({j: [j] for j in range(nbcomps)})
    for n in range(nbcomps):
        ref = g["sep%d" % n]
        if ref.dtype != np.int16:
            continue  # quiet component stored as int8 by the reference's writer
        assert sdr_db(ref, pcm[n]) > 50.0


def test_cuda_graph_api():
    a = build_model("mix_inst.wav", False, 2, 3, "float32").estim_param_a_post_model()
    b = build_model("mix_inst.wav", False, 2, 3, "float32",
                    use_cuda_graph=True).estim_param_a_post_model()
    assert_allclose(b, a, rtol=1e-6)


def test_stft_class_roundtrip():
    from pyfasst_b200.tftransforms.stft import STFT
    g = np.load(os.path.join(GOLDEN, "stft.npz"))
    tf = STFT(linFTLen=256, atomHopFactor=0.25, fs=8000)
    tf.computeTransform(g["x"])
    assert_allclose(tf.transfo, g["X"], atol=1e-11)
    assert_allclose(tf.freq_stamps, g["freqs"])
    assert_allclose(tf.time_stamps, g["times"])
    assert_allclose(tf.invertTransform(), g["y"], atol=1e-12)
    tf2 = STFT(linFTLen=2048, atomHopFactor=0.25, fs=44100)
    tf2.computeTransform(g["x2"])
    assert_allclose(tf2.transfo, g["X2"], atol=1e-10)
    assert_allclose(tf2.invertTransform(), g["y2"], atol=1e-12)
