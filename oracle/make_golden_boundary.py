#!/usr/bin/env python
"""Test infrastructure only -- NOT part of the product path.

Golden vectors for the rest of the drop-in surface of SURVEY.md section 8(b) and row a9, made by
EXECUTING THE REFERENCE (oracle/_py2shim.py) on the small seeded mixtures of make_golden.py:

  umm_*    FASST.update_mix_matrix called by hand (audioModel.py:766-889): instantaneous rank 2
           with one fixed component (the "other sources" branch), convolutive rank 2
  usc_*    FASST.update_spectral_components(hat_W) called by hand (:1469-1727)
  fs_*     tftransforms.stft.filter_stft (stft.py:133-227), W of 3 and 4 dimensions
  tb_*     a model whose spectral components have time blobs TB (:1931-1978, :2026-2030)
  lc_*     lambdaCorr > 0: the correlation penalty (:1484-1703)
  rd_*     renormalize_parameters when sum(TW) < eps: the random re-draw (:2023-2025)
  sub_*    separate_comps with spectral components left out (:1088-1236, :1327-1390)

    python oracle/make_golden_boundary.py
"""
import os
import sys
import warnings

import numpy as np
import scipy.io.wavfile as wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import _py2shim  # noqa: E402
from make_golden import snapshot  # noqa: E402

warnings.simplefilter("ignore")
KW = dict(wlen=256, hopsize=64, verbose=0, ann_PSD_lim=[None, None])


def model_inst(am, rank=2, iters=3, **kw):
    np.random.seed(0)
    args = dict(KW)
    args.update(kw)
    return am.MultiChanNMFInst_FASST(audio=os.path.join(GOLD, "mix_inst.wav"), nbComps=3,
                                     nbNMFComps=4, spatial_rank=rank, iter_num=iters, **args)


def model_conv(am, rank=2, iters=3, **kw):
    np.random.seed(0)
    args = dict(KW)
    args.update(kw)
    m = am.MultiChanNMFConv(audio=os.path.join(GOLD, "mix_conv.wav"), nbComps=2, nbNMFComps=4,
                            spatial_rank=rank, iter_num=iters, **args)
    m.makeItConvolutive()
    return m


def stats(model):
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    scp, mm, rpi = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(scp, mm)
    return hRxs, hRss, hWs, mm, rpi


def run_update_mix(am, out):
    for tag, model in (("inst", model_inst(am)), ("conv", model_conv(am))):
        if tag == "inst":
            model.spat_comps[1]["frdm_prior"] = "fixed"  # exercises upd_inst_other_ind (:811-818)
        hRxs, hRss, hWs, mm, rpi = stats(model)
        out["umm_%s_hat_Rxs" % tag], out["umm_%s_hat_Rss" % tag] = np.array(hRxs), np.array(hRss)
        out["umm_%s_mix0" % tag] = np.array(mm)
        model.update_mix_matrix(hRxs, hRss, mm, rpi)
        out["umm_%s_mix1" % tag] = np.array(mm)
        out["umm_%s_hat_Rxs_after" % tag] = np.array(hRxs)  # (modified in place for 'inst', :810)
        for j, sc in model.spat_comps.items():
            out["umm_%s_A%d" % (tag, j)] = np.array(sc["params"])


def run_update_spec(am, out):
    model = model_inst(am, nmfUpdateCoeff=0.7)
    hRxs, hRss, hWs, mm, rpi = stats(model)
    hat_W = np.zeros([len(model.spat_comps), model.nbFreqsSigRepr, model.nbFramesSigRepr])
    for j in range(len(model.spat_comps)):
        hat_W[j] = np.mean(hWs[rpi[j]], axis=0)   # (:408-414)
    out["usc_hat_W"] = hat_W
    model.spec_comps[1]["factor"][0]["FB_frdm_prior"] = "fixed"
    model.spec_comps[2]["factor"][0]["TW_frdm_prior"] = "fixed"
    model.update_spectral_components(hat_W)
    snapshot(model, "usc", out)


def run_filter_stft(ref, out):
    st = ref["stft"]
    rng = np.random.default_rng(31)
    T, nc, wlen, hop = 2000, 2, 256, 64
    data = rng.standard_normal((T, nc))
    F = wlen // 2 + 1
    N = int(np.ceil(T / float(hop)))
    W3 = rng.standard_normal((nc, nc, F)) + 1j * rng.standard_normal((nc, nc, F))
    W4 = rng.standard_normal((nc, nc, F, N)) + 1j * rng.standard_normal((nc, nc, F, N))
    win = np.hanning(wlen)
    out["fs_data"], out["fs_W3"], out["fs_W4"] = data, W3, W4
    out["fs_y3"] = st.filter_stft(data, W3, analysisWindow=None, synthWindow=win, hopsize=hop,
                                  nfft=wlen, fs=8000)
    out["fs_y4"] = st.filter_stft(data, W4, analysisWindow=None, synthWindow=win, hopsize=hop,
                                  nfft=wlen, fs=8000)
    # analysis window given (and of the synthesis window's length), longer transform than window
    out["fs_y3b"] = st.filter_stft(data, np.concatenate([W3, W3[:, :, :-1][:, :, ::-1]], axis=2)
                                   [:, :, :257], analysisWindow=ref["utils"].sinebell(wlen),
                                   synthWindow=win, hopsize=hop, nfft=512, fs=8000)


def add_time_blobs(model, L=12, seed=4):
    """TW [K, L] and TB [L, N]: smooth non-negative time patterns (audioModel.py:1525-1528)."""
    rng = np.random.RandomState(seed)
    N = model.nbFramesSigRepr
    for sp in model.spec_comps.values():
        fac = sp["factor"][0]
        K = fac["TW"].shape[0]
        fac["TW"] = 0.75 * np.abs(rng.randn(K, L)) + 0.25
        centres = np.linspace(0, N - 1, L)
        fac["TB"] = np.exp(-0.5 * ((np.arange(N)[None, :] - centres[:, None]) / (N / float(L))) ** 2) \
            + 0.05 * np.abs(rng.randn(L, N))
        fac["TB_frdm_prior"] = "free"


def run_time_blobs(am, out):
    model = model_inst(am, rank=1, iters=4)
    add_time_blobs(model)
    model.spec_comps[2]["factor"][0]["TB_frdm_prior"] = "fixed"
    for k, sp in model.spec_comps.items():
        out["tb_init_TW%d" % k] = np.array(sp["factor"][0]["TW"])
        out["tb_init_TB%d" % k] = np.array(sp["factor"][0]["TB"])
    model.renormalize_parameters()
    for k, sp in model.spec_comps.items():
        out["tb_renorm_TW%d" % k] = np.array(sp["factor"][0]["TW"])
        out["tb_renorm_TB%d" % k] = np.array(sp["factor"][0]["TB"])
        out["tb_renorm_FB%d" % k] = np.array(sp["factor"][0]["FB"])
    out["tb_V0"] = model.comp_spat_comp_power(0)
    out["tb_logliks"] = np.real(model.estim_param_a_post_model())
    snapshot(model, "tb_final", out)
    for k, sp in model.spec_comps.items():
        out["tb_final_TB%d" % k] = np.array(sp["factor"][0]["TB"])


def run_lambda_corr(am, out):
    for tag, lam in (("lc", 0.1),):
        model = model_inst(am, rank=2, iters=4, lambdaCorr=lam)
        out["%s_logliks" % tag] = np.real(model.estim_param_a_post_model())
        snapshot(model, "%s_final" % tag, out)


def run_redraw(am, out):
    model = model_inst(am, rank=1)
    model.spec_comps[1]["factor"][0]["TW"][:] = 1e-14
    np.random.seed(77)
    model.renormalize_parameters()
    snapshot(model, "rd", out)


def run_subset(am, out):
    model = model_inst(am, rank=2, iters=3)
    model.estim_param_a_post_model()
    snapshot(model, "sub_model", out)
    out["sub_noise_PSD"] = np.array(model.noise["PSD"])
    outdir = "/tmp/pyfasst_golden_out_subset"
    os.makedirs(outdir, exist_ok=True)
    model.separate_comps(dir_results=outdir, spec_comp_ind={0: [2], 1: [0]})
    for n, f in enumerate(model.files["spat_comp"]):
        out["sub_sep%d" % n] = wavfile.read(f)[1]
    model.separate_comps(dir_results=outdir)  # default: one source per spectral component
    for n, f in enumerate(model.files["spat_comp"]):
        out["sub_all%d" % n] = wavfile.read(f)[1]


def main():
    ref = _py2shim.load()
    am = ref["audioModel"]
    out = {}
    run_update_mix(am, out)
    print("update_mix_matrix ok")
    run_update_spec(am, out)
    print("update_spectral_components ok")
    run_filter_stft(ref, out)
    print("filter_stft ok", out["fs_y3"].shape)
    run_time_blobs(am, out)
    print("time blobs ok", out["tb_logliks"])
    run_lambda_corr(am, out)
    print("lambdaCorr ok", out["lc_logliks"])
    run_redraw(am, out)
    print("re-draw ok", out["rd_TW1"].sum())
    run_subset(am, out)
    print("subset ok")
    np.savez_compressed(os.path.join(GOLD, "boundary.npz"), **out)


if __name__ == "__main__":
    main()
