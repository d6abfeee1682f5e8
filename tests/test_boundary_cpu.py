"""The rest of the drop-in surface against golden vectors made by executing
the reference (oracle/make_golden_boundary.py): update_mix_matrix and update_spectral_components
called by hand, filter_stft, time-blob factors (TB), the lambdaCorr penalty, the random re-draw
of a vanished TW, separate_comps with spectral components left out, and the `pyfasst` alias
package.  Here on the NumPy kernel specification (float64); tests/test_boundary_gpu.py runs the
same checks on the CUDA kernels."""
import os
import sys

import numpy as np
import pytest

import pyfasst_b200.audioModel as am
from tests.test_api_cpu import GOLDEN, rel_err
from tests.test_sourcefilter_cpu import AllFakeKernels

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
KW = dict(wlen=256, hopsize=64, verbose=0, ann_PSD_lim=[None, None])


def load():
    return np.load(os.path.join(GOLDEN, "boundary.npz"))


def inst(kernels, dtype, rank=2, iters=3, **kw):
    np.random.seed(0)
    args = dict(KW)
    args.update(kw)
    return am.MultiChanNMFInst_FASST(audio=os.path.join(GOLDEN, "mix_inst.wav"), nbComps=3,
                                     nbNMFComps=4, spatial_rank=rank, iter_num=iters,
                                     kernels=kernels, compute_dtype=dtype, **args)


def conv(kernels, dtype, rank=2, iters=3, **kw):
    np.random.seed(0)
    args = dict(KW)
    args.update(kw)
    m = am.MultiChanNMFConv(audio=os.path.join(GOLDEN, "mix_conv.wav"), nbComps=2, nbNMFComps=4,
                            spatial_rank=rank, iter_num=iters, kernels=kernels,
                            compute_dtype=dtype, **args)
    m.makeItConvolutive()
    return m


def worst(model, g, prefix, names=("FB", "TW"), with_A=True):
    w = 0.0
    for k in range(len(model.spec_comps)):
        for nm in names:
            w = max(w, rel_err(model.spec_comps[k]["factor"][0][nm], g["%s_%s%d" % (prefix, nm, k)]))
        if with_A:
            w = max(w, rel_err(model.spat_comps[k]["params"], g["%s_A%d" % (prefix, k)]))
    return w


def check_update_mix_matrix(kernels, dtype, tol):
    g = load()
    for tag, m in (("inst", inst(kernels, dtype)), ("conv", conv(kernels, dtype))):
        if tag == "inst":
            m.spat_comps[1]["frdm_prior"] = "fixed"
        mm = np.array(g["umm_%s_mix0" % tag])
        _, _, rpi = m.retrieve_subsrc_params()
        m.update_mix_matrix(g["umm_%s_hat_Rxs" % tag], g["umm_%s_hat_Rss" % tag], mm, rpi)
        assert rel_err(mm, g["umm_%s_mix1" % tag]) < tol          # updated in place
        for j in m.spat_comps:
            assert rel_err(m.spat_comps[j]["params"], g["umm_%s_A%d" % (tag, j)]) < tol
    with pytest.raises(ValueError):
        m.update_mix_matrix(g["umm_conv_hat_Rxs"], g["umm_conv_hat_Rss"], mm[:1], rpi)


def check_update_spectral_components(kernels, dtype, tol):
    g = load()
    m = inst(kernels, dtype, nmfUpdateCoeff=0.7)
    m.spec_comps[1]["factor"][0]["FB_frdm_prior"] = "fixed"
    m.spec_comps[2]["factor"][0]["TW_frdm_prior"] = "fixed"
    m.update_spectral_components(g["usc_hat_W"])
    assert worst(m, g, "usc") < tol
    with pytest.raises(ValueError):
        m.update_spectral_components(g["usc_hat_W"][:2])


def check_filter_stft(kernels, tol):
    from pyfasst_b200.tftransforms.stft import filter_stft
    from pyfasst_b200.tools.utils import sinebell
    g = load()
    win = np.hanning(256)
    y3 = filter_stft(g["fs_data"], g["fs_W3"], synthWindow=win, hopsize=64, nfft=256, fs=8000,
                     kernels=kernels)
    y4 = filter_stft(g["fs_data"], g["fs_W4"], synthWindow=win, hopsize=64, nfft=256, fs=8000,
                     kernels=kernels)
    assert y3.shape == g["fs_y3"].shape and np.abs(y3 - g["fs_y3"]).max() < tol
    assert np.abs(y4 - g["fs_y4"]).max() < tol
    W3 = g["fs_W3"]
    Wb = np.concatenate([W3, W3[:, :, :-1][:, :, ::-1]], axis=2)[:, :, :257]
    yb = filter_stft(g["fs_data"], Wb, analysisWindow=sinebell(256), synthWindow=win, hopsize=64,
                     nfft=512, fs=8000, kernels=kernels)
    assert np.abs(yb - g["fs_y3b"]).max() < tol
    with pytest.raises(AttributeError):
        filter_stft(g["fs_data"], g["fs_W3"][:1], synthWindow=win, hopsize=64, nfft=256,
                    kernels=kernels)
    with pytest.raises(AttributeError):
        filter_stft(g["fs_data"], g["fs_W3"], synthWindow=win, hopsize=64, nfft=512,
                    kernels=kernels)


def check_time_blobs(kernels, dtype, tol_renorm, tol_ll, tol_final):
    from make_golden_boundary import add_time_blobs
    g = load()
    m = inst(kernels, dtype, rank=1, iters=4)
    add_time_blobs(m)
    m.spec_comps[2]["factor"][0]["TB_frdm_prior"] = "fixed"
    m.renormalize_parameters()
    assert worst(m, g, "tb_renorm", ("TW", "TB", "FB"), with_A=False) < tol_renorm
    assert rel_err(m.comp_spat_comp_power(0), g["tb_V0"]) < 10 * tol_renorm
    ll = m.estim_param_a_post_model()
    assert np.abs(ll / g["tb_logliks"] - 1).max() < tol_ll
    assert worst(m, g, "tb_final", ("FB", "TW", "TB")) < tol_final
    assert rel_err(m.spec_comps[2]["factor"][0]["TB"], g["tb_renorm_TB2"]) < tol_renorm  # fixed


def check_lambda_corr(kernels, dtype, tol_ll, tol_final):
    g = load()
    m = inst(kernels, dtype, rank=2, iters=4, lambdaCorr=0.1)
    ll = m.estim_param_a_post_model()
    assert np.abs(ll / g["lc_logliks"] - 1).max() < tol_ll
    assert worst(m, g, "lc_final") < tol_final


def check_redraw(kernels, dtype, tol):
    """A TW whose sum falls below eps is re-drawn with np.random, in renormalize_parameters()
    called by hand and inside the estimation loop (replayed from the iteration it happens at)."""
    g = load()
    m = inst(kernels, dtype, rank=1)
    m.spec_comps[1]["factor"][0]["TW"][:] = 1e-14
    np.random.seed(77)
    m.renormalize_parameters()
    assert worst(m, g, "rd") < tol
    follow = np.random.rand()
    np.random.seed(77)
    np.random.randn(*m.spec_comps[1]["factor"][0]["TW"].shape)
    assert follow == np.random.rand(), "exactly one randn(K, N) draw is consumed"
    # inside the loop: component 1 dies at iteration 0; every later iteration sees the re-drawn TW
    # (mixing parameters fixed: with a dead source the reference's mixing update is singular)
    m = inst(kernels, dtype, rank=1, iters=3)
    for sc in m.spat_comps.values():
        sc["frdm_prior"] = "fixed"
    m.spec_comps[1]["factor"][0]["TW"][:] = 1e-30
    np.random.seed(5)
    ll = m.estim_param_a_post_model()
    assert np.all(np.isfinite(ll))
    tw = m.spec_comps[1]["factor"][0]["TW"]
    assert tw.sum() > 1e-9, "the dead component came back to life"


def check_subset_separation(kernels, dtype, max_lsb):
    g = load()
    m = inst(kernels, dtype, rank=2, iters=3)
    m.estim_param_a_post_model()
    pcm = m.separate_comps_pcm({0: [2], 1: [0]})
    for n in range(2):
        d = np.abs(pcm[n].astype(int) - g["sub_sep%d" % n].astype(int))
        assert d.max() <= max_lsb, (n, d.max())
    pcm = m.separate_comps_pcm()
    for n in range(3):
        d = np.abs(pcm[n].astype(int) - g["sub_all%d" % n].astype(int))
        assert d.max() <= max_lsb


# ---- CPU: the float64 kernel specification -----------------------------------------------------
def test_update_mix_matrix():
    check_update_mix_matrix(AllFakeKernels(), "float64", 1e-12)


def test_update_spectral_components():
    check_update_spectral_components(AllFakeKernels(), "float64", 1e-12)


def test_filter_stft():
    check_filter_stft(AllFakeKernels(), 1e-12)


def test_time_blobs():
    check_time_blobs(AllFakeKernels(), "float64", 1e-12, 1e-12, 1e-11)


def test_lambda_corr():
    check_lambda_corr(AllFakeKernels(), "float64", 1e-12, 1e-11)


def test_tw_redraw():
    check_redraw(AllFakeKernels(), "float64", 1e-12)


def test_subset_separation():
    check_subset_separation(AllFakeKernels(), "float64", 0)


def test_cx_all_channel_pairs():
    m = inst(AllFakeKernels(), "float64", rank=1)
    g = np.load(os.path.join(GOLDEN, "fasst_inst_r1.npz"))
    np.testing.assert_allclose(m.Cx, g["Cx"], atol=1e-12)
    assert m.Cx.shape[0] == 3


def test_shared_arrays_route_to_general_engine():
    """Two components aliasing ONE host FB array: the reference rescales the shared object once per
    component (quirk Q11); the fast engine would upload two copies."""
    m = inst(AllFakeKernels(), "float64", rank=1)
    assert not m._general_structure()
    m.spec_comps[1]["factor"][0]["FB"] = m.spec_comps[0]["factor"][0]["FB"]
    assert m._general_structure()


def test_ann_psd_lim_default_is_not_shared():
    a = inst(AllFakeKernels(), "float64", rank=1)
    np.random.seed(0)
    b = am.MultiChanNMFInst_FASST(audio=os.path.join(GOLDEN, "mix_inst.wav"), nbComps=3,
                                  wlen=128, hopsize=32, kernels=AllFakeKernels(),
                                  compute_dtype="float64")
    c = am.MultiChanNMFInst_FASST(audio=os.path.join(GOLDEN, "mix_inst.wav"), nbComps=3,
                                  wlen=256, hopsize=64, kernels=AllFakeKernels(),
                                  compute_dtype="float64")
    assert b.noise["ann_PSD_lim"][0].shape == (65,) and c.noise["ann_PSD_lim"][0].shape == (129,)
    np.testing.assert_allclose(c.noise["ann_PSD_lim"][0], a.noise["ann_PSD_lim"][0])


def test_pyfasst_alias_package():
    import pyfasst.audioModel as ref_am
    import pyfasst.SeparateLeadStereo.SeparateLeadStereoTF as ref_sls
    from pyfasst.SeparateLeadStereo.SIMM.SIMM import SIMM, Stereo_SIMM  # noqa: F401
    from pyfasst.tftransforms.stft import STFT, filter_stft, istft, stft  # noqa: F401
    assert ref_am is am and ref_am.FASST is am.FASST
    assert hasattr(ref_sls, "SeparateLeadProcess")
    for name in ("update_mix_matrix", "update_spectral_components", "renormalize_parameters",
                 "GEM_iteration", "estim_param_a_post_model", "compute_suff_stat",
                 "retrieve_subsrc_params", "comp_spat_comp_power", "separate_spat_comps",
                 "separate_comps", "comp_transf_Cx"):
        assert callable(getattr(ref_am.FASST, name))
    with pytest.raises(ImportError):
        import pyfasst.demixTF  # noqa: F401  (out of scope)
