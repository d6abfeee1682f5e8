"""The general-I (3- and 4-channel) extension: the reference is stereo only, so the checker is the
generalised oracle (batched I x I inverse), which these tests first prove equal to the stereo
oracle -- itself pinned by the reference's golden vectors -- at I = 2; then the engine on the NumPy
kernel specifications is compared with it at I = 4.  CPU only."""
import os

import numpy as np
import pytest
import scipy.io.wavfile as wavfile
from numpy.testing import assert_allclose

from oracle import fasst_oracle as fo
from tests.fake_kernels import FakeKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def multichannel_audio(nch, L=3000, seed=0, fs=8000):
    """(fs, data[L, nch] scaled, maxdata): three coloured sources through random mixing filters."""
    rng = np.random.default_rng(seed)
    src = np.cumsum(rng.standard_normal((3, L)), axis=1)
    src -= src.mean(axis=1, keepdims=True)
    src[1] = np.diff(src[1], prepend=0) * 5
    env = (np.sin(2 * np.pi * np.arange(L)[None] / np.array([[700.], [450.], [1100.]])) > -0.3)
    mix = np.zeros((L, nch))
    for j in range(3):
        h = rng.standard_normal((nch, 6)) * np.exp(-np.arange(6) / 2.0)
        for c in range(nch):
            mix[:, c] += np.convolve(src[j] * env[j], h[c])[:L]
    mix += 0.01 * rng.standard_normal(mix.shape)
    pcm = np.int16(np.round(20000 * mix / np.abs(mix).max()))
    maxdata = 1.1 * np.abs(pcm).max()
    return pcm, (fs, pcm / maxdata, maxdata)


def make_oracle(audio, conv, rank, iters, generalised=None, seed=3):
    np.random.seed(seed)
    m = fo.OracleFASST(audio, nbComps=3, nbNMFComps=4, spatial_rank=rank, wlen=256, hopsize=64,
                       iter_num=iters, generalised=generalised)
    if conv:
        m.makeItConvolutive()
    return m


@pytest.mark.parametrize("conv,rank", [(False, 1), (True, 2)])
def test_generalised_oracle_equals_stereo_oracle(conv, rank):
    wav = os.path.join(GOLDEN, "mix_conv.wav" if conv else "mix_inst.wav")
    a = make_oracle(wav, conv, rank, 3, generalised=False)
    b = make_oracle(wav, conv, rank, 3, generalised=True)
    la, lb = a.estim_param_a_post_model(), b.estim_param_a_post_model()
    assert_allclose(lb, la, rtol=1e-10)
    for j in a.spat_comps:
        assert_allclose(b.spat_comps[j]["params"], a.spat_comps[j]["params"], rtol=1e-7, atol=1e-12)
        for key in ("FB", "TW"):
            assert_allclose(b.spec_comps[j]["factor"][0][key], a.spec_comps[j]["factor"][0][key],
                            rtol=1e-7, atol=1e-300)
    sa, sb = a.separate_signals(), b.separate_signals()
    assert_allclose(sb, sa, rtol=0, atol=1e-9 * np.abs(sa).max())


def run_engine(kernels, nch, conv, rank, iters, dtype):
    """The engine through the public API on a synthetic nch-channel mixture, next to the oracle."""
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    pcm, audio = multichannel_audio(nch)
    ref = make_oracle(audio, conv, rank, iters)
    a = ao.AudioObject("mix%d.wav" % nch)
    a._samplerate = audio[0]
    a._set_raw(pcm)
    np.random.seed(3)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    model = cls(audio=a, nbComps=3, nbNMFComps=4, spatial_rank=rank, wlen=256, hopsize=64,
                iter_num=iters, compute_dtype=dtype, kernels=kernels,
                ann_PSD_lim=[None, None])  # (the default list is shared between instances, as
    #                                          in the reference: audioModel.py:166, :305-321)
    if conv:
        model.makeItConvolutive()
    return ref, model


def compare(ref, model, tol):
    ll_ref = ref.estim_param_a_post_model()
    ll = model.estim_param_a_post_model()
    assert_allclose(ll, ll_ref, rtol=tol)
    for j in ref.spat_comps:
        pa, pb = np.asarray(model.spat_comps[j]["params"]), np.asarray(ref.spat_comps[j]["params"])
        assert np.abs(pa - pb).max() <= 10 * tol * np.abs(pb).max(), ("A", j)
        for key in ("FB", "TW"):
            xa = model.spec_comps[j]["factor"][0][key]
            xb = ref.spec_comps[j]["factor"][0][key]
            assert np.abs(xa - xb).max() <= 10 * tol * np.abs(xb).max(), (key, j)
    sig = ref.separate_signals()
    pcm_ref = [fo.pcm_from_float(sig[n][:ref.nframes_audio], ref.maxdata) for n in range(len(sig))]
    pcm = model.separate_comps_pcm()
    assert pcm.shape == (len(sig), ref.nframes_audio, ref.channels)
    for n in range(len(sig)):
        d = np.abs(pcm[n].astype(int) - pcm_ref[n].astype(int))
        assert d.max() <= (2 if tol < 1e-6 else 8), (n, d.max())


@pytest.mark.parametrize("nch,conv,rank", [(4, False, 2), (4, True, 2), (3, True, 1)])
def test_engine_multichannel_float64_spec(nch, conv, rank):
    ref, model = run_engine(FakeKernels(), nch, conv, rank, 3, "float64")
    compare(ref, model, 1e-9)


def test_stereo_through_the_general_kernels(monkeypatch):
    """PYFASST_FORCE_MULTI=1 routes a stereo model through the general-I kernels: same result as
    the reference-pinned stereo path."""
    monkeypatch.setenv("PYFASST_FORCE_MULTI", "1")
    ref, model = run_engine(FakeKernels(), 2, True, 2, 3, "float64")
    compare(ref, model, 1e-9)
