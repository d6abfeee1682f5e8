"""The parameter matrices a model creates itself live in page-locked host memory
(FASST._host_param): they must stay plain, writable NumPy arrays, be updated IN PLACE by
estim_param_a_post_model like the reference's (audioModel.py:1573, :1725), and give bit-identical
results to pageable arrays (PYFASST_PINNED_PARAMS=0) -- the memory kind only changes the DMA."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FS = 44100


def _model(monkeypatch, pinned):
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    from bench import synth_mix
    monkeypatch.setenv("PYFASST_PINNED_PARAMS", "1" if pinned else "0")
    audio = ao.AudioObject("synthetic_mix.wav")
    audio._samplerate = FS
    audio._set_raw(synth_mix(50.0))  # 4 309 frames: TW = 32 x 4309 float64 = 1.1 MB per source
    np.random.seed(0)
    return am.MultiChanNMFInst_FASST(audio=audio, nbComps=2, nbNMFComps=32, spatial_rank=2,
                                     wlen=2048, hopsize=512, iter_num=3,
                                     ann_PSD_lim=[None, None], compute_dtype="float32")


def test_parameter_matrices_are_page_locked_numpy_arrays(monkeypatch):
    m = _model(monkeypatch, True)
    tws = [m.spec_comps[s]['factor'][0]['TW'] for s in range(2)]
    for tw in tws:
        assert type(tw) is np.ndarray and tw.dtype == np.float64 and tw.flags.writeable
        assert tw.flags.c_contiguous and tw.nbytes >= (1 << 20)
        assert torch.from_numpy(tw).is_pinned()
    before = [tw.copy() for tw in tws]
    ll = np.asarray(m.estim_param_a_post_model())
    assert np.isfinite(ll).all()
    for s, tw in enumerate(tws):  # the SAME arrays hold the updated parameters
        assert m.spec_comps[s]['factor'][0]['TW'] is tw
        assert not np.array_equal(tw, before[s]) and np.isfinite(tw).all()

    p = _model(monkeypatch, False)
    assert not torch.from_numpy(p.spec_comps[0]['factor'][0]['TW']).is_pinned()
    llp = np.asarray(p.estim_param_a_post_model())
    assert np.array_equal(ll, llp)
    for s in range(2):
        for name in ('FB', 'TW'):
            assert np.array_equal(m.spec_comps[s]['factor'][0][name],
                                  p.spec_comps[s]['factor'][0][name]), (s, name)
        assert np.array_equal(m.spat_comps[s]['params'], p.spat_comps[s]['params'])
