"""pyfasst_b200.batch: several independent clips on their own streams / CUDA graphs give exactly
what the one-after-the-other public API gives."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_models(n, iters):
    import pyfasst_b200.audioModel as am
    models = []
    for i in range(n):
        np.random.seed(100 + i)
        wav = os.path.join(GOLDEN, "mix_inst.wav" if i % 2 == 0 else "mix_conv.wav")
        models.append(am.MultiChanNMFInst_FASST(audio=wav, nbComps=3, nbNMFComps=4,
                                                spatial_rank=1 + i % 2, wlen=256, hopsize=64,
                                                iter_num=iters))
    return models


@pytest.mark.parametrize("use_graph", [False, True])
def test_batch_matches_sequential(use_graph):
    from pyfasst_b200 import batch
    seq = make_models(4, 5)
    ll_seq = [m.estim_param_a_post_model() for m in seq]
    pcm_seq = [m.separate_comps_pcm() for m in seq]
    bat = make_models(4, 5)
    ll_bat = batch.estimate_batch(bat, use_cuda_graph=use_graph)
    pcm_bat = batch.separate_batch(bat)
    for a, b in zip(ll_seq, ll_bat):
        np.testing.assert_array_equal(a, b)
    for ms, mb in zip(seq, bat):
        for j in ms.spec_comps:
            for key in ("FB", "TW"):
                np.testing.assert_array_equal(ms.spec_comps[j]["factor"][0][key],
                                              mb.spec_comps[j]["factor"][0][key])
            np.testing.assert_array_equal(ms.spat_comps[j]["params"], mb.spat_comps[j]["params"])
    for a, b in zip(pcm_seq, pcm_bat):
        np.testing.assert_array_equal(a, b)
