#!/usr/bin/env python
"""One 30-s clip (configs[4] shape), a few GEM iterations on the default stream -- for the ncu launch
list (kernel times of a clip-sized iteration)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    pcm = bench.synth_mix(30.0, seed=1234)
    a = ao.AudioObject("clip.wav")
    a._samplerate = bench.FS
    a._set_raw(pcm)
    np.random.seed(0)
    m = am.MultiChanNMFInst_FASST(audio=a, nbComps=bench.NSRC, nbNMFComps=bench.NNMF, spatial_rank=1,
                                  wlen=bench.WLEN, hopsize=bench.HOP, iter_num=6,
                                  ann_PSD_lim=[None, None])
    os.environ["PYFASST_STREAMS"] = os.environ.get("PYFASST_STREAMS", "0")
    eng = m._engine()
    ll = torch.ones(6, dtype=torch.float64, device=eng.dev)
    eng.iter_dev.zero_(); eng.flags.zero_(); eng.totals.zero_()
    for _ in range(6):
        eng.gem_iteration(6, ll)
    torch.cuda.synchronize()
    print("ok", ll.cpu().numpy()[-1])


if __name__ == "__main__":
    main()
