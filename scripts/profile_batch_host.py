#!/usr/bin/env python
"""Host-side profile (cProfile) of the configs[4] stages on a few 30-s clips: where the time of
construct / estimate / separate goes when the GPU work is small."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    from pyfasst_b200 import batch
    nclips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    pcms = [bench.synth_mix(30.0, seed=1234 + i) for i in range(nclips)]

    def build(iters):
        models = []
        for i, pcm in enumerate(pcms):
            a = ao.AudioObject("clip_%d.wav" % i)
            a._samplerate = bench.FS
            a._set_raw(pcm)
            np.random.seed(i)
            models.append(am.MultiChanNMFInst_FASST(audio=a, nbComps=bench.NSRC, nbNMFComps=bench.NNMF,
                                                    spatial_rank=1, wlen=bench.WLEN, hopsize=bench.HOP,
                                                    iter_num=iters, ann_PSD_lim=[None, None]))
        return models
    ms = build(3)
    batch.estimate_batch(ms)
    batch.separate_batch(ms)
    torch.cuda.synchronize()
    for name, fn in (("construct", lambda: build(iters)),):
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        ms = fn()
        torch.cuda.synchronize()
        pr.disable()
        print("== %s: %.3f s for %d clips" % (name, time.perf_counter() - t0, nclips))
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    for name, fn in (("estimate", lambda: batch.estimate_batch(ms)), ("separate", lambda: batch.separate_batch(ms))):
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        fn()
        torch.cuda.synchronize()
        pr.disable()
        print("== %s: %.3f s for %d clips" % (name, time.perf_counter() - t0, nclips))
        pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
