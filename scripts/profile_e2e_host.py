#!/usr/bin/env python
"""Host-side profile (cProfile) of the end-to-end leg of bench.py at N = 1: comp_transf_Cx +
estim_param_a_post_model on the 10-min mixture (20 iterations)."""
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
    pcm = bench.synth_mix(600.0)
    pinned = torch.from_numpy(pcm).pin_memory()

    def make():
        a = ao.AudioObject("synthetic_mix.wav")
        a._samplerate = bench.FS
        a._set_raw(pcm)
        np.random.seed(0)
        return am.MultiChanNMFInst_FASST(audio=a, nbComps=bench.NSRC, nbNMFComps=bench.NNMF,
                                         spatial_rank=2, wlen=bench.WLEN, hopsize=bench.HOP,
                                         iter_num=20, ann_PSD_lim=[None, None])

    def run(m):
        m.audioObject._set_raw(pinned)
        m.comp_transf_Cx()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        m.estim_param_a_post_model()
        torch.cuda.synchronize()
        return t1
    for _ in range(2):
        run(make())
    m = make()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    t1 = run(m)
    pr.disable()
    t2 = time.perf_counter()
    print("comp_transf_Cx %.2f ms, estim_param_a_post_model %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    print(m._last_engine_stats)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(40)


if __name__ == "__main__":
    main()
