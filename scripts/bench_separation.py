#!/usr/bin/env python
"""Separation phase of configs[1] on one B200: Wiener filter (K6) + inverse STFT with overlap-add
+ int16 conversion for all sources, end to end through FASST.separate_comps_pcm (device work +
device->host copy of the PCM)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    dur = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    pcm = bench.synth_mix(dur)
    a = ao.AudioObject("synthetic_mix.wav")
    a._samplerate = bench.FS
    a._set_raw(pcm)
    np.random.seed(0)
    m = am.MultiChanNMFInst_FASST(audio=a, nbComps=bench.NSRC, nbNMFComps=bench.NNMF,
                                  spatial_rank=bench.RANK, wlen=bench.WLEN, hopsize=bench.HOP,
                                  iter_num=2, ann_PSD_lim=[None, None])
    m.estim_param_a_post_model()
    times = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = m.separate_comps_pcm()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    bins = m.nbFreqsSigRepr * m.nbFramesSigRepr
    print(json.dumps({"workload": "separate_comps_pcm, %d sources, %.0f-s stereo" % (bench.NSRC, dur),
                      "tf_bins": bins, "seconds": times, "best_s": min(times),
                      "output_mb": out.nbytes / 1e6}))


if __name__ == "__main__":
    main()
