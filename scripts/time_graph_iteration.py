#!/usr/bin/env python
"""GEM iteration of the bench workload: eager launches against CUDA-graph replay (CUDA events)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyfasst_b200.audioModel as am  # noqa: E402
import pyfasst_b200.audioObject as ao  # noqa: E402


def main():
    pcm = bench.synth_mix(600.0)
    a = ao.AudioObject("synthetic_mix.wav")
    a._samplerate = bench.FS
    a._set_raw(pcm)
    np.random.seed(0)
    n = 46
    m = am.MultiChanNMFInst_FASST(audio=a, nbComps=bench.NSRC, nbNMFComps=bench.NNMF, spatial_rank=2,
                                  wlen=bench.WLEN, hopsize=bench.HOP, iter_num=n, ann_PSD_lim=[None, None])
    eng = m._engine()
    ll = torch.ones(n, dtype=torch.float64, device=eng.dev)
    eng.iter_dev.zero_(); eng.flags.zero_(); eng.totals.zero_()
    for _ in range(3):
        eng.gem_iteration(n, ll)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(20):
        eng.gem_iteration(n, ll)
    t1.record()
    torch.cuda.synchronize()
    print("eager  %.4f ms per iteration" % (t0.elapsed_time(t1) / 20))
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        g.capture_begin()
        eng.gem_iteration(n, ll)
        g.capture_end()
    torch.cuda.current_stream().wait_stream(cap)
    g.replay()
    torch.cuda.synchronize()
    t0.record()
    for _ in range(20):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    print("graph  %.4f ms per iteration" % (t0.elapsed_time(t1) / 20))
    os.environ["PYFASST_STREAMS"] = "0"


if __name__ == "__main__":
    main()
