#!/usr/bin/env python
"""GEM iteration time of multiChanSourceF0Filter (two-factor source/filter sources sharing the
1093-comb glottal dictionary + one residual NMF component) on one B200.

    python scripts/bench_sourcefilter.py [--duration-s 60] [--iters 5] [--comps 3]
"""
import argparse
import json
import os
import sys
import tempfile
import time


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--duration-s", type=float, default=60.0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--comps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bench
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    os.chdir(tempfile.mkdtemp())
    pcm = bench.synth_mix(args.duration_s, nsrc=args.comps)
    a = ao.AudioObject("synthetic_mix.wav")
    a._samplerate = bench.FS
    a._set_raw(pcm)
    t0 = time.perf_counter()
    m = am.multiChanSourceF0Filter(audio=a, nbComps=args.comps, iter_num=args.iters + args.warmup,
                                   ann_PSD_lim=[None, None])
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    k = m._k()
    eng = m._engine()
    total = args.iters + args.warmup
    logliks = torch.ones(total, dtype=torch.float64, device=eng.dev)
    eng.iter_dev.zero_()
    eng.flags.zero_()
    for _ in range(args.warmup):
        eng.gem_iteration(total, logliks)
    phases = {}
    marks = []

    def mark(label):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((label, ev))
    torch.cuda.synchronize()
    l0 = k.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        eng.gem_iteration(total, logliks, mark)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    for (la, ea), (lb, eb) in zip(marks[:-1], marks[1:]):
        if lb != "begin":
            phases[lb] = phases.get(lb, 0.0) + ea.elapsed_time(eb) / args.iters
    eng.check_flags()
    bins = m.nbFreqsSigRepr * m.nbFramesSigRepr
    print(json.dumps({
        "workload": "multiChanSourceF0Filter, %d sources (%d source/filter + 1 residual), "
                    "%d-comb dictionary, %.0f-s stereo 44.1 kHz, STFT 2048/512" %
                    (args.comps, args.comps - 1, m.nbSourceComps, args.duration_s),
        "F": m.nbFreqsSigRepr, "N": m.nbFramesSigRepr, "tf_bins": bins,
        "ms_per_iteration": ms, "tf_bins_iters_per_s": bins / (ms * 1e-3),
        "phases_ms": phases, "launches_per_iteration": (k.launch_count() - l0) / args.iters,
        "construction_s": t_build, "loglik_last": float(logliks[total - 1].item())}))


if __name__ == "__main__":
    main()
