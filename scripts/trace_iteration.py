#!/usr/bin/env python
"""Kernel timeline of ONE GEM iteration (torch.profiler / CUPTI) of the bench workload, optionally
sharded (run under torchrun): start, duration, stream and name of every kernel / memcpy on rank 0 --
to see what the collectives of the frequency / frame partitions overlap with.

    torchrun --nproc-per-node 2 scripts/trace_iteration.py --shard freq [--channels 4 --model conv --rank 4]
"""
import argparse
import os
import sys

import numpy as np


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard", default="freq")
    ap.add_argument("--duration-s", type=float, default=600.0)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--model", default="inst")
    ap.add_argument("--rank", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/trace.txt")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    from pyfasst_b200.engine import Comm
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        # (NCCL's own stream at high priority: the collectives of the frequency partition overlap
        # the contraction kernels of the next component, which fill the GPU)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
        comm = Comm()
    block = bench.synth_mix(args.duration_s, channels=args.channels)
    pcm = block if world == 1 else np.ascontiguousarray(np.tile(block, (world, 1)))
    a = ao.AudioObject("synthetic_mix.wav")
    a._samplerate = bench.FS
    a._set_raw(pcm)
    np.random.seed(0)
    cls = am.MultiChanNMFConv if args.model == "conv" else am.MultiChanNMFInst_FASST
    m = cls(audio=a, nbComps=bench.NSRC, nbNMFComps=bench.NNMF, spatial_rank=args.rank,
            wlen=bench.WLEN, hopsize=bench.HOP, iter_num=8, ann_PSD_lim=[None, None],
            comm=comm, shard=args.shard)
    if args.model == "conv":
        m.makeItConvolutive()
    eng = m._engine()
    ll = torch.ones(8, dtype=torch.float64, device=eng.dev)
    eng.iter_dev.zero_(); eng.flags.zero_(); eng.totals.zero_()
    for _ in range(4):
        eng.gem_iteration(8, ll)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.gem_iteration(8, ll)
        eng.gem_iteration(8, ll)
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        # the second iteration: after the largest gap-free midpoint -- simply print everything
        with open(args.out, "w") as f:
            for e in evs:
                f.write("%9.1f %8.1f  %s\n" % (e.time_range.start - t0,
                                               e.time_range.end - e.time_range.start, e.name[:90]))
        print("wrote", args.out, len(evs), "events; span %.1f us" % (evs[-1].time_range.end - t0))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
