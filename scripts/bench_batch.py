#!/usr/bin/env python
"""BASELINE.json configs[4]: a batch of synthetic 30-s stereo clips end to end -- PCM on the host
-> STFT -> GEM (100 iterations) -> Wiener -> inverse STFT -> PCM on the host -- with
pyfasst_b200.batch (one stream + one CUDA graph per clip).  One process per GPU; each rank takes
clips[rank::world]; no collective on the data path.  Prints one JSON line (rank 0): aggregate
TF-bins*iterations/s over all ranks (max of the per-rank wall times).

    python scripts/bench_batch.py [--clips 32] [--iters 100] [--seconds 30] [--no-graph]
    torchrun --nproc-per-node 8 scripts/bench_batch.py --clips 256
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# independent clips: the default 8 hardware queues (32 streams alias onto them) keep fewer clips in
# flight at once, i.e. more of a clip's planes stay L2 resident -- measured faster than 32
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("PF_BATCH_CONNECTIONS", "8"))
import bench  # noqa: E402  (synth_mix, constants)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=32, help="total number of clips (all ranks)")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--sequential", action="store_true",
                    help="the plain loop over models (public API), for comparison")
    ap.add_argument("--sequential-graph", action="store_true",
                    help="the plain loop over models, each replaying its iteration as a CUDA graph")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    from pyfasst_b200 import batch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine = list(range(rank, args.clips, world))
    pcms = [bench.synth_mix(args.seconds, seed=1234 + i) for i in mine]

    seq = args.sequential or args.sequential_graph
    stages = {}

    def run(pcm_list, iters):
        models = []
        t_a = time.perf_counter()
        for i, pcm in zip(mine, pcm_list):
            a = ao.AudioObject("clip_%d.wav" % i)
            a._samplerate = bench.FS
            a._set_raw(pcm)
            np.random.seed(i)
            models.append(am.MultiChanNMFInst_FASST(audio=a, nbComps=bench.NSRC,
                                                    nbNMFComps=bench.NNMF, spatial_rank=1,
                                                    wlen=bench.WLEN, hopsize=bench.HOP,
                                                    iter_num=iters, ann_PSD_lim=[None, None],
                                                    use_cuda_graph=args.sequential_graph))
        torch.cuda.synchronize()
        t_b = time.perf_counter()
        if seq:
            lls = [m.estim_param_a_post_model() for m in models]
        else:
            lls = batch.estimate_batch(models, use_cuda_graph=not args.no_graph)
        torch.cuda.synchronize()
        t_c = time.perf_counter()
        if seq:
            out = [m.separate_comps_pcm() for m in models]
        else:
            out = batch.separate_batch(models)
        torch.cuda.synchronize()
        t_d = time.perf_counter()
        stages.update(construct_stft_s=t_b - t_a, estimate_s=t_c - t_b, separate_s=t_d - t_c)
        return models, lls, out

    run(pcms[:2], 3)  # warm-up (library load, kernel attributes, allocator)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    models, lls, out = run(pcms, args.iters)
    dt = time.perf_counter() - t0
    dt_t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    dt = float(dt_t.item())
    if rank == 0:
        F, N = models[0].nbFreqsSigRepr, models[0].nbFramesSigRepr
        bins = F * N * args.clips
        line = {"metric": "gem_tf_bins_iters_per_s", "value": bins * args.iters / dt,
                "unit": "TF-bins*iters/s", "n_gpus": world, "wall_s": dt,
                "config": {"workload": "configs[4]: %d synthetic %.0f-s stereo clips, STFT -> GEM "
                                       "(%d iterations, 4 sources x K=32, rank 1) -> Wiener -> iSTFT, "
                                       "host PCM in / host PCM out" % (args.clips, args.seconds,
                                                                       args.iters),
                           "clips_per_gpu": len(mine), "F": F, "N": N, "tf_bins_per_clip": F * N,
                           "mode": "sequential" if args.sequential else
                                   "sequential+graphs" if args.sequential_graph else
                                   ("streams" if args.no_graph else "streams+graphs")},
                "stages_s": stages,
                "scaling": "weak (independent clips, no collective)",
                "loglik_last_clip0": float(lls[0][-1]),
                "finite": bool(all(np.isfinite(l).all() for l in lls))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
