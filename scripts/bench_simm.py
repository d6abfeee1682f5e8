#!/usr/bin/env python
"""Kernel-level benchmark of BASELINE.json configs[2]: Stereo_SIMM on a synthetic 10-min stereo
spectrogram with a 480-pitch F0 dictionary (F=1025, N=103362 frames at hop 256, NF0=480, P=30,
K=4, R=40).  Prints one JSON line: TF-bins*iterations/s of the device-resident
update loop (CUDA events), the tensor-pipe figure of the dense contractions, and the oracle
(NumPy float64 restatement of the reference) on a crop as the CPU baseline.

    python scripts/bench_simm.py [--frames N] [--steps K] [--warmup W] [--no-cpu-baseline]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

F, NF0, P, K, R = 1025, 480, 30, 4, 40


def problem(N, seed=1234):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    WF0 = np.abs(rng.standard_normal((F, NF0), dtype=f32))
    WF0 /= WF0.sum(axis=0)
    WG = np.abs(rng.standard_normal((F, P), dtype=f32))
    tilt = np.linspace(3.0, 0.1, F, dtype=f32)[:, None]
    SXR = rng.standard_normal((F, N), dtype=f32) ** 2 * tilt
    SXL = rng.standard_normal((F, N), dtype=f32) ** 2 * tilt
    init = dict(HGAMMA=np.abs(rng.standard_normal((P, K), dtype=f32)),
                HPHI=np.abs(rng.standard_normal((K, N), dtype=f32)),
                HF0=np.abs(rng.standard_normal((NF0, N), dtype=f32)),
                WM=np.abs(rng.standard_normal((F, R), dtype=f32)),
                HM=np.abs(rng.standard_normal((R, N), dtype=f32)),
                beta=rng.random(R))
    return SXR, SXL, WF0, WG, init


def dense_flops(N):
    """Algorithmic flops of the contractions with an F x N operand in one Stereo_SIMM iteration
    as this implementation runs them (2 m n k each)."""
    ldn = (N + 3) // 4 * 4
    g = 0
    g += 2 * NF0 * (2 * ldn) * F          # WF0^T (num | den)
    g += 2 * F * N * NF0                  # SF0 = WF0 HF0
    g += 2 * K * (2 * ldn) * F            # WPHI^T (num | den)
    g += 2 * R * (4 * ldn) * F            # WM^T (T_R T_L I_R I_L)
    g += 2 * 2 * F * K * ldn              # (num | den) HPHI^T
    g += 2 * 4 * 2 * F * R * ldn          # plane_q HM^T for the WM and beta updates
    g += 4 * 2 * 2 * F * N * R            # SM_c = (WM beta_c^2) HM, 4 times, 2 channels
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=103362)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    from pyfasst_b200._lib import CudaKernels
    from pyfasst_b200.simm_engine import SimmEngine
    N = args.frames
    SXR, SXL, WF0, WG, init = problem(N)
    k = CudaKernels()
    eng = SimmEngine(k, [SXR, SXL], WF0, WG, init["HGAMMA"], init["HPHI"], init["HF0"], init["WM"],
                     init["HM"], betaR=init["beta"], n_iter=args.steps + args.warmup)
    for _ in range(args.warmup):
        eng.iterate()
    torch.cuda.synchronize()
    l0 = k.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        eng.iterate()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    launches = k.launch_count() - l0
    r = eng.results()
    finite = all(np.isfinite(np.asarray(v)).all() for v in r.values())
    bins = F * N
    line = {"metric": "simm_tf_bins_iters_per_s", "value": bins / (ms * 1e-3),
            "unit": "TF-bins*iters/s", "n_gpus": 1, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "data": "synthetic", "ms_per_step": ms, "steps": args.steps,
            "warmup": args.warmup, "dtype": "f32 (3xTF32 tensor-core products)",
            "config": {"workload": "configs[2]: Stereo_SIMM, F=%d N=%d NF0=%d P=%d K=%d R=%d"
                                   % (F, N, NF0, P, K, R), "tf_bins": bins},
            "gpu_launches": int(launches), "finite": bool(finite),
            "dense": {"gflop_per_iter": dense_flops(N) / 1e9,
                      "tflops_fp32_equivalent": dense_flops(N) / (ms * 1e-3) / 1e12,
                      "note": "algorithmic fp32 flops of the F x N contractions / whole-iteration "
                              "time; each product issues 3 tf32 MMAs (3xTF32)"}}
    if not args.no_cpu_baseline:
        from oracle import simm_oracle as so
        Nc = args.cpu_frames
        a = [np.asarray(x[:, :Nc], dtype=np.float64) for x in (SXR, SXL)]
        t = time.perf_counter()
        so.stereo_simm(a[0], a[1], WF0.astype(np.float64), WG.astype(np.float64),
                       init["HGAMMA"].astype(np.float64), init["HPHI"][:, :Nc].astype(np.float64),
                       init["HF0"][:, :Nc].astype(np.float64), init["WM"].astype(np.float64),
                       init["HM"][:, :Nc].astype(np.float64), init["beta"], numberOfIterations=2)
        dt = (time.perf_counter() - t) / 2
        line["cpu_baseline"] = {"value": F * Nc / dt, "unit": "TF-bins*iters/s",
                                "cores": os.cpu_count(), "kind": "port",
                                "sample": "2 Stereo_SIMM iterations of the oracle on %d frames "
                                          "(%.2f s per iteration)" % (Nc, dt)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
