#!/usr/bin/env python
"""Small driver for ncu: a few GEM iterations of the bench workload (configs[1] shape) on
random device-resident inputs -- no STFT, no host transfers, no CPU baseline."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyfasst_b200._lib import CudaKernels  # noqa: E402
from pyfasst_b200.engine import GemEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--frames", type=int, default=51682)
    ap.add_argument("--F", type=int, default=1025)
    ap.add_argument("--J", type=int, default=4)
    ap.add_argument("--K", type=int, default=32)
    ap.add_argument("--rank", type=int, default=2)
    ap.add_argument("--dtype", default="float32")
    args = ap.parse_args()
    k = CudaKernels()
    F, N, J, K = args.F, args.frames, args.J, args.K
    eng = GemEngine(k, F, N, dtype=args.dtype)
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((4, F, eng.ld), generator=g, device="cuda", dtype=eng.tdtype)
    X[:, :, N:] = 0
    eng.set_X_planes(X)
    rng = np.random.default_rng(0)
    spat, spec = {}, {}
    for j in range(J):
        spat[j] = {"time_dep": "indep", "mix_type": "inst", "frdm_prior": "free",
                   "params": rng.standard_normal((2, args.rank))}
        spec[j] = {"spat_comp_ind": j, "factor": {0: {
            "FB": np.abs(rng.standard_normal((F, K))) + 0.25, "FW": np.eye(K),
            "TW": np.abs(rng.standard_normal((K, N))) + 0.25, "TB": [],
            "FB_frdm_prior": "free", "FW_frdm_prior": "fixed", "TW_frdm_prior": "free",
            "TB_frdm_prior": [], "TW_constr": "NMF"}}}
    psd = np.full(F, 0.05)
    eng.set_noise("ann", psd, psd / 100, psd)
    eng.set_model(spat, spec)
    ll = eng.run(args.iters)
    torch.cuda.synchronize()
    print("logliks", ll)


if __name__ == "__main__":
    main()
