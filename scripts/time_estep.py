#!/usr/bin/env python
"""E-step launch time (CUDA events) at the configs[1] shape on random device-resident inputs --
for kernel-variant sweeps (PYFASST_B200_LIB=... selects the build)."""
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
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--frames", type=int, default=51682)
    ap.add_argument("--F", type=int, default=1025)
    ap.add_argument("--J", type=int, default=4)
    ap.add_argument("--I", type=int, default=2)
    ap.add_argument("--K", type=int, default=32)
    ap.add_argument("--rank", type=int, default=2)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--conv", action="store_true")
    ap.add_argument("--for-update", action="store_true",
                    help="the E-step as the GEM loop calls it (instantaneous mixing: pf_estep_stereo_inst)")
    args = ap.parse_args()
    k = CudaKernels()
    F, N, J, K, I = args.F, args.frames, args.J, args.K, args.I
    eng = GemEngine(k, F, N, dtype=args.dtype)
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((2 * I, F, eng.ld), generator=g, device="cuda", dtype=eng.tdtype)
    X[:, :, N:] = 0
    eng.set_X_planes(X)
    rng = np.random.default_rng(0)
    spat, spec = {}, {}
    for j in range(J):
        p = rng.standard_normal((I, args.rank))
        if args.conv:
            p = rng.standard_normal((args.rank, I, F)) + 1j * rng.standard_normal((args.rank, I, F))
        spat[j] = {"time_dep": "indep", "mix_type": "conv" if args.conv else "inst",
                   "frdm_prior": "free", "params": p}
        spec[j] = {"spat_comp_ind": j, "factor": {0: {
            "FB": np.abs(rng.standard_normal((F, K))) + 0.25, "FW": np.eye(K),
            "TW": np.abs(rng.standard_normal((K, N))) + 0.25, "TB": [],
            "FB_frdm_prior": "free", "FW_frdm_prior": "fixed", "TW_frdm_prior": "free",
            "TB_frdm_prior": [], "TW_constr": "NMF"}}}
    psd = np.full(F, 0.05)
    eng.set_noise("no_ann", psd, psd / 100, psd)
    eng.set_model(spat, spec)
    eng.compute_powers()
    for _ in range(3):
        eng.estep(args.for_update)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.reps):
        eng.estep(args.for_update)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.reps
    sz = 4 if args.dtype == "float32" else 8
    gb = sz * (2 * I + 2 * J) * F * N / 1e9
    print("%s estep %.4f ms  %.0f GB/s  (lib %s)" % (args.dtype, ms, gb / ms * 1e3,
                                                    os.environ.get("PYFASST_B200_LIB", "default")))


if __name__ == "__main__":
    main()
