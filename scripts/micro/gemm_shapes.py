#!/usr/bin/env python
"""Times pf_gemm_tf32x3(_splitk) on the seven contraction shapes of one Stereo_SIMM iteration of
BASELINE.json configs[2] (F=1025, N=103362, NF0=480, K=4, R=40; simm_engine.py: iterate), one
launch each between CUDA events, operands far larger than L2.  Prints one line per shape:
duration, fp32-equivalent TFLOP/s and the HBM rate of the compulsory traffic.

    python scripts/micro/gemm_shapes.py [--frames N] [--reps 5] [--only name,name]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

F, NF0, K, R = 1025, 480, 4, 40


def ru4(n):
    return (n + 3) // 4 * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=103362)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from pyfasst_b200._lib import CudaKernels
    k = CudaKernels()
    dev = torch.device("cuda", 0)
    N = args.frames
    ldn = ru4(N)
    g = torch.Generator(device=dev).manual_seed(1)

    def rnd(*shape):
        return torch.rand(shape, generator=g, device=dev, dtype=torch.float32)

    WF0, HF0 = rnd(F, NF0), rnd(NF0, ldn)
    WPHI, HPHI = rnd(F, ru4(K)), rnd(K, ldn)
    WM, HM = rnd(F, R), rnd(R, ldn)
    work = rnd(F, 4 * ldn)          # work_acc; its first half doubles as work_lead
    lead = work[:, :2 * ldn]
    C_f0 = torch.empty(NF0, 2 * ldn, device=dev)
    SF0 = torch.empty(F, ldn, device=dev)
    C_phi = torch.empty(K, 2 * ldn, device=dev)
    C_hm = torch.empty(R, 4 * ldn, device=dev)
    SM = torch.empty(F, 2 * ldn, device=dev)
    D = torch.empty(F, R, device=dev)
    tn = torch.empty(F, ru4(K), device=dev)
    ws = torch.empty(max(k.gemm_splitk_workspace_bytes(F, K, ldn),
                         k.gemm_splitk_workspace_bytes(F, R, ldn), 16) // 4, device=dev)
    plane = rnd(F, ldn)             # one contiguous plane (row stride ldn instead of 4 ldn)
    pl = F * ldn * 4  # bytes of one F x N plane
    # name, call, flops (2 m n k), compulsory bytes
    shapes = [
        ("C_f0 = WF0^T lead  [480 x 2N, K=F]",
         lambda: k.gemm_view(WF0, lead, C_f0, NF0, 2 * ldn, F, transA=True),
         2.0 * NF0 * 2 * ldn * F, 2 * pl + NF0 * 2 * ldn * 4),
        ("SF0 = WF0 HF0      [F x N, K=480]",
         lambda: k.gemm_view(WF0, HF0, SF0, F, N, NF0),
         2.0 * F * N * NF0, pl + NF0 * ldn * 4),
        ("C_phi = WPHI^T lead [4 x 2N, K=F]",
         lambda: k.gemm_view(WPHI, lead, C_phi, K, 2 * ldn, F, transA=True),
         2.0 * K * 2 * ldn * F, 2 * pl),
        ("C_hm = WM^T work   [40 x 4N, K=F]",
         lambda: k.gemm_view(WM, work, C_hm, R, 4 * ldn, F, transA=True),
         2.0 * R * 4 * ldn * F, 4 * pl + R * 4 * ldn * 4),
        ("SM_c = WMs HM      [F x N, K=40]",
         lambda: k.gemm_view(WM, HM, SM[:, :ldn], F, N, R),
         2.0 * F * N * R, pl + R * ldn * 4),
        ("D_q = plane HM^T   [F x 40, K=N] split-K",
         lambda: k.gemm_view(work[:, ldn:2 * ldn], HM, D, F, R, ldn, transB=True, workspace=ws),
         2.0 * F * R * ldn, pl + R * ldn * 4),
        ("tn = plane HPHI^T  [F x 4, K=N] split-K",
         lambda: k.gemm_view(work[:, :ldn], HPHI, tn, F, K, ldn, transB=True, workspace=ws),
         2.0 * F * K * ldn, pl + K * ldn * 4),
        ("D_q compact plane  [F x 40, K=N] split-K",
         lambda: k.gemm_view(plane, HM, D, F, R, ldn, transB=True, workspace=ws),
         2.0 * F * R * ldn, pl + R * ldn * 4),
        ("C_phi compact      [4 x N, K=F]",
         lambda: k.gemm_view(WPHI, plane, C_phi[:, :ldn], K, ldn, F, transA=True),
         2.0 * K * ldn * F, pl),
    ]
    only = [s for s in args.only.split(",") if s]
    print("PYFASST_GEMM_* =", {kk: v for kk, v in os.environ.items() if kk.startswith("PYFASST_GEMM")})
    for name, fn, flops, nbytes in shapes:
        if only and not any(o in name for o in only):
            continue
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print("%-42s %8.1f us  %7.1f TFLOP/s fp32-eq  %6.2f TB/s compulsory" %
              (name, ms * 1e3, flops / ms * 1e-9, nbytes / ms * 1e-9))


if __name__ == "__main__":
    main()
