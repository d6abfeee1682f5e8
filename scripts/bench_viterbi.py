import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pyfasst_b200._lib import CudaKernels
k = CudaKernels()
for S, N in ((480, 20000), (480, 103362), (1092, 20000)):
    g = torch.Generator(device="cuda").manual_seed(0)
    dens = torch.randn((S, N), generator=g, device="cuda", dtype=torch.float64)
    prior = torch.zeros(S, device="cuda", dtype=torch.float64)
    trans = torch.log(torch.rand((S, S), generator=g, device="cuda", dtype=torch.float64) + 1e-3)
    k.viterbi(dens, prior, trans); torch.cuda.synchronize()
    t0 = time.perf_counter(); p = k.viterbi(dens, prior, trans); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("S=%d N=%d: %.1f ms (%.2f us per frame)" % (S, N, 1e3 * (t1 - t0), 1e6 * (t1 - t0) / N))
from oracle import build_ref
trk = build_ref.load()
if trk is not None:
    S, N = 480, 2000
    rng = np.random.default_rng(0)
    dens = rng.standard_normal((S, N)); prior = np.zeros(S); trans = np.log(rng.random((S, S)) + 1e-3)
    t0 = time.perf_counter(); trk.viterbiTracking(S, N, dens, prior, trans); t1 = time.perf_counter()
    print("reference Cython S=%d N=%d: %.1f ms (%.2f us per frame)" % (S, N, 1e3 * (t1 - t0), 1e6 * (t1 - t0) / N))
