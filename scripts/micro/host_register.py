"""Micro-benchmark: pageable vs cudaHostRegister-ed NumPy arrays for the parameter upload / download
(4 x 13 MB float64 = the TW matrices of configs[1])."""
import time
import numpy as np
import torch

rt = torch.cuda.cudart()
arrs = [np.random.rand(32, 51682) for _ in range(4)]
dev = [torch.empty((32, 51682), dtype=torch.float64, device="cuda") for _ in range(4)]
torch.cuda.synchronize()


def h2d():
    for a, d in zip(arrs, dev):
        d.copy_(torch.from_numpy(a))
    torch.cuda.synchronize()


def d2h():
    for a, d in zip(arrs, dev):
        torch.from_numpy(a).copy_(d)
    torch.cuda.synchronize()


def t(fn, n=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return 1e3 * (time.perf_counter() - t0) / n


print("pageable  : H2D %.2f ms  D2H %.2f ms" % (t(h2d), t(d2h)))
t0 = time.perf_counter()
for a in arrs:
    r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
t1 = time.perf_counter()
print("register 4 x %.1f MB: %.2f ms (rc %s)" % (arrs[0].nbytes / 1e6, 1e3 * (t1 - t0), r))
print("registered: H2D %.2f ms  D2H %.2f ms" % (t(h2d), t(d2h)))
t0 = time.perf_counter()
for a in arrs:
    rt.cudaHostUnregister(a.ctypes.data)
print("unregister: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
