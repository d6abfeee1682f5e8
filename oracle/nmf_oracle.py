"""CPU oracle for the IS-NMF initialisers.  TEST INFRASTRUCTURE ONLY.

float64 NumPy restatement of pyfasst/tools/nmf.py: `NMF_decomposition` (:24-62) and
`NMF_decomp_init` (:64-159), with the initial matrices as arguments (the reference draws them
from the global RNG).  Parity status: PINNED by tests/golden/nmf.npz, produced by running the
reference itself (oracle/make_golden.py: run_nmf) -- see tests/test_nmf_cpu.py.
"""
import numpy as np

EPS = 1e-10  # ref: nmf.py:22


def nmf_decomposition(SX, W0, H0, niter=10):
    """ref: nmf.py:33-62 (W0 is normalised here like :36)."""
    W, H = W0 / W0.sum(axis=0), H0.copy()
    for _ in range(niter):
        hat = W @ H
        W = W * ((SX / np.maximum(hat ** 2, EPS)) @ H.T) / np.maximum(
            (1 / np.maximum(hat, EPS)) @ H.T, EPS)
        s = W.sum(axis=0)
        s[s == 0] = 1.0
        W = W / s
        H = H * s[:, None]
        hat = W @ H
        H = H * (W.T @ (SX / np.maximum(hat ** 2, EPS))) / np.maximum(
            W.T @ (1 / np.maximum(hat, EPS)), EPS)
    return W, H


def nmf_decomp_init(SX, W0, H0, niter=10, updateW=True, updateH=True):
    """ref: nmf.py:108-159 (H0 is [nbComps, nframes])."""
    W, H = W0.copy(), H0.copy()
    if updateW:
        W = W / W.sum(axis=0)
    for _ in range(niter):
        if updateW:
            hat = W @ H
            W = W * ((SX / np.maximum(hat ** 2, EPS)) @ H.T) / np.maximum(
                (1 / np.maximum(hat, EPS)) @ H.T, EPS)
            s = W.sum(axis=0)
            s[s == 0] = 1.0
            W = W / s
            H = H * s[:, None]
        if updateH:
            hat = W @ H
            H = H * (W.T @ (SX / np.maximum(hat ** 2, EPS))) / np.maximum(
                W.T @ (1 / np.maximum(hat, EPS)), EPS)
    return W, H
