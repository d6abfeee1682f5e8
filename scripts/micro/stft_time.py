#!/usr/bin/env python
"""Times the forward STFT kernel alone on the 10-min stereo mixture of configs[1] (int16 PCM
already in HBM, 2048 / 512, float32 planes): median of 5 launches between CUDA events and a
checksum of the planes (the variants must agree bit for bit).  PYFASST_B200_LIB / PYFASST_STFT_NB
select the build and the frames per batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyfasst_b200._lib import CudaKernels  # noqa: E402
from pyfasst_b200.tftransforms import stft as _stft  # noqa: E402

k = CudaKernels()
L = 600 * 44100
g = torch.Generator(device="cuda").manual_seed(3)
pcm = (torch.randn(L, 2, generator=g, device="cuda") * 3000).to(torch.int16)
win = np.sin(np.pi * (np.arange(2048) + 0.5) / 2048)
ts = []
for _ in range(6):
    psd = torch.zeros(1025, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    X, N = _stft.stft_planes(k, pcm, win, 512, 2048, "float32", psd, pcm_div=32768.0)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("lib=%s nb=%s: stft_planes %.3f ms (median of 5, incl. plane allocation + psd sum), checksum %.10e" %
      (os.environ.get("PYFASST_B200_LIB", "default"), os.environ.get("PYFASST_STFT_NB", "auto"),
       sorted(ts[1:])[2], float(X.double().abs().sum().item())))
