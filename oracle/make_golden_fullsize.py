#!/usr/bin/env python
"""Test infrastructure only -- NOT part of the product path.

Golden vectors at BASELINE.json's REAL sizes, made by EXECUTING THE REFERENCE ITSELF (the
Python-2 sources under /root/reference through oracle/_py2shim.py) on the reference's own
audio file `data/tamy.wav` (L = 573 301 samples, stereo, 44.1 kHz; F = 1025, N = 1122):

  tamy_inst_r1   BASELINE configs[0]: MultiChanNMFInst_FASST, 3 sources, rank-1 spatial,
                 nbNMFComps = 4 (default), STFT 2048 / hop 512, 50 GEM iterations
                 (kwargs of pyfasst_tests/pyfasst/test_audioModel.py:13-22,
                 audioModel.py:330-382 estimation, :1063-1236 separation)
  tamy_conv_r2   MultiChanNMFConv + makeItConvolutive, 3 sources, rank-2 spatial, 2048 / 512,
                 50 GEM iterations

For each: the parameters after the constructor (the np.random.seed(0) draws), after ONE GEM
iteration and after all 50; logliks[50]; the E-step statistics of the initial parameters;
the separated signals written by separate_spat_comps (int16).  The 50-iteration runs take
minutes of CPU (the reference does about 5e5 TF bins x iterations per second), which is why
this is a separate script from make_golden.py:

    python oracle/make_golden_fullsize.py [inst|conv|all]

The GPU box never runs this; it reads tests/golden/tamy.wav (a copy of the reference's
CC BY-NC 3.0 fixture, see tests/golden/tamy.COPY) and the committed .npz files.
"""
import os
import shutil
import sys
import time
import warnings

import numpy as np
import scipy.io.wavfile as wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import _py2shim  # noqa: E402
from make_golden import snapshot  # noqa: E402

warnings.simplefilter("ignore")

#: frequency rows whose posterior powers hat_Ws[:, f, :] of the first E-step are kept
ROWS = (0, 1, 7, 64, 200, 511, 777, 1000, 1024)


def build(am, wav, conv, rank, iters):
    np.random.seed(0)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    model = cls(audio=wav, nbComps=3, spatial_rank=rank, iter_num=iters, verbose=0)
    if conv:
        model.makeItConvolutive()
    return model


def run(ref, name, wav, conv, rank, iters=50):
    am = ref["audioModel"]
    out = {}
    t0 = time.time()
    model = build(am, wav, conv, rank, 1)
    out["ann0"] = np.array(model.noise["ann_PSD_lim"][0])
    out["ann1"] = np.array(model.noise["ann_PSD_lim"][1])
    snapshot(model, "init", out)
    # one E-step on the initial parameters (audioModel.py:580-764)
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    scp, mm, rpi = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(scp, mm)
    out["e0_hat_Rxs"], out["e0_hat_Rss"] = hRxs, hRss
    out["e0_rows"] = np.array(ROWS)
    out["e0_hat_Ws_rows"] = np.array(hWs[:, list(ROWS), :])
    out["e0_loglik"] = np.real(ll)
    del scp, hWs
    ll1 = model.estim_param_a_post_model()
    snapshot(model, "it1", out)
    out["ll_it1"] = np.real(ll1)
    print(name, "one iteration done", time.time() - t0, flush=True)
    model = build(am, wav, conv, rank, iters)
    lls = model.estim_param_a_post_model()
    out["logliks"] = np.real(lls)
    snapshot(model, "final", out)
    out["noise_PSD_final"] = np.array(model.noise["PSD"])
    print(name, "logliks", out["logliks"], time.time() - t0, flush=True)
    outdir = "/tmp/pyfasst_golden_out_%s" % name
    os.makedirs(outdir, exist_ok=True)
    model.separate_spat_comps(dir_results=outdir)
    for n, f in enumerate(model.files["spat_comp"]):
        fs, y = wavfile.read(f)
        out["sep%d" % n] = y
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "done", time.time() - t0, flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ref = _py2shim.load()
    wav = os.path.join(GOLD, "tamy.wav")
    if not os.path.exists(wav):
        shutil.copyfile(os.path.join(_py2shim.REF, "data", "tamy.wav"), wav)
        shutil.copyfile(os.path.join(_py2shim.REF, "data", "COPY"), os.path.join(GOLD, "tamy.COPY"))
        os.chmod(wav, 0o644)
        os.chmod(os.path.join(GOLD, "tamy.COPY"), 0o644)
    if which in ("inst", "all"):
        run(ref, "tamy_inst_r1", wav, conv=False, rank=1)
    if which in ("conv", "all"):
        run(ref, "tamy_conv_r2", wav, conv=True, rank=2)


if __name__ == "__main__":
    main()
