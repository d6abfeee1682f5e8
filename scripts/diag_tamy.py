"""Diagnostic: per-parameter error of the CUDA path against the reference's golden vectors on
tamy.wav after one GEM iteration and over 50 (which parameter carries the float32 error)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyfasst_b200.audioModel as am  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def build(conv, rank, dtype, iters):
    np.random.seed(0)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    m = cls(audio=os.path.join(GOLDEN, "tamy.wav"), nbComps=3, spatial_rank=rank, iter_num=iters,
            verbose=0, compute_dtype=dtype)
    if conv:
        m.makeItConvolutive()
    return m


def main():
    dtype = sys.argv[1] if len(sys.argv) > 1 else "float32"
    for name, conv, rank in (("tamy_conv_r2", True, 2), ("tamy_inst_r1", False, 1)):
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        m = build(conv, rank, dtype, 1)
        m.noise["PSD"] = m.noise["ann_PSD_lim"][0]
        powers, mix, ranks = m.retrieve_subsrc_params()
        hRxx, hRxs, hRss, hWs, ll = m.compute_suff_stat(powers, mix)
        sc = np.abs(g["e0_hat_Rss"]).max(axis=(1, 2), keepdims=True)
        e_rss = np.abs(hRss - g["e0_hat_Rss"]) / sc
        sc = np.abs(g["e0_hat_Rxs"]).max(axis=(1, 2), keepdims=True)
        e_rxs = np.abs(hRxs - g["e0_hat_Rxs"]) / sc
        rows = list(g["e0_rows"])
        e_w = np.abs(hWs[:, rows, :] / g["e0_hat_Ws_rows"] - 1)
        print("%s %s E-step: Rss max %.2g rms %.2g | Rxs max %.2g rms %.2g | hatW max %.2g rms %.2g | ll %.2g"
              % (name, dtype, e_rss.max(), np.sqrt((e_rss ** 2).mean()), e_rxs.max(),
                 np.sqrt((e_rxs ** 2).mean()), e_w.max(), np.sqrt((e_w ** 2).mean()),
                 abs(ll / g["e0_loglik"] - 1)))
        lls = m.estim_param_a_post_model()
        line = "  it1: ll %.2g" % abs(lls[0] / g["ll_it1"][0] - 1)
        for j in range(3):
            line += " | A%d %.2g" % (j, rel_err(m.spat_comps[j]["params"], g["it1_A%d" % j]))
            for nm in ("FB", "TW"):
                line += " %s %.2g" % (nm, rel_err(m.spec_comps[j]["factor"][0][nm],
                                                  g["it1_%s%d" % (nm, j)]))
        print(line)
        m = build(conv, rank, dtype, 50)
        lls = m.estim_param_a_post_model()
        e = np.abs(lls - g["logliks"]) / np.abs(g["logliks"])
        print("  50 iterations: ll rel err at 1,2,5,10,20,30,40,50: " +
              " ".join("%.2g" % e[i - 1] for i in (1, 2, 5, 10, 20, 30, 40, 50)) + " max %.2g" % e.max())
        line = "  final:"
        for j in range(3):
            line += " | A%d %.2g" % (j, rel_err(m.spat_comps[j]["params"], g["final_A%d" % j]))
            for nm in ("FB", "TW"):
                line += " %s %.2g" % (nm, rel_err(m.spec_comps[j]["factor"][0][nm],
                                                  g["final_%s%d" % (nm, j)]))
        print(line)


if __name__ == "__main__":
    main()
