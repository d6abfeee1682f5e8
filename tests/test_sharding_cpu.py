"""Frequency-sharded GEM loop on two ranks (gloo, CPU, NumPy stand-in kernels): the
collectives of pyfasst_b200/engine.py (TW numerators/denominators, log-likelihood,
instantaneous-mixing statistics, spatial energies, FB column maxima) must reproduce the
single-process result."""
import copy
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_engine_cpu import engine_for, oracle_model, rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, conv, rank_sp, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyfasst_b200.engine import Comm
        m = oracle_model("mix_conv.wav" if conv else "mix_inst.wav", conv, rank_sp, 3)
        if not conv:
            m.spat_comps[1]["frdm_prior"] = "fixed"  # exercises the `oth` statistics
        eng = engine_for(m, "float64", comm=Comm())
        assert eng.F < m.nbFreqsSigRepr
        lls = eng.run(4)
        spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
        eng.read_model(spat, spec)
        psd = eng.noise_psd()
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), lls=lls, psd=psd,
                     **{"A%d" % j: spat[j]["params"] for j in spat},
                     **{"%s%d" % (nm, s): spec[s]["factor"][0][nm] for s in spec
                        for nm in ("FB", "FW", "TW")})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("conv,rank_sp", [(False, 2), (True, 1)])
def test_two_rank_frequency_sharding_matches_single_process(tmp_path, conv, rank_sp):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), conv, rank_sp, str(tmp_path)), nprocs=world,
             join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    m = oracle_model("mix_conv.wav" if conv else "mix_inst.wav", conv, rank_sp, 3)
    if not conv:
        m.spat_comps[1]["frdm_prior"] = "fixed"
    eng = engine_for(m, "float64")
    lls = eng.run(4)
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng.read_model(spat, spec)
    np.testing.assert_allclose(got["lls"], lls, rtol=1e-11)
    np.testing.assert_allclose(got["psd"], eng.noise_psd(), rtol=1e-13)
    for j in spat:
        assert rel_err(got["A%d" % j], spat[j]["params"]) < 1e-10
    for s in spec:
        for nm in ("FB", "FW", "TW"):
            assert rel_err(got["%s%d" % (nm, s)], spec[s]["factor"][0][nm]) < 1e-10
