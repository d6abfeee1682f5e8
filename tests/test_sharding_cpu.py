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


def _worker_time(rank, world, port, conv, rank_sp, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pyfasst_b200.engine import Comm, GemEngine
        from tests.fake_kernels import FakeKernels
        m = oracle_model("mix_conv.wav" if conv else "mix_inst.wav", conv, rank_sp, 3)
        if not conv:
            m.spat_comps[1]["frdm_prior"] = "fixed"
        eng = GemEngine(FakeKernels(), m.nbFreqsSigRepr, m.nbFramesSigRepr, dtype="float64",
                        comm=Comm(), shard="time")
        assert eng.N < m.nbFramesSigRepr and eng.F == m.nbFreqsSigRepr
        eng.set_X_host(m.X)
        lim = m.noise["ann_PSD_lim"]
        eng.set_noise(m.noise["sim_ann_opt"], lim[0], lim[1], m.noise["PSD"])
        eng.set_model(m.spat_comps, m.spec_comps, m.nmfUpdateCoeff)
        lls = eng.run(4)
        spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
        eng.read_model(spat, spec)
        if rank == 0:
            np.savez(os.path.join(out_dir, "sharded.npz"), lls=lls, psd=eng.noise_psd(),
                     **{"A%d" % j: spat[j]["params"] for j in spat},
                     **{"%s%d" % (nm, s): spec[s]["factor"][0][nm] for s in spec
                        for nm in ("FB", "FW", "TW")})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("conv,rank_sp", [(False, 2), (True, 2)])
def test_three_rank_frame_sharding_matches_single_process(tmp_path, conv, rank_sp):
    """shard='time': only the per-frequency E-step statistics and the FB numerators /
    denominators cross the ranks."""
    world = 3
    mp.spawn(_worker_time, args=(world, _free_port(), conv, rank_sp, str(tmp_path)),
             nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    m = oracle_model("mix_conv.wav" if conv else "mix_inst.wav", conv, rank_sp, 3)
    if not conv:
        m.spat_comps[1]["frdm_prior"] = "fixed"
    eng = engine_for(m, "float64")
    lls = eng.run(4)
    spat, spec = copy.deepcopy(m.spat_comps), copy.deepcopy(m.spec_comps)
    eng.read_model(spat, spec)
    np.testing.assert_allclose(got["lls"], lls, rtol=1e-11)
    for j in spat:
        assert rel_err(got["A%d" % j], spat[j]["params"]) < 1e-10
    for s in spec:
        for nm in ("FB", "FW", "TW"):
            assert rel_err(got["%s%d" % (nm, s)], spec[s]["factor"][0][nm]) < 1e-10


def _worker_api(rank, world, port, shard, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pyfasst_b200.audioModel as am
        from pyfasst_b200.engine import Comm
        from tests.fake_kernels import FakeKernels
        from tests.test_engine_cpu import GOLDEN
        np.random.seed(0)
        model = am.MultiChanNMFConv(audio=os.path.join(GOLDEN, "mix_conv.wav"), nbComps=2,
                                    nbNMFComps=4, spatial_rank=2, wlen=256, hopsize=64,
                                    iter_num=6, ann_PSD_lim=[None, None],
                                    compute_dtype="float64", kernels=FakeKernels(), comm=Comm(),
                                    shard=shard)
        model.makeItConvolutive()
        lls = model.estim_param_a_post_model()
        pcm = model.separate_comps_pcm({j: [j] for j in range(2)})  # (on the rank-local shards)
        model.gather_parameters()
        if rank == 0:
            np.savez(os.path.join(out_dir, "api.npz"), lls=lls, pcm=pcm,
                     FB0=model.spec_comps[0]["factor"][0]["FB"],
                     TW1=model.spec_comps[1]["factor"][0]["TW"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["time", "freq"])
def test_sharded_public_api_matches_reference_golden(tmp_path, shard):
    """The drop-in classes on 2 ranks (STFT of the rank's own frames / rows, sharded GEM,
    gathered Wiener + iSTFT) against the golden vectors produced by the reference."""
    from tests.test_engine_cpu import GOLDEN
    mp.spawn(_worker_api, args=(2, _free_port(), shard, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "api.npz"))
    g = np.load(os.path.join(GOLDEN, "fasst_conv_r2.npz"))
    np.testing.assert_allclose(got["lls"], g["logliks"], rtol=1e-9)
    assert rel_err(got["FB0"], g["final_FB0"]) < 1e-7
    assert rel_err(got["TW1"], g["final_TW1"]) < 1e-7
    for n in range(2):
        diff = np.abs(got["pcm"][n].astype(int) - g["sep%d" % n].astype(int))
        assert diff.max() <= 1
