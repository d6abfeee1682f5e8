#!/usr/bin/env python
"""bench.py -- GEM throughput of the FASST hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one GEM iteration (E-step + spatial / spectral M-step + renormalisation,
ref pyfasst/audioModel.py:384-428) over the whole synthetic mixture.  Workload at N=1 =
BASELINE.json configs[1]: synthetic 10-min stereo 44.1 kHz mixture, STFT 2048 / hop 512
(F=1025, N=51682, 52.97 M TF bins), MultiChanNMFInst_FASST with 4 sources x K=32 NMF
components, full-rank (rank 2) spatial model.  N>1: WEAK scaling -- one mixture of N x 10
minutes (80 min at N=8, the size class of BASELINE configs[3]) whose frames are sharded over
the GPUs (one process per GPU, NCCL all-reduce of the per-frequency statistics).

Printed JSON (one line, rank 0):
  value     TF-bins*iterations/s with X and the parameters resident in HBM (CUDA events)
  e2e       the same metric through the public API with host buffers:
            MultiChanNMFInst_FASST.comp_transf_Cx() (PCM host->device + STFT) +
            estim_param_a_post_model() (parameters host->device, K iterations,
            parameters + log-likelihoods device->host), wall clock around synchronised calls
  roofline  fused E-step kernel: algorithmic bytes 4*(I^2+2J) per bin / its CUDA-event time
  cpu_baseline  the oracle (NumPy restatement of the reference, float64) on a bounded crop
`--impl reference` times that oracle instead (the reference is Python 2 and cannot run
here) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 44100
WLEN, HOP = 2048, 512
NSRC, NNMF, RANK = 4, 32, 2
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused E-step kernel on the
# default workload, from the committed `ncu --set full` capture
# profiles/r02/ncu_estep_stereo_kernel_inst.txt (1.696543 GB + 0.821810 GB); the algorithmic figure
# is 48 B x 52,974,050 bins = 2.543 GB, i.e. no wasted re-reads.
ESTEP_DRAM_BYTES_PER_LAUNCH = 2.518353e9
METRIC = "gem_tf_bins_iters_per_s"
UNIT = "TF-bins*iters/s"


# --------------------------------------------------------------------------- workload
def synth_mix(duration_s, seed=1234, fs=FS, nsrc=NSRC, channels=2):
    """int16 mixture: AR(2)-coloured Gaussian sources with random on/off envelopes, white noise
    at -40 dB.  Stereo: instantaneous mixing; more channels (configs[3]): every
    source reaches every channel through its own random 64-tap FIR filter."""
    import scipy.signal as sps
    rng = np.random.default_rng(seed)
    L = int(round(duration_s * fs))
    mix = np.zeros((L, channels))
    for j in range(nsrc):
        r, th = 0.97 - 0.02 * j, np.pi * (0.05 + 0.11 * j)
        s = sps.lfilter([1.0], [1.0, -2 * r * np.cos(th), r * r], rng.standard_normal(L))
        # random on/off envelope, segments of 0.1 - 1 s
        nseg = int(duration_s / 0.1) + 2
        lens = rng.integers(int(0.1 * fs), int(1.0 * fs), nseg)
        state = rng.random(nseg) < 0.6
        env = np.repeat(state, lens)[:L].astype(np.float64)
        if env.size < L:
            env = np.pad(env, (0, L - env.size))
        s = s * env / (np.abs(s).max() + 1e-12)
        if channels == 2:
            ang = (j + 1) * np.pi / (2.0 * (nsrc + 1))
            mix[:, 0] += np.sin(ang) * s
            mix[:, 1] += np.cos(ang) * s
        else:
            for c in range(channels):
                h = rng.standard_normal(64) * np.exp(-np.arange(64) / 12.0)
                mix[:, c] += sps.fftconvolve(s, h)[:L]
    mix += 10 ** (-40 / 20.0) * np.abs(mix).max() * rng.standard_normal(mix.shape)
    mix = 0.9 * mix / np.abs(mix).max()
    return np.int16(np.round(mix * 32767))


def write_wav(path, pcm):
    import scipy.io.wavfile as wavfile
    wavfile.write(path, FS, pcm)


def n_frames(L):
    return int(np.ceil(L / float(HOP)) + 2)


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region, polled through NVML
    (nvidia-ml-py) every ~12 ms from a thread; nvidia-smi is the fallback."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
               "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}

    PERIOD_S = 0.012

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.query_s = []
        self.reasons = set()
        self.sm_max = None
        self._stop = threading.Event()
        self.thread = None
        self.error = None

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA index through the UUID-free
            # common case (CUDA_VISIBLE_DEVICES unset or a prefix of the device list)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    idx = int(ids[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                t0 = time.perf_counter()
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.query_s.append(time.perf_counter() - t0)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                # (a 40-ms timed region gets 3-4 samples; polling faster buys nothing and every
                # NVML query goes through the driver the launching thread is also calling into)
                self._stop.wait(self.PERIOD_S)
        except Exception as exc:  # noqa: BLE001 -- reported in the JSON line
            self.error = repr(exc)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        # NVML initialisation takes a few ms: do not let a short timed region end before the
        # first sample is in
        t0 = time.time()
        while not self.samples and self.error is None and time.time() - t0 < 2.0:
            time.sleep(0.005)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=5)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
               "sm_max_mhz": self.sm_max, "samples": len(self.samples),
               "reasons": sorted(self.reasons),
               "nvml_query_ms": round(1e3 * float(np.median(self.query_s)), 3) if self.query_s else None,
               "how": "NVML polled every %d ms during the timed region" % round(1e3 * self.PERIOD_S)}
        if self.error:
            out["error"] = self.error
        return out


# --------------------------------------------------------------------------- CPU arm
CPU_CROP_S = 8.0   # crop of the synthetic mixture the CPU arms run on (both of them)
CPU_ITERS, CPU_WARM = 5, 1  # in-line cpu_baseline; `--impl reference` takes --steps / --warmup


def oracle_model(pcm, iters, nsrc=NSRC, nnmf=NNMF, rank=RANK):
    from oracle import fasst_oracle as fo
    maxdata = np.maximum(1.1 * np.abs(pcm).max(), 1e-10)
    np.random.seed(0)
    return fo.OracleFASST((FS, pcm / maxdata, maxdata), nbComps=nsrc, nbNMFComps=nnmf,
                          spatial_rank=rank, wlen=WLEN, hopsize=HOP, iter_num=iters)


def cpu_gem_rate(crop_s, iters, warm=0, workload="configs1"):
    """TF-bins*iters/s of the oracle (reference algorithm, NumPy float64) on a crop."""
    if workload == "tamy":  # configs[0] is the reference's own CPU-runnable case: the whole file
        m = oracle_model(tamy_pcm(), iters + warm, 3, 4, 1)
    else:
        m = oracle_model(synth_mix(crop_s), iters + warm)
    bins = m.nbFreqsSigRepr * m.nbFramesSigRepr
    for _ in range(warm):
        m.GEM_iteration()
    t0 = time.perf_counter()
    for _ in range(iters):
        m.GEM_iteration()
    dt = time.perf_counter() - t0
    return bins * iters / dt, bins, dt


def host_threads():
    """Gives the BLAS behind NumPy every host core (torchrun exports OMP_NUM_THREADS=1 to its
    workers) and returns the number of threads it will actually use -- the `cores` of the CPU
    arm."""
    want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        return max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])
    except Exception:  # noqa: BLE001 -- without threadpoolctl the environment decides
        return int(os.environ.get("OMP_NUM_THREADS", want))


def run_reference(args, rank):
    if rank != 0:
        return
    crop_s = args.cpu_crop_s
    cores = host_threads()
    rate, bins, dt = cpu_gem_rate(crop_s, args.steps, args.warmup, args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": data_kind(args),
        "config": workload_config(args, note="CPU arm: the oracle (NumPy restatement of the "
                                  "reference's GEM_iteration, float64) on %s (%d TF bins per "
                                  "step); the reference itself is Python 2 and cannot be run" %
                                  ("the whole file" if args.workload == "tamy" else
                                   "a %.1f s crop of the same synthetic mixture" % crop_s, bins)),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d GEM iterations on a %.1f s crop (%d bins), NumPy "
                                   "float64, BLAS threads = %d" % (args.steps, crop_s, bins, cores)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, note=None):
    if getattr(args, "workload", "configs1") == "tamy":
        L = tamy_pcm().shape[0]
        cfg = {"workload": "configs[0]: data/tamy.wav (the reference's fixture, %.1f s stereo 44.1 kHz), "
                           "MultiChanNMFInst_FASST 3 sources x K=4, spatial rank 1, STFT %d/hop %d; "
                           "one step = one GEM iteration" % (L / float(FS), WLEN, HOP),
               "channels": 2, "F": WLEN // 2 + 1, "N": n_frames(L),
               "tf_bins": (WLEN // 2 + 1) * n_frames(L), "sources": 3, "nmf_comps": 4,
               "spatial_rank": 1, "sharding": "none (one GPU)",
               "l2": "the whole problem (14 planes x 4.6 MB) is L2 resident: an iteration of this "
                     "size is bound by its ~60 launches, not by HBM"}
        if note:
            cfg["note"] = note
        return cfg
    blocks = getattr(args, "blocks", 1)
    L = int(round(args.duration_s * FS)) * args.gpus * blocks
    model = "MultiChanNMFConv (convolutive mixing)" if getattr(args, "model", "inst") == "conv" \
        else "MultiChanNMFInst_FASST"
    shard = getattr(args, "shard", "time")
    nch, rank = getattr(args, "channels", 2), getattr(args, "rank", RANK)
    cfg = {"workload": "%s: synthetic %.0f-s %s 44.1 kHz mix per GPU (x%d GPUs = %.0f s, "
                       "weak scaling), %s %d sources x K=%d, spatial rank %d, "
                       "STFT %d/hop %d" % ("configs[1]" if nch == 2 else "configs[3]-like",
                                           args.duration_s * blocks,
                                           "stereo" if nch == 2 else "%d-channel" % nch,
                                           args.gpus, args.duration_s * args.gpus * blocks,
                                           model, NSRC, NNMF, rank, WLEN, HOP),
           "channels": nch,
           "F": WLEN // 2 + 1, "N": n_frames(L), "tf_bins": (WLEN // 2 + 1) * n_frames(L),
           "sources": NSRC, "nmf_comps": NNMF, "spatial_rank": rank,
           "sharding": ("frames over %d GPU(s); all-reduce of the per-frequency E-step statistics "
                        "and of the FB numerators/denominators (NCCL)" % args.gpus)
           if shard == "time" else
           ("frequency bins over %d GPU(s); all-reduce of the TW numerators/denominators, the "
            "log-likelihood and the renormalisation maxima (NCCL)" % args.gpus),
           "l2": "inputs larger than L2 (X + V + hat_W planes >> 126 MB), no flush needed"}
    if note:
        cfg["note"] = note
    return cfg


def data_kind(args):
    return "tests/golden/tamy.wav (copy of the reference's data/tamy.wav)" \
        if args.workload == "tamy" else "synthetic"


def tamy_pcm():
    """int16 PCM of the reference's own fixture (committed copy: tests/golden/tamy.wav)."""
    from scipy.io import wavfile
    fs, pcm = wavfile.read(os.path.join(ROOT, "tests", "golden", "tamy.wav"))
    assert fs == FS and pcm.ndim == 2 and pcm.shape[1] == 2
    return np.ascontiguousarray(pcm)


# --------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    import pyfasst_b200.audioModel as am
    from pyfasst_b200.engine import Comm

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        # (NCCL's own stream at high priority: the collectives of the frequency partition overlap
        # the contraction kernels of the next component, which fill the GPU)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
        comm = Comm()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N = 1: the 10-min mixture of configs[1].  N > 1: weak scaling -- the mixture is N such
    # blocks back to back (one per GPU) and its frames are sharded over the ranks.
    tamy = args.workload == "tamy"
    nsrc, nnmf = (3, 4) if tamy else (NSRC, NNMF)
    block = tamy_pcm() if tamy else synth_mix(args.duration_s, channels=args.channels)
    # (--blocks B: B such blocks per GPU, e.g. --duration-s 450 --blocks 8 on one GPU is the
    # same 1-hour mixture as --duration-s 450 on 8 GPUs: strong-scaling pairs)
    reps = world * args.blocks
    pcm = block if reps == 1 else np.ascontiguousarray(np.tile(block, (reps, 1)))
    L = pcm.shape[0]
    F, N = WLEN // 2 + 1, n_frames(L)
    bins = F * N
    dtype = "float32" if args.dtype == "f32" else "float64"
    import pyfasst_b200.audioObject as ao

    def make_audio(raw):
        a = ao.AudioObject("synthetic_mix.wav")
        a._samplerate = FS
        a._set_raw(raw)  # the scaling factor is scanned on the device by the model
        return a

    def make_model(iters, raw=None):
        np.random.seed(0)
        cls = am.MultiChanNMFConv if args.model == "conv" else am.MultiChanNMFInst_FASST
        m = cls(audio=make_audio(pcm if raw is None else raw), nbComps=nsrc, nbNMFComps=nnmf,
                spatial_rank=args.rank, wlen=WLEN, hopsize=HOP, iter_num=iters,
                ann_PSD_lim=[None, None], compute_dtype=dtype, comm=comm, shard=args.shard)
        if args.model == "conv":
            m.makeItConvolutive()
        return m

    # ---- device-resident throughput (`value`) and per-phase / E-step kernel times -------
    model = make_model(args.steps + args.warmup)
    kern = model._k()
    eng = model._engine()
    total_iters = args.steps + args.warmup
    logliks = torch.ones(total_iters, dtype=torch.float64, device=eng.dev)
    eng.iter_dev.zero_()
    eng.flags.zero_()
    eng.totals.zero_()
    for _ in range(args.warmup):
        eng.gem_iteration(total_iters, logliks)
    events = []

    def mark(label):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        events.append((label, ev))

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = kern.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    kern.estep_timing(True)  # an event pair around the per-bin E-step kernel alone, on its stream
    t_start.record()
    for _ in range(args.steps):
        eng.gem_iteration(total_iters, logliks, mark)
    t_end.record()
    barrier()
    kern.estep_timing(False)
    estep_kernel_ms, estep_launches = kern.estep_timing_read()
    estep_kernel_ms /= max(estep_launches, 1)
    clocks = sampler.stop()
    launches = kern.launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    eng.check_flags()
    phases = {}
    for (l0, e0), (l1, e1) in zip(events[:-1], events[1:]):
        if l1 != "begin":
            phases[l1] = phases.get(l1, 0.0) + e0.elapsed_time(e1)
    phases = {k: v / args.steps for k, v in phases.items()}
    ms_t = torch.tensor([ms], dtype=torch.float64, device=eng.dev)
    est = torch.tensor([estep_kernel_ms], dtype=torch.float64, device=eng.dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item())
    estep_ms = float(est.item())
    value = bins * args.steps / (ms * 1e-3)
    final_ll = logliks.cpu().numpy()
    del eng, model
    torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers (`e2e`) ---------------------
    # the mixture as a host buffer (int16 PCM)
    pinned = None if args.no_e2e else torch.from_numpy(pcm).pin_memory()

    def api_run(iters):
        m = make_model(iters)  # constructor = reference behaviour (reads the WAV, STFT, init)
        stages = {}
        barrier()
        t0 = time.perf_counter()
        m.audioObject._set_raw(pinned)  # host PCM -> AudioObject (no host pass over the samples)
        m.comp_transf_Cx()     # PCM host->device + STFT kernels (+ annealing limits D2H)
        torch.cuda.synchronize()
        stages["comp_transf_Cx_s"] = time.perf_counter() - t0
        ll = m.estim_param_a_post_model()  # params H2D, GEM iterations, params + LL D2H
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        stages["estim_param_a_post_model_s"] = t1 - t0 - stages["comp_transf_Cx_s"]
        for key in ("pack_s", "gem_s", "unpack_s"):
            stages["estim." + key] = m._last_engine_stats[key]
        npar = sum(np.asarray(sc["params"]).nbytes for sc in m.spat_comps.values()) + \
            sum(f["FB"].nbytes + f["FW"].nbytes + f["TW"].nbytes
                for sp in m.spec_comps.values() for f in sp["factor"].values())
        return t1 - t0, npar, ll, stages

    e2e = None
    if not args.no_e2e:  # (scaling-evidence runs may skip it; the driver's runs never do)
        for _ in range(2):  # warm-up: library state and the caching allocator's block sizes
            api_run(max(1, args.warmup))
        dt, npar, ll_api, stages = api_run(args.steps)
        dt_t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t.item())
        h2d = (pcm.nbytes + npar) / float(args.steps)
        d2h = (npar + 8 * args.steps + 8 * F) / float(args.steps)
        e2e = {"value": bins * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "wall_s": dt, "stages": stages,
               "what": "AudioObject._set_raw(pinned int16 PCM) + comp_transf_Cx() + "
                       "estim_param_a_post_model() with iter_num=steps; host numpy/PCM in, host "
                       "numpy parameters and log-likelihoods out"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused E-step) -----------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sz = 4 if args.dtype == "f32" else 8
    # 2 I reals of x (Cx = x x^H is rank one; = I^2 at I = 2) + V_j + hat_W_j
    bytes_per_bin = sz * (2 * args.channels + 2 * nsrc)
    local_bins = bins / float(world)  # frames are split evenly over the ranks
    achieved = bytes_per_bin * local_bins / (estep_ms * 1e-3) / 1e9
    traffic = args.traffic
    if traffic is None and world == 1 and args.dtype == "f32" and args.duration_s == 600.0 and args.blocks == 1 \
            and args.channels == 2 and args.rank == RANK and args.model == "inst" \
            and args.workload == "configs1":
        traffic = ESTEP_DRAM_BYTES_PER_LAUNCH  # same workload as the committed capture
    roofline = {"bound": "hbm",
                "kernel": "estep_stereo_kernel" if args.channels == 2 else "estep_multi_kernel",
                "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "bytes_per_bin": bytes_per_bin, "ms_per_launch": estep_ms,
                "how": "CUDA events recorded by the library on the launching stream right before and "
                       "after the per-bin kernel, every launch of the timed region (average); "
                       "phases_ms.estep also holds the two small per-frequency kernels around it",
                "algorithmic_bytes_per_launch": bytes_per_bin * local_bins,
                "traffic": traffic}

    # ---- CPU baseline: the oracle on a bounded crop -----------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.channels == 2:
        cores = host_threads()
        # the same sample as `--impl reference`: the same crop, warm iterations first
        rate, cbins, cdt = cpu_gem_rate(args.cpu_crop_s, CPU_ITERS, CPU_WARM, args.workload)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d GEM iterations (after %d warm) of the oracle (NumPy float64 "
                         "restatement of the reference) on a %.1f s crop (%d bins, %.1f s), "
                         "BLAS threads = %d" %
                         (CPU_ITERS, CPU_WARM, args.cpu_crop_s, cbins, cdt, cores)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": data_kind(args),
        "config": workload_config(args), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": phases, "loglik_last": float(final_ll[total_iters - 1]),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--duration-s", type=float, default=600.0)
    ap.add_argument("--cpu-crop-s", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the end-to-end leg (scaling-evidence runs only)")
    ap.add_argument("--blocks", type=int, default=1,
                    help="blocks of --duration-s per GPU (strong-scaling pairs: N GPUs x B "
                         "blocks = the same mixture)")
    ap.add_argument("--model", default="inst", choices=["inst", "conv"],
                    help="inst: MultiChanNMFInst_FASST (configs[1], default); conv: "
                         "MultiChanNMFConv + makeItConvolutive (the model of configs[3])")
    ap.add_argument("--channels", type=int, default=2, choices=[2, 3, 4],
                    help="channels of the synthetic mixture (default 2 = configs[1]; 4 with "
                         "--model conv --rank 4 = the model of configs[3]; the CPU arms are stereo)")
    ap.add_argument("--rank", type=int, default=RANK, help="spatial rank of every source")
    ap.add_argument("--shard", default="time", choices=["time", "freq"],
                    help="multi-GPU partition: frames (default) or frequency bins")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "tamy", "simm"],
                    help="configs1 (default): BASELINE configs[1], the 10-min synthetic stereo "
                         "mixture; tamy: configs[0], the reference's data/tamy.wav, 3 sources, "
                         "rank 1 (one GPU); simm: configs[2], Stereo_SIMM with the 480-pitch "
                         "dictionary (scripts/bench_simm.py, one GPU)")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per E-step launch from the committed ncu capture")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.workload == "simm":  # configs[2]: its own script, the same JSON schema
        if rank == 0:
            sys.argv = [os.path.join(ROOT, "scripts", "bench_simm.py"), "--steps", str(args.steps),
                        "--warmup", str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import bench_simm
            bench_simm.main()
        return
    if args.workload == "tamy":
        args.rank = 1
        if world != 1:
            sys.stderr.write("bench.py: --workload tamy runs on one GPU (rank 0)\n")
            if rank != 0:
                return
            world = args.gpus = 1
    if args.cpu_crop_s is None:
        args.cpu_crop_s = CPU_CROP_S  # one crop for both CPU arms
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE\n"
                         % (args.gpus, world))
        args.gpus = world
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
