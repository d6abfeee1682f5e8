"""Batches of independent mixtures on one GPU (BASELINE.json configs[4]: many short clips).

The reference has no batch API -- a user loops over files, one model per file.  A 30-s clip
(2.65 M TF bins) cannot fill a B200 and its GEM iteration is dominated by the ~70 small
per-iteration kernels, so the loop is restructured here: every model gets its own CUDA stream
and its GEM iteration is captured once as a CUDA graph; the graphs of all the clips are replayed
round-robin, so that the small kernels of one clip overlap the bandwidth-bound kernels of the
others.  Results are bit-identical to running `model.estim_param_a_post_model()` one model
after the other (same kernels, same order within a clip; nothing is shared between clips).

Multi-GPU: the clips are independent -- each rank takes `clips[rank::world]`, no collective.
"""
import numpy as np


def estimate_batch(models, use_cuda_graph=True):
    """Run `estim_param_a_post_model` of every model in `models` (all with the same
    `iter_num`) concurrently on the current device.  Returns the list of log-likelihood
    arrays; the models' `spat_comps` / `spec_comps` / `noise` are updated in place."""
    import torch
    if not models:
        return []
    n_iter = models[0].iter_num
    assert all(m.iter_num == n_iter for m in models), "all models must share iter_num"
    dev = models[0]._k().device
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in models]
    engines, logliks = [], []
    # pack: STFT (if not done) + parameters to HBM, each on its own stream
    for m, s in zip(models, streams):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            eng = m._engine()
            ll = torch.ones([max(n_iter, 1)], dtype=torch.float64, device=dev)
            eng.iter_dev.zero_()
            eng.flags.zero_()
            eng.totals.zero_()
        engines.append(eng)
        logliks.append(ll)
    if use_cuda_graph and n_iter > 1:
        # iteration 0 doubles as the warm-up run CUDA graph capture needs (the iteration counter
        # and the annealed noise PSD live in device memory, so one graph serves every iteration)
        for eng, ll, s in zip(engines, logliks, streams):
            with torch.cuda.stream(s):
                eng.gem_iteration(n_iter, ll)
        torch.cuda.synchronize(dev)
        # one graph per clip, replayed round-robin on the clips' streams.  (capture_begin /
        # capture_end directly: the `torch.cuda.graph` context manager synchronises, collects
        # garbage and EMPTIES the caching allocator on every entry -- 25 ms per clip, and every
        # later allocation of the batch, the separation's buffers included, then pays a
        # cudaMalloc.  Every graph keeps its own memory pool: they are replayed concurrently.
        # Measured alternatives, profiles/r02/batch_experiments.txt: ONE graph with a branch per
        # clip is slower, 1.65 s against 1.26 s for 32 clips x 100 iterations; so are 32 hardware
        # queues instead of the default 8 -- fewer clips in flight keep more of a clip L2 resident.)
        graphs = []
        for eng, ll, s in zip(engines, logliks, streams):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                g.capture_begin()
                try:
                    eng.gem_iteration(n_iter, ll)
                finally:
                    g.capture_end()
            graphs.append(g)
        for _ in range(n_iter - 1):
            for g, s in zip(graphs, streams):
                with torch.cuda.stream(s):
                    g.replay()
    else:
        for _ in range(n_iter):
            for eng, ll, s in zip(engines, logliks, streams):
                with torch.cuda.stream(s):
                    eng.gem_iteration(n_iter, ll)
    out = []
    for m, eng, ll, s in zip(models, engines, logliks, streams):
        with torch.cuda.stream(s):
            eng.n_iter_done = n_iter
            eng.check_flags()
            eng.read_model(m.spat_comps, m.spec_comps)
            m.noise['PSD'] = eng.noise_psd()
            out.append(ll[:n_iter].cpu().numpy())
        main.wait_stream(s)
    return out


def separate_batch(models, spec_comp_ind=None):
    """`separate_comps_pcm` of every model, each on its own stream.  Returns the list of int16
    arrays [nbSources, L, 2]."""
    import torch
    if not models:
        return []
    dev = models[0]._k().device
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in models]
    out = []
    for m, s in zip(models, streams):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            out.append(m.separate_comps_pcm(spec_comp_ind))
        main.wait_stream(s)
    return [np.asarray(o) for o in out]
