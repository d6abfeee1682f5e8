"""Device-resident GEM engine: orchestrates the sm_100a kernels of one FASST model.

This is the host side of the hot path: it mirrors the control flow of
FASST.estim_param_a_post_model / GEM_iteration / update_spectral_components /
renormalize_parameters (ref: pyfasst/audioModel.py:330-428, :1469-1727, :1980-2040)
but every array lives in HBM and every numerical step is a kernel of the C ABI
(include/pyfasst_b200.h).  One iteration enqueues ~40 kernels on the current
stream and never synchronises with the host (the iteration counter, the annealed
noise PSD and the log-likelihood trace live on the device), so the whole loop can
be captured in a CUDA graph.

Multi-GPU (one process per GPU, `Comm` over torch.distributed), two sharding modes:

* shard='freq' (what BASELINE.json prescribes): each rank owns the rows [f_lo, f_hi) of X, V,
  hat_W, FB, A and the noise PSD; TW (and FW) are replicated.  The E-step, the FB update and
  the convolutive mixing update are local; per iteration the ranks all-reduce the TW
  numerators/denominators [S,2,K,N] (SUM), the scalar log-likelihood, the instantaneous-mixing
  statistics, the spatial energies (SUM) and the FB column maxima (MAX).
* shard='time': each rank owns the frames [n_lo, n_hi) of X, V, hat_W and TW; FB, FW, A and the
  noise PSD are replicated.  The TW update is local; per iteration the ranks all-reduce the
  per-frequency E-step statistics (hat_Rss, hat_Rxs, log-likelihood: O(F R^2)) and the FB
  numerators/denominators [S,2,F,K] -- a few MB whatever the length of the mixture, where the
  frequency mode moves O(K N) per iteration.  This is the mode bench.py scales with.

Supported model structure (what MultiChanNMFInst_FASST / MultiChanNMFConv build,
audioModel.py:2349-2393, :2488-2508): stereo; one spectral component per spatial
component, one factor each with FB [F,Kb], FW [Kb,Kw] (fixed), TW [Kw,N] 'NMF'
constrained, TB empty; FB / TW / mixing parameters individually 'free' or 'fixed';
all spatial components instantaneous, or all convolutive and free (Q7).  Anything
else raises NotImplementedError -- there is no CPU fallback.
"""
import os

import numpy as np

EPS = 1e-10  # ref: audioModel.py:61


class Comm(object):
    """Collectives of the frequency-sharded GEM loop over torch.distributed."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def allreduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)

    def allgather(self, t):
        out = [t.new_empty(t.shape) for _ in range(self.world)]
        self.dist.all_gather(out, t.contiguous(), group=self.group)
        return out

    def reduce_scatter_sum(self, out, inp):
        """out = this rank's chunk of the sum over ranks of inp ([world, *out.shape], contiguous)."""
        if self.dist.get_backend(self.group) == "gloo":  # (CPU tests: gloo has no reduce-scatter)
            tmp = inp.clone()
            self.dist.all_reduce(tmp, op=self.dist.ReduceOp.SUM, group=self.group)
            out.copy_(tmp[self.rank])
        else:
            self.dist.reduce_scatter_tensor(out, inp, op=self.dist.ReduceOp.SUM, group=self.group)

    def allgather_into(self, out, inp):
        """out [world, *inp.shape] = the ranks' inp, in rank order."""
        if self.dist.get_backend(self.group) == "gloo":
            parts = [out[r] for r in range(self.world)]
            self.dist.all_gather(parts, inp.contiguous(), group=self.group)
        else:
            self.dist.all_gather_into_tensor(out, inp, group=self.group)


def shard_bounds(F, world):
    """Contiguous frequency shards, sizes differing by at most one (1025 = 129 + 7*128)."""
    base, extra = divmod(F, world)
    lo = [r * base + min(r, extra) for r in range(world)]
    return [(lo[r], lo[r] + base + (1 if r < extra else 0)) for r in range(world)]


def _round_up(n, m):
    return (n + m - 1) // m * m


class GemEngine(object):
    def __init__(self, kernels, F_total, N, dtype="float32", comm=None, f_range=None,
                 shard="freq", n_range=None):
        import torch
        self.torch = torch
        self.k = kernels
        self.dev = kernels.device
        self.tdtype = {"float32": torch.float32, "float64": torch.float64}[dtype]
        self.F_total = int(F_total)
        self.N_total = int(N)
        self.comm = comm
        if shard not in ("freq", "time"):
            raise ValueError("shard must be 'freq' or 'time'")
        self.shard = shard
        multi = comm is not None and comm.world > 1
        if f_range is None:
            f_range = shard_bounds(self.F_total, comm.world)[comm.rank] \
                if (multi and shard == "freq") else (0, self.F_total)
        if n_range is None:
            n_range = shard_bounds(self.N_total, comm.world)[comm.rank] \
                if (multi and shard == "time") else (0, self.N_total)
        self.f_lo, self.f_hi = f_range
        self.n_lo, self.n_hi = n_range
        self.F = self.f_hi - self.f_lo
        self.N = self.n_hi - self.n_lo   # local frames
        # padded row length: every row starts on a 128-byte boundary; under frequency sharding
        # the frames of the TW numerators / denominators are reduce-scattered over the ranks in
        # equal chunks of whole 128-byte lines
        self.ld = _round_up(self.N, 32 * comm.world if (multi and shard == "freq") else 32)
        if self.F <= 0 or self.N <= 0:
            raise ValueError("empty shard f=%r n=%r" % (f_range, n_range))
        self.X = None
        self.I = 2  # channels; set_X_planes / set_X_host update it (3 and 4: general-I kernels)
        self.J = 0
        self.n_iter_done = 0
        # The spectral components are independent of one another inside each phase of the
        # M-step: their kernel chains run on side streams, so that the small kernels of one
        # component (~30 % of the launches, a few microseconds each) overlap the bandwidth-bound
        # kernels of the others.  PYFASST_STREAMS=0 keeps everything on one stream.
        self._side = []
        self._use_streams = self.dev.type == "cuda" and os.environ.get("PYFASST_STREAMS", "1") != "0"

    # ------------------------------------------------------------------ streams
    def _for_each(self, items, fn):
        """fn(index, item) for every item; on a GPU each item's launches go to its own side
        stream, forked from and joined back into the current stream (so the whole thing can
        still be captured in a CUDA graph)."""
        items = list(items)
        if not self._use_streams or len(items) < 2:
            for i, it in enumerate(items):
                fn(i, it)
            return
        torch = self.torch
        cur = torch.cuda.current_stream(self.dev)
        while len(self._side) < len(items):
            self._side.append(torch.cuda.Stream(device=self.dev))
        for i, it in enumerate(items):
            st = self._side[i]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                fn(i, it)
        for i in range(len(items)):
            cur.wait_stream(self._side[i])

    def _comm_stream(self):
        if getattr(self, "_cstream", None) is None:
            self._cstream = self.torch.cuda.Stream(device=self.dev)
        return self._cstream

    # ------------------------------------------------------------------ allocation
    def _zeros(self, shape, dtype=None):
        return self.torch.zeros(shape, dtype=dtype or self.tdtype, device=self.dev)

    def _upload(self, arr, dtype=None):
        """Host array -> new device tensor (always a copy, never an alias of `arr`).  On a GPU
        the dtype conversion runs on the device (the host side of a multi-rank job is often
        limited to one thread per process)."""
        a = np.ascontiguousarray(arr)
        if self.dev.type == "cuda":
            t = self.torch.from_numpy(a).to(self.dev)  # H2D copies: never aliases `arr`
            if dtype is not None and t.dtype != dtype:
                t = t.to(dtype)
            return t.contiguous()
        t = self.torch.tensor(a)
        if dtype is not None:
            t = t.to(dtype)
        return t.contiguous()

    def _upload_cols(self, arr, lo, hi, dtype):
        """arr[:, lo:hi] of a 2-D host array -> new device tensor: a proper column range (the
        frames of one rank) goes as ONE strided DMA instead of a host-side gather."""
        if self.dev.type == "cuda" and (lo > 0 or hi < arr.shape[1]) and arr.flags.c_contiguous \
                and hasattr(self.k, "copy_cols_to_device"):
            t = self.torch.empty((arr.shape[0], hi - lo), device=self.dev,
                                 dtype=self.torch.from_numpy(arr[:1, :1]).dtype)
            self.k.copy_cols_to_device(arr, lo, hi, t)
            return t if t.dtype == dtype else t.to(dtype)
        return self._upload(arr[:, lo:hi], dtype)

    def _f64(self, arr):
        return self._upload(np.asarray(arr, dtype=np.float64))

    def _sharded(self):
        return self.comm is not None and self.comm.world > 1

    def _fshard(self):
        return self._sharded() and self.shard == "freq"

    def _tshard(self):
        return self._sharded() and self.shard == "time"

    # ------------------------------------------------------------------ inputs
    def set_X_planes(self, X):
        """X: device planes [2 I, F_local, ld] (re, im per channel) of dtype self.tdtype.
        I = 2 is the reference's case; I = 3, 4 run the general-I kernels (an extension: the
        reference raises for them, audioModel.py:394, :605)."""
        assert X.dim() == 3 and tuple(X.shape[1:]) == (self.F, self.ld) and X.dtype == self.tdtype
        assert X.shape[0] in (4, 6, 8), "2..4 channels"
        self.X = X
        self.I = X.shape[0] // 2

    def set_X_host(self, Xc):
        """Xc: complex [I, F_total, N_total] host array (e.g. an STFT computed elsewhere)."""
        Xc = np.asarray(Xc)[:, self.f_lo:self.f_hi, self.n_lo:self.n_hi]
        nch = Xc.shape[0]
        planes = np.zeros([2 * nch, self.F, self.ld])
        for c in range(nch):
            planes[2 * c, :, :self.N], planes[2 * c + 1, :, :self.N] = Xc[c].real, Xc[c].imag
        self.set_X_planes(self._upload(planes, self.tdtype))

    def set_noise(self, sim_ann_opt, lim0, lim1, psd):
        """Mirrors the noise handling of estim_param_a_post_model (audioModel.py:355-373)."""
        sl = slice(self.f_lo, self.f_hi)
        self.sim_ann_opt = sim_ann_opt
        self.anneal = sim_ann_opt in ("ann", "ann_ns_inj")
        self.sqrt0 = self._f64(np.sqrt(np.asarray(lim0, dtype=np.float64)[sl]))
        self.sqrt1 = self._f64(np.sqrt(np.asarray(lim1, dtype=np.float64)[sl]))
        if sim_ann_opt == "ann":
            psd0 = np.asarray(lim0, dtype=np.float64)
        elif sim_ann_opt == "no_ann":
            psd0 = np.asarray(lim1, dtype=np.float64)
        else:
            psd0 = np.asarray(psd, dtype=np.float64)
        self.noise = self._f64(np.array(np.broadcast_to(psd0, (self.F_total,))[sl]))

    def set_model(self, spat_comps, spec_comps, nmfUpdateCoeff=1.0, lambdaCorr=0.0):
        """Validates the structure and uploads the parameters (host dicts -> HBM)."""
        torch = self.torch
        self.lambdaCorr = float(lambdaCorr)
        J = len(spat_comps)
        if sorted(spat_comps.keys()) != list(range(J)):
            raise ValueError("spat_comps keys must be 0..J-1")
        mix_types = [spat_comps[j]["mix_type"] for j in range(J)]
        self.free_A = [spat_comps[j]["frdm_prior"] == "free" for j in range(J)]
        if all(m == "inst" for m in mix_types):
            self.mix_type = "inst"
        elif all(m == "conv" for m in mix_types):
            self.mix_type = "conv"
            if any(self.free_A) and not all(self.free_A):
                raise NotImplementedError(
                    "convolutive mixing with fixed components: the reference solves with the "
                    "full hat_Rss (audioModel.py:857) and cannot run this either")
        else:
            raise NotImplementedError("mixed instantaneous / convolutive spatial components")
        # sub-source layout (retrieve_subsrc_params, audioModel.py:541-556)
        self.ranks, self.src_of_sub = [], []
        for j in range(J):
            p = np.asarray(spat_comps[j]["params"])
            r = p.shape[1] if mix_types[j] == "inst" else p.shape[0]
            nc = p.shape[0] if mix_types[j] == "inst" else p.shape[1]
            if nc != self.I:
                raise AttributeError("Nb channels %d not implemented yet" % nc
                                     if nc not in (2, 3, 4) else
                                     "mixing parameters for %d channels, signal has %d" % (nc, self.I))
            start = len(self.src_of_sub)
            self.ranks.append(list(range(start, start + r)))
            self.src_of_sub.extend([j] * r)
        R = len(self.src_of_sub)
        A = np.zeros([R, self.I, self.F], dtype=np.complex128)
        for j in range(J):
            p = np.asarray(spat_comps[j]["params"])
            if mix_types[j] == "inst":
                A[self.ranks[j]] = p.T[:, :, None]
            else:
                if p.shape[2] != self.F_total:
                    raise ValueError("convolutive params must be [rank, channels, F]")
                A[self.ranks[j]] = p[:, :, self.f_lo:self.f_hi]
        # (instantaneous parameters are real in the reference; a user-supplied complex array keeps
        # the general E-step kernel)
        self.real_mixing = self.mix_type == "inst" and not np.any(A.imag)
        self.J, self.R = J, R
        self.A = self._upload(A)
        self.omega = float(nmfUpdateCoeff)
        self._set_spectral(spec_comps)
        self._alloc_work()

    def _set_spectral(self, spec_comps):
        """Spectral components of the fast path: one per spatial component, single NMF factor,
        fixed FW (the structures of MultiChanNMFInst_FASST / MultiChanNMFConv,
        audioModel.py:2355-2391).  Anything else: GeneralGemEngine (engine_general.py)."""
        J = self.J
        if self.lambdaCorr > 0:
            raise NotImplementedError("lambdaCorr > 0 runs on GeneralGemEngine")
        # spectral components: one per spatial component, single NMF factor
        S = len(spec_comps)
        if sorted(spec_comps.keys()) != list(range(S)):
            raise ValueError("spec_comps keys must be 0..S-1")
        owner = [spec_comps[s]["spat_comp_ind"] for s in range(S)]
        if sorted(owner) != list(range(J)):
            raise NotImplementedError(
                "the device GEM path needs exactly one spectral component per spatial component")
        self.spec = []
        for s in range(S):
            facs = spec_comps[s]["factor"]
            if len(facs) != 1:
                raise NotImplementedError("multi-factor spectral components (source/filter "
                                          "models) are not on the device path yet")
            fac = facs[list(facs.keys())[0]]
            if len(fac["TB"]):
                raise NotImplementedError("time-blob factors (TB) run on GeneralGemEngine")
            if fac.get("TW_constr", "NMF") != "NMF":
                raise NotImplementedError("discrete-state TW constraints (GMM/HMM)")
            if fac["FW_frdm_prior"] == "free":
                raise NotImplementedError("free FW update is not on the device path yet")
            FB = np.asarray(fac["FB"], dtype=np.float64)
            FW = np.asarray(fac["FW"], dtype=np.float64)
            TW = np.asarray(fac["TW"], dtype=np.float64)
            if FB.shape[0] != self.F_total or TW.shape[1] != self.N_total or \
                    FW.shape != (FB.shape[1], TW.shape[0]):
                raise ValueError("inconsistent factor shapes FB%s FW%s TW%s"
                                 % (FB.shape, FW.shape, TW.shape))
            Kb, Kw = FW.shape
            TWd = self._zeros([Kw, self.ld])
            TWd[:, :self.N] = self._upload_cols(TW, self.n_lo, self.n_hi, self.tdtype)
            ent = {
                "j": owner[s], "Kb": Kb, "Kw": Kw,
                "FB_free": fac["FB_frdm_prior"] == "free",
                "TW_free": fac["TW_frdm_prior"] == "free",
                "FB": self._upload(FB[self.f_lo:self.f_hi], self.tdtype),
                "FW": self._upload(FW, self.tdtype),
                "TW": TWd,
            }
            # Plain NMF (the standard models, audioModel.py:2377-2380) has FW = identity, fixed:
            # W = FB FW and G = FW TW are then exact copies of FB and TW -- alias them instead
            # of launching the two products.  (The renormalisation keeps FW at exactly the
            # identity: it scales row k by c_k and divides column k by the same c_k.)
            ent["FW_identity"] = Kb == Kw and np.array_equal(FW, np.eye(Kb))
            if ent["FW_identity"]:
                ent["W"], ent["G"] = ent["FB"], ent["TW"]
            else:
                ent["W"] = self._zeros([self.F, Kw])
                ent["G"] = self._zeros([Kb, self.ld])
            self.spec.append(ent)

    def _alloc_work(self):
        self._alloc_common()
        self._alloc_spectral()

    def _alloc_common(self):
        torch, k = self.torch, self.k
        F, ld, J, R, N = self.F, self.ld, self.J, self.R, self.N
        f64, c128 = torch.float64, torch.complex128
        self.V = self._zeros([J, F, ld])
        self.hatW = self._zeros([J, F, ld])
        # hat_Rss, hat_Rxs and the per-frequency log-likelihood sums share one buffer so that
        # the time-sharded mode reduces them with a single all-reduce
        I = self.I
        self.estat = self._zeros([F * (2 * R * R + 2 * I * R + 1)], f64)
        o1, o2 = 2 * F * R * R, 2 * F * R * R + 2 * I * F * R
        self.Rss = torch.view_as_complex(self.estat[:o1].view(F, R, R, 2))
        self.Rxs = torch.view_as_complex(self.estat[o1:o2].view(F, I, R, 2))
        self.ll_f = self.estat[o2:]
        self.ll_sum = self._zeros([1], f64)
        self.flags = self._zeros([1], torch.int32)
        self.iter_dev = self._zeros([1], torch.int32)
        # first iteration (iter_dev at that time) at which a TW vanished, see redraw_vanished_TW
        self.first_vanish = torch.full([1], 2 ** 30, dtype=torch.int32, device=self.dev)
        self.sync_redraw = False
        self.redrawn = 0
        code = k.dtype_code(self.V)
        # the stereo kernels for I = 2 (PYFASST_FORCE_MULTI=1: the general-I ones, for tests)
        self.multi = I != 2 or os.environ.get("PYFASST_FORCE_MULTI") == "1"
        nbytes = k.estep_multi_workspace_bytes(I, J, F, N) if self.multi else \
            k.estep_workspace_bytes(J, F, N, code)
        self.ws = self._zeros([(nbytes + 7) // 8], f64)
        self.stats = self._zeros([I * R + R * R], f64)
        self.sums = self._zeros([J], f64)
        counts = np.array([len(self.ranks[j]) * I * self.F_total for j in range(J)], dtype=np.float64)
        self.counts = self._f64(counts)

    def _alloc_spectral(self):
        torch, k = self.torch, self.k
        F, ld, N = self.F, self.ld, self.N
        f64 = torch.float64
        code = k.dtype_code(self.V)
        Kmax = max(max(e["Kb"], e["Kw"]) for e in self.spec)
        S = len(self.spec)
        self.colmax = self._zeros([S, Kmax], f64)
        self.wcol = self._zeros([S, Kmax], f64)
        self.w2 = self._zeros([S, Kmax], f64)
        self._totals2 = self._zeros([2, S], f64)
        self._tot_parity = 0
        self.totals = self._totals2[0]
        # FB update partial sums (one buffer, reused per component)
        self.fb_plan, fb_size = {}, 0
        for e in self.spec:
            chunk, nsplit = k.fb_plan(F, e["Kb"], N, code)
            self.fb_plan[id(e)] = (chunk, nsplit)
            fb_size = max(fb_size, nsplit * F * e["Kb"])
        self.fb_part = self._zeros([S, 2, fb_size], f64)  # per component: they run concurrently
        self.fb_nd = self._zeros([S, 2, F * Kmax], f64)
        # TW update partial sums; the reduced num/den of all components share one
        # buffer so that a single all-reduce serves the whole iteration
        self.tw_plan, tw_size = {}, 0
        for e in self.spec:
            fchunk, fsplit = k.tw_plan(F, e["Kw"], N, code)
            self.tw_plan[id(e)] = (fchunk, fsplit)
            tw_size = max(tw_size, fsplit * e["Kw"] * ld)
        self.tw_part = self._zeros([S, 2, tw_size], f64)
        # the views the contraction kernels write, made once (tensor indexing costs the host
        # several microseconds per view, 16 views per iteration)
        for s, e in enumerate(self.spec):
            nsplit = self.fb_plan[id(e)][1]
            cnt = F * e["Kb"]
            e["fb_pn"] = self.fb_part[s, 0, :nsplit * cnt].view(nsplit, F, e["Kb"])
            e["fb_pd"] = self.fb_part[s, 1, :nsplit * cnt].view(nsplit, F, e["Kb"])
            fsplit = self.tw_plan[id(e)][1]
            cnt = e["Kw"] * ld
            e["tw_pn"] = self.tw_part[s, 0, :fsplit * cnt].view(fsplit, e["Kw"], ld)
            e["tw_pd"] = self.tw_part[s, 1, :fsplit * cnt].view(fsplit, e["Kw"], ld)
        # work planes for P' = W'H of the two-kernel tensor-core TW path (float32 planes only;
        # PYFASST_TW_FUSED=0).  Not zeroed: that path writes a plane in full before it reads it,
        # the default fused kernel never touches it (a 1.7 GB fill per model of configs[1]).
        self.scratch = torch.empty([S, F, ld], dtype=self.tdtype, device=self.dev) \
            if self.tdtype == torch.float32 else None
        for s, e in enumerate(self.spec):
            e["scratch"] = None if self.scratch is None else self.scratch[s]
        if self._fshard():
            # reduce-scatter over the frames (plane type, fixed order) -> shard-local update ->
            # all-gather of TW, per component on its side stream: 2.7x less NVLink
            # traffic than the float64 all-reduce of num and den, and it overlaps the contraction
            # of the next component
            world = self.comm.world
            c = ld // world
            self.tw_rs_in = self._zeros([S, world, 2, Kmax, c])
            self.tw_rs_out = self._zeros([S, 2, Kmax, c])
            self.tw_ag = self._zeros([S, world, Kmax, c])

    # ------------------------------------------------------------------ pieces
    def compute_powers(self, with_G=True):
        """W = FB FW, V_j = W H (comp_spat_comp_power, audioModel.py:430-498), G = FW H."""
        k = self.k

        def powers(s, e):
            if not e["FW_identity"]:
                k.small_matmul(e["FB"], e["FW"], e["W"])
            k.spec_power(e["W"], e["TW"], self.V[e["j"]], self.N, False)
            if with_G and e["FB_free"] and not e["FW_identity"]:
                k.spec_power(e["FW"], e["TW"], e["G"], self.N, False)
        self._for_each(self.spec, powers)

    def estep(self, for_update=False):
        """compute_suff_stat (audioModel.py:580-764) on the current parameters.  for_update: the
        statistics only feed this iteration's mixing update -- with instantaneous mixing (real
        mixing vectors) that update takes their real parts (audioModel.py:818-820), and the stereo
        kernel then skips the imaginary moments (pf_estep_stereo_inst)."""
        if self.multi:
            estep = self.k.estep_multi
        elif for_update and getattr(self, "real_mixing", False) and hasattr(self.k, "estep_stereo_inst") \
                and os.environ.get("PYFASST_ESTEP_INST", "1") != "0":
            estep = self.k.estep_stereo_inst
        else:
            estep = self.k.estep_stereo
        estep(self.X, self.V, self.A, self.src_of_sub, self.noise, self.N, self.hatW, self.Rss,
              self.Rxs, self.ll_f, self.ws, self.N_total)

    def reduce_estat(self):
        """Frame sharding: the statistics are means over all frames -- sum the ranks' parts."""
        if self._tshard():
            self.comm.allreduce_sum(self.estat)

    def update_mix(self):
        """update_mix_matrix (audioModel.py:766-889)."""
        k = self.k
        if not any(self.free_A):
            return
        if self.mix_type == "inst":
            upd = [r for j in range(self.J) if self.free_A[j] for r in self.ranks[j]]
            oth = [r for j in range(self.J) if not self.free_A[j] for r in self.ranks[j]]
            n = self.I * len(upd) + len(upd) ** 2
            stats = self.stats[:n]
            k.mix_inst_stats(self.Rss, self.Rxs, self.A, upd, oth, stats)
            if self._fshard():
                self.comm.allreduce_sum(stats)
            k.mix_inst_solve(stats, self.F_total, upd, self.A, self.flags)
        else:
            k.mix_conv_solve(self.Rss, self.Rxs, self.A, self.flags)

    def update_spectral(self):
        """NMF branch of update_spectral_components (audioModel.py:1509-1727): for every
        component FB (frames contracted, local rows) then TW (frequencies contracted,
        all-reduced under sharding).  V_j still holds the pre-update power: it is both
        `comp_spat_comp_power(spat_ind)` of the FB update (Q3) and the stale
        `other_fact_power` (Q1/Q2)."""
        k, N, F = self.k, self.N, self.F
        # FB: contraction over the (local) frames.  Different components are independent given
        # hat_W and V, so under time sharding all FB sums are reduced with one all-reduce.
        fb = [(s, e) for s, e in enumerate(self.spec) if e["FB_free"]]

        def fb_sums(_, se):
            s, e = se
            j = e["j"]
            chunk, nsplit = self.fb_plan[id(e)]
            cnt = F * e["Kb"]
            pn, pd = e["fb_pn"], e["fb_pd"]
            k.fb_contract(self.hatW[j], self.V[j], self.V[j], e["G"], N, pn, pd, chunk, nsplit)
            if self._tshard():
                k.sum_splits(pn, self.fb_nd[s, 0, :cnt])
                k.sum_splits(pd, self.fb_nd[s, 1, :cnt])
            else:
                k.mult_update_splits(e["FB"], pn, pd, F, e["Kb"], self.omega)
                if not e["FW_identity"]:
                    k.small_matmul(e["FB"], e["FW"], e["W"])

        def fb_apply(_, se):
            s, e = se
            cnt = F * e["Kb"]
            k.mult_update(e["FB"], self.fb_nd[s, 0, :cnt].view(F, e["Kb"]),
                          self.fb_nd[s, 1, :cnt].view(F, e["Kb"]), F, e["Kb"], self.omega)
            if not e["FW_identity"]:
                k.small_matmul(e["FB"], e["FW"], e["W"])
        if self._fshard() and self._use_streams and os.environ.get("PYFASST_FREQ_PIPELINE", "1") != "0":
            # Frequency partition: one component after the other on this stream (FB_s, TW_s), the
            # exchange of component s on the high-priority stream while component s + 1 is
            # contracted -- only the last exchange is exposed.  Equivalent to the phase order
            # below: FB_s reads its own component's TW only, which changes after FB_s.
            for s, e in enumerate(self.spec):
                if e["FB_free"]:
                    fb_sums(0, (s, e))
                if e["TW_free"]:
                    self._tw_contract_and_exchange(s, e)
            self.torch.cuda.current_stream(self.dev).wait_stream(self._exchange_stream())
            return
        if not self._sharded() and self._use_streams and os.environ.get("PYFASST_MSTEP_CHAINS", "1") != "0":
            # One chain per component (FB_s, then TW_s) on its side stream instead of two joined
            # phases: FB_s reads its own component's TW only, which changes after FB_s, and
            # nothing is shared between the components of this engine -- the same arithmetic,
            # one join less (its bubble: the tails of four bandwidth-bound kernels).
            def chain(s, e):
                if e["FB_free"]:
                    fb_sums(0, (s, e))
                if e["TW_free"]:
                    self._tw_contract_and_exchange(s, e)
            self._for_each(list(enumerate(self.spec)), lambda _, se: chain(*se))
            return
        # NB every FB update reads G_s = FW_s TW_s of its own component only, and no TW changes
        # before all the FB updates are done (Gauss-Seidel order FB -> TW, Q2)
        self._for_each(fb, fb_sums)
        if fb and self._tshard():
            self.comm.allreduce_sum(self.fb_nd)
            self._for_each(fb, fb_apply)
        # TW: contraction over the (local) frequencies with the updated W
        tw = [(s, e) for s, e in enumerate(self.spec) if e["TW_free"]]

        self._for_each(tw, lambda _, se: self._tw_contract_and_exchange(*se))
        if tw and self._fshard() and self._use_streams:
            self.torch.cuda.current_stream(self.dev).wait_stream(self._exchange_stream())

    def _tw_contract_and_exchange(self, s, e):
        """TW update of one component: contraction over the (local) frequencies with the updated W,
        then the multiplicative update (after the exchange under frequency sharding)."""
        k = self.k
        j = e["j"]
        fchunk, fsplit = self.tw_plan[id(e)]
        pn, pd = e["tw_pn"], e["tw_pd"]
        k.tw_contract(self.hatW[j], self.V[j], e["W"], e["TW"], self.N, pn, pd, fchunk, fsplit,
                      e["scratch"])
        if self._fshard():
            self._tw_exchange(s, e, pn, pd)
        else:  # reduce the frequency splits inside the update kernel
            k.mult_update_splits(e["TW"], pn, pd, e["Kw"], self.N, self.omega)

    def _exchange_stream(self):
        """High-priority stream of the frequency partition's TW exchange: the chain of component s
        (pack, reduce-scatter, update, all-gather, unpack) starts as soon as its contraction is
        done and its small kernels get SM slots ahead of the CTAs of the next component's
        contraction, which fills the GPU."""
        if getattr(self, "_xstream", None) is None:
            self._xstream = self.torch.cuda.Stream(device=self.dev, priority=-1)
        return self._xstream

    def _tw_exchange(self, s, e, pn, pd):
        """Frequency partition: sum over the ranks of the TW numerators / denominators as a
        reduce-scatter over the FRAMES (plane type, fixed order), shard-local multiplicative
        update, all-gather of the updated TW."""
        k = self.k
        Kw, world, rank = e["Kw"], self.comm.world, self.comm.rank
        c = self.ld // world

        def chain():
            # split sums -> plane type, chunk-major [world, 2, Kmax, c]: one kernel
            k.tw_pack_chunks(pn, pd, self.tw_rs_in[s], world)
            self.comm.reduce_scatter_sum(self.tw_rs_out[s], self.tw_rs_in[s])
            # this rank's frames of TW (padding frames: 0 * (0 / eps) = 0)
            chunk = e["TW"][:, rank * c:(rank + 1) * c]
            k.mult_update_same(chunk, self.tw_rs_out[s, 0, :Kw], self.tw_rs_out[s, 1, :Kw],
                               Kw, c, self.omega)
            self.tw_ag[s, rank, :Kw].copy_(chunk)
            self.comm.allgather_into(self.tw_ag[s], self.tw_ag[s, rank])
            e["TW"].view(Kw, world, c).copy_(self.tw_ag[s, :, :Kw].permute(1, 0, 2))
        if not self._use_streams:
            chain()
            return
        torch = self.torch
        xs = self._exchange_stream()
        xs.wait_stream(torch.cuda.current_stream(self.dev))  # (the component's side stream)
        with torch.cuda.stream(xs):
            chain()

    def renormalize(self):
        """renormalize_parameters (audioModel.py:1980-2040)."""
        k = self.k
        if self._tshard() and self._use_streams:  # (the check issued one iteration ago: long done)
            self.torch.cuda.current_stream(self.dev).wait_stream(self._comm_stream())
        k.spat_energy(self.A, self.src_of_sub, self.J, self.sums)
        if self._fshard():
            self.comm.allreduce_sum(self.sums)
        k.spat_scale(self.A, self.src_of_sub, self.sums, self.counts)
        def colmax(s, e):
            k.fb_scale_colmax(e["FB"], self.sums, self.counts, e["j"], self.colmax[s])

        def rescale(s, e):
            k.fw_renorm(e["FW"], self.colmax[s], self.wcol[s], self.w2[s])
            k.scale_matrix(e["FB"], self.F, e["Kb"], self.wcol[s], False, True)
            k.scale_matrix(e["TW"], e["Kw"], self.N, self.w2[s], True, False, self.totals[s:s + 1])
        if self._fshard():
            self._for_each(self.spec, colmax)
            self.comm.allreduce_max(self.colmax)
            self._for_each(self.spec, rescale)
        else:
            self._for_each(self.spec, lambda s, e: (colmax(s, e), rescale(s, e)))
        if self._tshard() and self._use_streams and not self.sync_redraw:
            # the sum over the ranks only feeds the vanished-TW check: reduce and check it on
            # the side stream, off the critical path (the next renormalisation accumulates into
            # the other buffer and waits for this check before it comes back to this one)
            torch = self.torch
            cur, side = torch.cuda.current_stream(self.dev), self._comm_stream()
            tot = self.totals
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.comm.allreduce_sum(tot)
                k.check_totals(tot, EPS, self.flags, self.iter_dev, self.first_vanish)
            self._tot_parity ^= 1
            self.totals = self._totals2[self._tot_parity]
            return
        if self._tshard():
            self.comm.allreduce_sum(self.totals)
        if self.sync_redraw:
            self.redraw_vanished_TW()
        k.check_totals(self.totals, EPS, self.flags, self.iter_dev, self.first_vanish)

    def redraw_vanished_TW(self):
        """Host-synchronous part of the renormalisation: a TW whose sum fell below eps is re-drawn
        with np.random (randn(K, N)^2 * 1e3 eps), like the reference (audioModel.py:2023-2025).
        For the single-factor components of this engine nothing of the renormalisation depends on
        TW afterwards, so doing it after the device kernels is exact.  Under frame sharding every
        rank draws the whole matrix (same np.random state assumed) and keeps its frames."""
        tot = self.totals.cpu().numpy()
        for s, e in enumerate(self.spec):
            if tot[s] < EPS:
                Z = np.random.randn(e["Kw"], self.N_total) ** 2 * (1e3 * EPS)
                e["TW"][:, :self.N] = self._upload(Z[:, self.n_lo:self.n_hi], self.tdtype)
                self.totals[s] = float(Z.sum())
                self.redrawn += 1

    def gem_iteration(self, n_iter_total, logliks, mark=None):
        """One GEM iteration (audioModel.py:364-376, :384-428), fully stream-ordered.
        `mark(label)` (optional) is called between the phases -- bench.py records CUDA
        events there."""
        k = self.k
        mark = mark or (lambda label: None)
        mark("begin")
        if self.anneal:
            k.noise_anneal(self.sqrt0, self.sqrt1, self.iter_dev, n_iter_total, self.noise)
        self.compute_powers()
        mark("powers")
        self.estep(for_update=True)
        mark("estep")

        def stats_and_mix():
            self.reduce_estat()
            k.ll_reduce(self.ll_f, self.ll_sum)
            if self._fshard():
                self.comm.allreduce_sum(self.ll_sum)
            k.ll_store(self.ll_sum, float(self.F_total) * self.N_total, logliks, self.iter_dev, True)
            self.update_mix()
        if self._use_streams:
            # only the mixing update needs the (all-reduced) statistics: it runs, with its
            # collectives, on a side stream while the spectral M-step (which reads hat_W, V and
            # the factors, never A) starts on this one
            torch = self.torch
            cur = torch.cuda.current_stream(self.dev)
            side = self._comm_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                stats_and_mix()
            mark("mix")
            self.update_spectral()
            cur.wait_stream(side)
        else:
            stats_and_mix()
            mark("mix")
            self.update_spectral()
        mark("spectral")
        self.renormalize()
        mark("renorm")

    # ------------------------------------------------------------------ drivers
    def run(self, n_iter, use_graph=False, careful_from=None):
        """estim_param_a_post_model: n_iter GEM iterations; returns logliks (host).  The loop is
        enqueued without a single host synchronisation.  `careful_from` = i: from iteration i on
        the renormalisation synchronises with the host so that a vanished TW is re-drawn with
        np.random exactly where the reference would (see vanished_iteration)."""
        torch = self.torch
        logliks = torch.ones([max(n_iter, 1)], dtype=torch.float64, device=self.dev)
        self.iter_dev.zero_()
        self.flags.zero_()
        self._totals2.zero_()
        self.first_vanish.fill_(2 ** 30)
        if use_graph and n_iter > 1 and self.dev.type == "cuda" and careful_from is None:
            self._run_graph(n_iter, logliks)
        else:
            for it in range(n_iter):
                self.sync_redraw = careful_from is not None and it >= careful_from
                self.gem_iteration(n_iter, logliks)
            self.sync_redraw = False
        self.n_iter_done = n_iter
        self.check_flags()
        return logliks[:n_iter].cpu().numpy()

    def vanished_iteration(self):
        """Index of the first iteration of the last run() whose renormalisation found a TW with
        sum < eps and did NOT re-draw it (None: none).  The caller replays the run from the same
        initial parameters with careful_from = that index: the kernels are deterministic, so the
        replay reproduces the iterations before it bit for bit."""
        first = int(self.first_vanish.cpu().item())
        return None if first >= 2 ** 30 else max(first - 1, 0)

    def _run_graph(self, n_iter, logliks):
        torch = self.torch
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.gem_iteration(n_iter, logliks)  # warm-up iteration, also iteration 0
        torch.cuda.current_stream(self.dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        # (not the `torch.cuda.graph` context manager: it empties the caching allocator on entry)
        cap = torch.cuda.Stream(device=self.dev)
        cap.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(cap):
            graph.capture_begin()
            try:
                self.gem_iteration(n_iter, logliks)
            finally:
                graph.capture_end()
        torch.cuda.current_stream(self.dev).wait_stream(cap)
        for _ in range(n_iter - 1):
            graph.replay()

    def check_flags(self):
        flags = int(self.flags.cpu().item())
        if flags & 1:
            raise np.linalg.LinAlgError("Singular Matrix")
        # (flags & 2: a TW vanished and was not re-drawn on the spot -- see vanished_iteration)

    def suff_stat(self):
        """One E-step on the current parameters; returns host arrays shaped like the
        reference's compute_suff_stat outputs (local frequency rows)."""
        self.compute_powers(with_G=False)
        self.estep()
        self.reduce_estat()
        self.k.ll_reduce(self.ll_f, self.ll_sum)
        if self._fshard():
            self.comm.allreduce_sum(self.ll_sum)
        ll = -float(self.ll_sum.cpu().item()) / (float(self.F_total) * self.N_total)
        hat_Rxs = self.Rxs.cpu().numpy()
        hat_Rss = self.Rss.cpu().numpy()
        hat_W = self._gather_n(self.hatW[:, :, :self.N].to(self.torch.float64), 2)
        return hat_Rxs, hat_Rss, hat_W, ll

    # ------------------------------------------------------------------ outputs
    def _gather(self, t, axis, total, which):
        """Host array of a tensor sharded along `axis` (`which` = 'freq' or 'time'),
        concatenated over the ranks; replicated tensors are read locally."""
        if not self._sharded() or self.shard != which:
            return t.cpu().numpy()
        shards = shard_bounds(total, self.comm.world)
        smax = max(hi - lo for lo, hi in shards)
        t = t.movedim(axis, 0)
        pad = t.new_zeros((smax,) + tuple(t.shape[1:]))
        pad[:t.shape[0]] = t
        parts = self.comm.allgather(pad)
        full = self.torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, shards)], dim=0)
        return full.movedim(0, axis).cpu().numpy()

    def _gather_f(self, t, axis):
        return self._gather(t, axis, self.F_total, "freq")

    def _gather_n(self, t, axis):
        """`t` holds the local frames (without padding) along `axis`."""
        return self._gather(t, axis, self.N_total, "time")

    def read_model(self, spat_comps, spec_comps, gather=True):
        """Writes the device parameters back into the user-visible dicts.  Sharded models:
        `gather` = True all-gathers the sharded factor (TW under frame sharding, FB under frequency
        sharding) so that every rank holds the whole model; False writes only this rank's rows /
        frames into the host arrays (a copy whose size does not grow with the number of ranks --
        what estim_param_a_post_model does; FASST.gather_parameters() completes them)."""
        A = self._gather_f(self.A, 2)
        for j in range(self.J):
            if self.mix_type == "inst":
                # Q6: complex dtype after the update (audioModel.py:878-882)
                spat_comps[j]["params"] = np.ascontiguousarray(A[self.ranks[j], :, 0].T)
            else:
                spat_comps[j]["params"] = np.ascontiguousarray(A[self.ranks[j]])
        for s, e in enumerate(self.spec):
            fac = spec_comps[s]["factor"]
            fac = fac[list(fac.keys())[0]]
            f64 = self.torch.float64
            if gather or not self._sharded():
                fac["FB"] = self._to_host(
                    self._gather_dev(e["FB"].to(f64), 0, self.F_total, "freq"), fac["FB"])
                fac["TW"] = self._to_host(
                    self._gather_dev(e["TW"][:, :self.N].to(f64), 1, self.N_total, "time"),
                    fac["TW"])
            else:
                fac["FB"] = self._to_host_part(e["FB"].to(f64), fac["FB"], 0, self.f_lo, self.f_hi,
                                               (self.F_total, e["Kb"]))
                fac["TW"] = self._to_host_part(e["TW"][:, :self.N].to(f64), fac["TW"], 1,
                                               self.n_lo, self.n_hi, (e["Kw"], self.N_total))
            fac["FW"] = e["FW"].to(f64).cpu().numpy()

    def _to_host_part(self, t, old, axis, lo, hi, full_shape):
        """Device tensor holding the slice [lo, hi) along `axis` of a host array of `full_shape`:
        written in place into `old` when that is a writable float64 array of that shape (else into
        a copy of it / a zero array)."""
        if not (isinstance(old, np.ndarray) and old.dtype == np.float64 and old.flags.writeable
                and old.shape == tuple(full_shape)):
            old = np.array(old, dtype=np.float64) if np.shape(old) == tuple(full_shape) \
                else np.zeros(full_shape)
        view = old[lo:hi] if axis == 0 else old[:, lo:hi]
        if t.is_cuda and axis == 1 and old.ndim == 2 and old.flags.c_contiguous \
                and t.dtype == self.torch.float64 and hasattr(self.k, "copy_cols_to_host"):
            self.k.copy_cols_to_host(t.contiguous(), old, lo, hi)  # one strided DMA
        elif t.is_cuda:
            self.torch.from_numpy(view).copy_(t)
        else:
            view[...] = t.numpy()
        return old

    def _to_host(self, t, old):
        """Device tensor -> host array.  When the user-visible array it replaces has the same
        shape and is a writable C-contiguous float64 array, the copy lands in it IN PLACE (the
        reference updates its parameter arrays in place as well, audioModel.py:1573, :1725): the
        pages are already mapped, which halves the device-to-host time of the 10-minute TW
        matrices compared with a freshly allocated array."""
        if isinstance(old, np.ndarray) and old.dtype == np.float64 and old.flags.c_contiguous \
                and old.flags.writeable and old.shape == tuple(t.shape) and t.is_cuda:
            self.torch.from_numpy(old).copy_(t)
            return old
        return t.cpu().numpy()

    def _gather_dev(self, t, axis, total, which):
        """Like _gather, but the result stays a device tensor (replicated / unsharded case) or
        is the gathered device tensor."""
        if not self._sharded() or self.shard != which:
            return t
        shards = shard_bounds(total, self.comm.world)
        smax = max(hi - lo for lo, hi in shards)
        t = t.movedim(axis, 0)
        pad = t.new_zeros((smax,) + tuple(t.shape[1:]))
        pad[:t.shape[0]] = t
        parts = self.comm.allgather(pad)
        full = self.torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, shards)], dim=0)
        return full.movedim(0, axis)

    def noise_psd(self):
        return self._gather_f(self.noise, 0)

    def component_power(self, s):
        """Power plane [F, ld] of spectral component s (after compute_powers)."""
        return self.V[self.spec[s]["j"]]

    def wiener_custom(self, V, A, src_of_sub, group_of_src, ngroups):
        """Wiener filter for an explicit list of sources: powers V [Jv, F, ld], mixing vectors
        A [Rv, I, F] of their sub-sources, output group of each source."""
        Y = self._zeros([ngroups * 2 * self.I, self.F, self.ld])
        wiener = self.k.wiener_multi if self.multi else self.k.wiener_stereo
        wiener(self.X, V, A, src_of_sub, self.noise, group_of_src, ngroups, self.N, Y, self.ws)
        return Y

    def wiener(self, group_of_src, ngroups):
        """Wiener-filtered STFTs Y[g] = Sigma_g Sigma_x^-1 x as planes [ngroups * 2 I, F, ld]
        (compute_sigma_comp_2d / compute_inv_sigma_mix_2d / compute_Wiener_gain_2d,
        audioModel.py:1327-1467, with the last iteration's noise PSD, :1385)."""
        self.compute_powers(with_G=False)
        Y = self._zeros([ngroups * 2 * self.I, self.F, self.ld])
        wiener = self.k.wiener_multi if self.multi else self.k.wiener_stereo
        wiener(self.X, self.V, self.A, self.src_of_sub, self.noise, group_of_src, ngroups, self.N,
               Y, self.ws)
        return Y
