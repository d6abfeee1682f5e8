"""FASST audio source separation models on B200 (drop-in for pyfasst/audioModel.py).

Same class names, constructor arguments, public methods and user-visible state as the
reference (`FASST` audioModel.py:66-2294, `MultiChanNMFInst_FASST` :2296-2420,
`MultiChanNMFConv` :2422-2508, `multiChanSourceF0Filter` :2551-3014): `spat_comps`, `spec_comps` and `noise` stay plain
NumPy-holding dicts that the user may read and edit between calls
(doc/source/description.rst:125-197).  Every numerical method packs that state into HBM,
runs the sm_100a kernels through the C ABI (GemEngine, pyfasst_b200/engine.py) and
writes the result back -- there is no NumPy implementation of the maths in here and no
CPU fallback: without the CUDA library / a GPU the methods raise.

Extra keyword arguments (not in the reference):
    compute_dtype : 'float32' (default, fast path) or 'float64' (the reference's precision)
    kernels       : kernel provider (default: the CUDA kernels)
    comm          : pyfasst_b200.engine.Comm to shard one mixture over several GPUs
    shard         : 'time' (frames, default: a few MB of collectives per iteration) or 'freq'
                    (frequency bins; the TW numerators/denominators are all-reduced)
    use_cuda_graph: replay the GEM iteration as a CUDA graph
"""
import os
import warnings

import numpy as np

from . import audioObject as ao
from .engine import GemEngine
from .engine_general import GeneralGemEngine
from .tftransforms.tft import tftransforms
from .tftransforms import stft as _stft

eps = 1e-10  # ref: audioModel.py:61


class FASST(object):
    """Flexible Audio Source Separation Toolbox model (ref: audioModel.py:66-2294)."""

    implemented_transf = ['stft']

    def __init__(self, audio, transf='stft', wlen=2048, hopsize=512, iter_num=50,
                 sim_ann_opt='ann', ann_PSD_lim=[None, None], verbose=0, nmfUpdateCoeff=1.,
                 tffmin=25, tffmax=18000, tfWinFunc=None, tfbpo=48, lambdaCorr=0.,
                 compute_dtype='float32', kernels=None, comm=None, use_cuda_graph=False,
                 shard='time'):
        self.verbose = verbose
        self.nmfUpdateCoeff = nmfUpdateCoeff
        if isinstance(audio, ao.AudioObject):
            self.audioObject = audio
        elif isinstance(audio, str):
            self.audioObject = ao.AudioObject(filename=audio)
        else:
            raise AttributeError("The provided audio parameter is not a supported format.")
        self.sig_repr_params = {
            'transf': transf.lower(),
            'wlen': ao.nextpow2(wlen),   # ref: audioModel.py:192-193
            'fsize': ao.nextpow2(wlen),
            'hopsize': hopsize,
            'tffmin': tffmin, 'tffmax': tffmax, 'tfbpo': tfbpo, 'tfWinFunc': tfWinFunc,
        }
        self.sig_repr_params['hopfactor'] = 1. * hopsize / self.sig_repr_params['wlen']
        if self.sig_repr_params['transf'] not in self.implemented_transf \
                or self.sig_repr_params['transf'] not in tftransforms:
            raise NotImplementedError(self.sig_repr_params['transf'] + " not yet implemented.")
        if compute_dtype not in ('float32', 'float64'):
            raise ValueError("compute_dtype must be 'float32' or 'float64'")
        self.compute_dtype = compute_dtype
        self._kernels = kernels
        self._comm = comm
        self._shard = shard
        self._use_cuda_graph = use_cuda_graph
        # NB the reference does not forward tfWinFunc: the window is always Hann
        # (audioModel.py:206-214 / stft.py:361)
        self.tft = tftransforms[self.sig_repr_params['transf']](
            fmin=tffmin, fmax=tffmax, bins=tfbpo, fs=self.audioObject.samplerate, perfRast=1,
            linFTLen=self.sig_repr_params['fsize'],
            atomHopFactor=self.sig_repr_params['hopfactor'], kernels=kernels)
        self.demixParams = {
            'tffmin': tffmin, 'tffmax': tffmax, 'tfbpo': tfbpo,
            'tfrepresentation': transf.lower(), 'wlen': self.sig_repr_params['wlen'],
            'hopsize': self.sig_repr_params['wlen'] // 2, 'neighbors': 20, 'winFunc': tfWinFunc}
        self.noise = {
            'PSD': np.zeros(self.sig_repr_params['fsize'] // 2 + 1),
            'sim_ann_opt': sim_ann_opt,
            # (a copy: the reference stores the argument itself and fills it in place, so that
            # its mutable default [None, None] leaks the limits of the first model of a process
            # into every later one, audioModel.py:171,:236,:316-322 -- not reproduced)
            'ann_PSD_lim': list(ann_PSD_lim),
        }
        self.spat_comps = {}
        self.spec_comps = {}
        self.iter_num = iter_num
        self.lambdaCorr = lambdaCorr
        self._X = None
        self._Cx = None

    # ------------------------------------------------------------------ plumbing
    def _k(self):
        if self._kernels is None:
            self._kernels = _stft.default_kernels()
        return self._kernels

    def _engine(self, psd_mode=None):
        """A GEM engine loaded with the current user-visible state (H9: the dicts are
        the source of truth at every public-method entry)."""
        if self._X is None:
            self.comp_transf_Cx()
        cls = GeneralGemEngine if self._general_structure() else GemEngine
        eng = cls(self._k(), self.nbFreqsSigRepr, self.nbFramesSigRepr,
                  dtype=self.compute_dtype, comm=self._comm, shard=self._shard)
        eng.set_X_planes(self._X)
        lim = self.noise['ann_PSD_lim']
        opt = self.noise['sim_ann_opt'] if psd_mode is None else psd_mode
        eng.set_noise(opt, lim[0], lim[1], self.noise['PSD'])
        eng.set_model(self.spat_comps, self.spec_comps, self.nmfUpdateCoeff, self.lambdaCorr)
        # (the sparsity re-weighting belongs to estim_param_a_post_model, not to GEM_iteration)
        eng.sparsity_enabled = psd_mode is None
        return eng

    def _general_structure(self):
        """True when the spectral structure is not the single-factor NMF of the standard models
        (several factors, free FW, time blobs, several spectral components per source, parameter
        arrays shared between components, large dictionaries, lambdaCorr > 0): GeneralGemEngine
        (engine_general.py) runs those."""
        if self.lambdaCorr > 0:
            return True
        owners, seen = [], set()
        for spec in self.spec_comps.values():
            owners.append(spec['spat_comp_ind'])
            if len(spec['factor']) != 1:
                return True
            fac = list(spec['factor'].values())[0]
            if fac['FW_frdm_prior'] == 'free' or np.shape(fac['FW'])[0] > 64 or len(fac['TB']):
                return True
            for name in ('FB', 'FW', 'TW'):
                # arrays shared between components stay ONE device buffer on GeneralGemEngine
                # (the reference rescales the shared object once per component, quirk Q11)
                if id(fac[name]) in seen:
                    return True
                seen.add(id(fac[name]))
        return sorted(owners) != list(range(len(self.spat_comps)))

    # ------------------------------------------------------------------ K1
    def comp_transf_Cx(self):
        """Signal representation: STFT of every channel on the device, noise-annealing
        limits from the mixture PSD (ref: audioModel.py:250-328).  `Cx` (rank one,
        Cx[n1,n2] = X[n1] conj(X[n2]), :293-302) is not stored: the kernels read X; the
        `Cx` attribute is built on demand."""
        import torch
        if self.sig_repr_params['transf'] not in self.implemented_transf:
            raise ValueError(self.sig_repr_params['transf'] + " not implemented - yet?")
        k = self._k()
        aobj = self.audioObject
        hop, nfft = self.sig_repr_params['hopsize'], self.sig_repr_params['fsize']
        wlen = self.tft.window.size
        if not hasattr(aobj, '_data') and not hasattr(aobj, '_raw'):
            aobj._read_raw()
        L = np.asarray(aobj._data).shape[0] if hasattr(aobj, '_data') else aobj._nframes
        N = _stft.number_of_frames(L, hop)
        multi = self._comm is not None and self._comm.world > 1
        frames, s_lo, s_hi = None, 0, L
        if multi and self._shard == 'time':
            # this rank transforms its own frames only and ships only the samples they cover
            from .engine import shard_bounds
            frames = shard_bounds(N, self._comm.world)[self._comm.rank]
            s_lo = min(max(0, frames[0] * hop - wlen // 2), L - 1)
            s_hi = max(min(L, (frames[1] - 1) * hop + wlen - wlen // 2), s_lo + 1)
        if hasattr(aobj, '_data'):
            # the user already holds the scaled float64 samples (e.g. set through `.data`)
            data = np.asarray(aobj._data, dtype=np.float64)
            if data.ndim == 1:
                data = data[:, None]
            nc = data.shape[1]
            pcm, div = torch.tensor(np.ascontiguousarray(data[s_lo:s_hi].T)).to(k.device), 1.0
        else:
            # ship the samples as stored (int16: 4x fewer bytes than float64) and let the
            # STFT kernel apply the reference's scaling data / (1.1 max|data|)
            if not hasattr(aobj, '_raw'):
                aobj._read_raw()
            raw = aobj._raw if hasattr(aobj._raw, "numpy") else torch.from_numpy(
                np.ascontiguousarray(aobj._raw))
            if raw.dim() == 1:
                raw = raw[:, None]
            nc = raw.shape[1]
            raw = raw[s_lo:s_hi]
            if raw.dtype not in (torch.int16, torch.int32, torch.float32):
                raw = raw.to(torch.float64).t().contiguous()  # planar float64 path
            pcm = raw.to(k.device, non_blocking=True)
            if not hasattr(aobj, '_maxdata'):
                # the scaling factor 1.1 max|x| (audioObject.py:124-126) from a scan of the
                # samples already in HBM (a host scan of a 10-min mixture costs more than its
                # STFT); sharded: every rank scans its own samples, then all-reduce MAX
                peak = torch.zeros(1, dtype=torch.float64, device=k.device)
                k.pcm_peak(pcm, peak)
                if multi:
                    self._comm.allreduce_max(peak)
                aobj._maxdata = np.maximum(1.1 * float(peak.item()), 1e-10)
            div = float(aobj._maxdata)
        if nc < 2 or nc > 4:
            # the reference accepts stereo only (audioModel.py:394, :605); 3 and 4 channels are
            # an extension of this package (general-I kernels, csrc/estep_multi.cu)
            raise AttributeError("Nb channels " + str(nc) + " not implemented yet")
        F = nfft // 2 + 1
        psd = torch.zeros(F, dtype=torch.float64, device=k.device)
        X, _ = _stft.stft_planes(k, pcm, self.tft.window, hop, nfft, self.compute_dtype, psd,
                                 pcm_div=div, frames=frames, sample0=s_lo, L_total=L)
        self.nbFreqsSigRepr, self.nbFramesSigRepr = F, N
        self._Cx = None
        del self.audioObject.data  # like the reference (:288); `_raw`, if any, is kept
        if multi and self._shard == 'time':
            self._comm.allreduce_sum(psd)
        elif multi:
            from .engine import shard_bounds
            lo, hi = shard_bounds(F, self._comm.world)[self._comm.rank]
            # (the engine's rows are a multiple of 32 * world frames long under frequency sharding)
            ldp = -(-N // (32 * self._comm.world)) * (32 * self._comm.world)
            Xs = torch.zeros([X.shape[0], hi - lo, ldp], dtype=X.dtype, device=X.device)
            Xs[:, :, :X.shape[2]] = X[:, lo:hi]
            X = Xs
        self._X = X
        lim = self.noise['ann_PSD_lim']
        if lim[0] is None or lim[1] is None:
            mix_psd = psd.cpu().numpy() / (N * nc)  # mean over frames, then over channels
            if lim[0] is None:
                lim[0] = mix_psd / 100.
            if lim[1] is None:
                lim[1] = mix_psd / 10000.
        if self.noise['sim_ann_opt'] in ('ann'):  # substring test, as in the reference (:324)
            self.noise['PSD'] = lim[0]

    @property
    def Cx(self):
        """Upper triangle of the empirical covariance Cx[n1, n2] = X[n1] conj(X[n2]), n1 <= n2, as
        [nc (nc + 1) / 2, F, N] complex128 in the reference's order (audioModel.py:293-302): a host
        copy built on demand, the kernels read X."""
        if self._Cx is None:
            if self._X is None:
                self.comp_transf_Cx()
            if self._comm is not None and self._comm.world > 1:
                raise NotImplementedError("Cx is not gathered when the model is sharded over GPUs")
            X = self._X[:, :, :self.nbFramesSigRepr].cpu().numpy().astype(np.float64)
            nc = X.shape[0] // 2
            x = [X[2 * c] + 1j * X[2 * c + 1] for c in range(nc)]
            self._Cx = np.array([x[n1] * np.conj(x[n2]) for n1 in range(nc)
                                 for n2 in range(n1, nc)])
        return self._Cx

    # ------------------------------------------------------------------ GEM
    def estim_param_a_post_model(self):
        """Runs `iter_num` GEM iterations on the device; returns the log-likelihoods
        (ref: audioModel.py:330-382)."""
        opt = self.noise['sim_ann_opt']
        if opt not in ('ann', 'no_ann', 'ann_ns_inj'):
            warnings.warn("To add noise to the signal, provide the sim_ann_opt from any of "
                          "'ann', 'no_ann' or 'ann_ns_inj' ")
        import time
        t0 = time.perf_counter()
        eng = self._engine()
        t1 = time.perf_counter()
        logliks = eng.run(self.iter_num, use_graph=self._use_cuda_graph)
        first = eng.vanished_iteration()
        if first is not None:
            # a TW vanished (sum < eps) at iteration `first`: the reference re-draws it with
            # np.random inside renormalize_parameters (:2023-2025).  The device loop never waits
            # for the host, so the run is replayed from the unchanged host parameters (the kernels
            # are deterministic) and synchronises with the host from that iteration on.
            eng = self._engine()
            logliks = eng.run(self.iter_num, careful_from=first)
        t2 = time.perf_counter()
        # (sharded over GPUs: every rank writes back its own rows / frames of the sharded factor;
        # gather_parameters() completes the model on every rank)
        eng.read_model(self.spat_comps, self.spec_comps, gather=False)
        self.noise['PSD'] = eng.noise_psd()
        t3 = time.perf_counter()
        # host-side wall times of the three stages (pack to HBM, GEM loop, unpack to NumPy)
        self._last_engine_stats = {'launches': self._k().launch_count(), 'pack_s': t1 - t0,
                                   'gem_s': t2 - t1, 'unpack_s': t3 - t2}
        return logliks

    def gather_parameters(self):
        """Model sharded over several GPUs (`comm`): after estim_param_a_post_model every rank holds
        its own rows (frequency sharding: FB) or frames (frame sharding: TW) of the sharded
        factor.  This collective call completes the parameter dicts on every rank."""
        if self._comm is None or self._comm.world == 1:
            return
        import torch
        from .engine import shard_bounds
        dev = self._k().device
        for sp in self.spec_comps.values():
            for fac in sp['factor'].values():
                name, axis = ('TW', 1) if self._shard == 'time' else ('FB', 0)
                arr = np.ascontiguousarray(np.asarray(fac[name], dtype=np.float64))
                bounds = shard_bounds(arr.shape[axis], self._comm.world)
                lo, hi = bounds[self._comm.rank]
                smax = max(b[1] - b[0] for b in bounds)
                mine = np.moveaxis(arr, axis, 0)[lo:hi]
                pad = torch.zeros((smax,) + mine.shape[1:], dtype=torch.float64, device=dev)
                pad[:hi - lo] = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
                parts = self._comm.allgather(pad)
                full = torch.cat([p[:b[1] - b[0]] for p, b in zip(parts, bounds)], dim=0)
                fac[name] = np.ascontiguousarray(np.moveaxis(full.cpu().numpy(), 0, axis))

    def GEM_iteration(self):
        """One GEM iteration with the current noise PSD (ref: audioModel.py:384-428)."""
        eng = self._engine(psd_mode='fixed')
        ll = eng.run(1, careful_from=0)
        eng.read_model(self.spat_comps, self.spec_comps)
        return float(ll[0])

    def comp_spat_comp_power(self, spat_comp_ind, spec_comp_ind=[], factor_ind=[]):
        """V = power of one spatial component [F, N] (ref: audioModel.py:430-498)."""
        eng = self._engine(psd_mode='fixed')
        eng.compute_powers(with_G=False)
        if isinstance(eng, GeneralGemEngine):
            # sum over the selected spectral components of this spatial component of the product
            # of the selected factors; an empty list selects everything (Q1, :476-497)
            V = None
            specs = list(spec_comp_ind) if len(spec_comp_ind) else list(range(len(eng.spec)))
            for s_ind in specs:
                sp = eng.spec[s_ind]
                if sp["j"] != spat_comp_ind:
                    continue
                keys = [fc["key"] for fc in sp["fac"]]
                which = list(factor_ind) if len(factor_ind) else keys
                C = None
                for fc in sp["fac"]:
                    if fc["key"] in which:
                        C = fc["P"].clone() if C is None else C * fc["P"]
                if C is None:
                    C = eng.torch.ones_like(sp["fac"][0]["P"])
                V = C if V is None else V + C
            if V is None:
                V = eng.torch.zeros_like(eng.V[0])
            return V[:, :self.nbFramesSigRepr].cpu().numpy().astype(np.float64)
        if (len(factor_ind) and list(factor_ind) != [0]) or (len(spec_comp_ind) and any(
                self.spec_comps[s]['spat_comp_ind'] != spat_comp_ind for s in spec_comp_ind)):
            raise NotImplementedError("factor / foreign spectral-component selection")
        V = eng._gather_f(eng.V[spat_comp_ind], 0)
        return V[:, :self.nbFramesSigRepr].astype(np.float64)

    def retrieve_subsrc_params(self):
        """(spat_comp_powers [Rtot,F,N], mix_matrix [Rtot,2,F], rank_part_ind)
        (ref: audioModel.py:514-578)."""
        eng = self._engine(psd_mode='fixed')
        eng.compute_powers(with_G=False)
        V = eng._gather_f(eng.V, 1)[:, :, :self.nbFramesSigRepr].astype(np.float64)
        A = eng._gather_f(eng.A, 2)
        rank_part_ind = {j: np.array(r) for j, r in enumerate(eng.ranks)}
        powers = np.zeros([eng.R, self.nbFreqsSigRepr, self.nbFramesSigRepr])
        for j, idx in rank_part_ind.items():
            powers[idx] = V[j][None]
        return powers, A, rank_part_ind

    def compute_suff_stat(self, spat_comp_powers, mix_matrix):
        """E-step sufficient statistics for given sub-source powers and mixing vectors
        (ref: audioModel.py:580-764).  Returns (hat_Rxx, hat_Rxs, hat_Rss, hat_Ws,
        loglik) with the reference's shapes; every sub-source is treated as its own
        rank-one component so that hat_Ws is per sub-source."""
        import torch
        if self._X is None:
            self.comp_transf_Cx()
        if self._comm is not None and self._comm.world > 1:
            raise NotImplementedError("compute_suff_stat under frequency sharding: use "
                                      "estim_param_a_post_model")
        k = self._k()
        powers = np.asarray(spat_comp_powers)
        mix = np.asarray(mix_matrix)
        R, F, N = powers.shape
        if mix.shape[1] != 2:
            raise ValueError("Nb channels not supported:" + str(mix.shape[1]))
        eng = GemEngine(k, F, N, dtype=self.compute_dtype)
        eng.set_X_planes(self._X)
        eng.set_noise('fixed', self.noise['PSD'], self.noise['PSD'], self.noise['PSD'])
        A = eng._upload(mix.astype(np.complex128))
        Rss = eng._zeros([F, R, R], torch.complex128)
        Rxs = eng._zeros([F, 2, R], torch.complex128)
        ll_f = eng._zeros([F], torch.float64)
        hat_Ws = np.zeros([R, F, N])
        # The kernel takes up to MAX_COMPS spatial components.  Sub-sources with the SAME power
        # (the rank columns of one source, as retrieve_subsrc_params returns them) may share a
        # component -- Sigma only sees sum_r a_r a_r^H per power -- but hat_Ws is wanted per
        # sub-source: each pass gives every sub-source of some groups its own component and merges
        # the others; hat_Rss, hat_Rxs and the log-likelihood are the same in every pass.
        groups = []
        for r in range(R):
            for g in groups:
                if np.array_equal(powers[g[0]], powers[r]):
                    g.append(r)
                    break
            else:
                groups.append([r])
        MAX_COMPS = 6
        if len(groups) > MAX_COMPS:
            raise NotImplementedError("compute_suff_stat: %d sub-sources with distinct powers "
                                      "(the E-step kernel takes %d)" % (len(groups), MAX_COMPS))
        pending = list(groups)
        while pending:
            chosen, ncomp = [], len(groups)
            for g in list(pending):
                if ncomp + len(g) - 1 <= MAX_COMPS:
                    chosen.append(g)
                    ncomp += len(g) - 1
                    pending.remove(g)
            src, comp_rows, own = [0] * R, [], {}
            for g in groups:
                if any(g is c for c in chosen):
                    for r in g:
                        src[r] = len(comp_rows)
                        own[r] = len(comp_rows)
                        comp_rows.append(r)
                else:
                    for r in g:
                        src[r] = len(comp_rows)
                    comp_rows.append(g[0])
            J = len(comp_rows)
            Vp = np.zeros([J, F, eng.ld])
            Vp[:, :, :N] = powers[comp_rows]
            V = eng._upload(Vp, eng.tdtype)
            hatW = eng._zeros([J, F, eng.ld])
            ws = eng._zeros([(k.estep_workspace_bytes(J, F, N, k.dtype_code(V)) + 7) // 8],
                            torch.float64)
            k.estep_stereo(eng.X, V, A, src, eng.noise, N, hatW, Rss, Rxs, ll_f, ws)
            hw = hatW[:, :, :N].cpu().numpy().astype(np.float64)
            for r, c in own.items():
                hat_Ws[r] = hw[c]
        loglik = -float(ll_f.sum().cpu().item()) / (F * N)
        hat_Rxx = np.mean(self.Cx, axis=-1)
        return (hat_Rxx, Rxs.cpu().numpy(), Rss.cpu().numpy(), hat_Ws, loglik)

    def renormalize_parameters(self):
        """Energy normalisation across A, FB, FW, TW(, TB) (ref: audioModel.py:1980-2040).  A TW
        whose sum fell below eps is re-drawn with np.random like the reference (:2023-2025)."""
        eng = self._engine(psd_mode='fixed')
        eng.flags.zero_()
        eng.totals.zero_()
        eng.sync_redraw = True
        eng.renormalize()
        eng.check_flags()
        eng.read_model(self.spat_comps, self.spec_comps)

    # ------------------------------------------------------------------ M-step, by hand
    def update_mix_matrix(self, hat_Rxs, hat_Rss, mix_matrix, rank_part_ind):
        """Spatial M-step from given statistics (ref: audioModel.py:766-889): `mix_matrix`
        [Rtot, nc, F] is updated IN PLACE and the 'params' of the free spatial components are
        replaced, like the reference does.  (hat_Rxs is left untouched; the reference modifies it
        in place when some instantaneous components are fixed, :810-818.)"""
        import torch
        eng = self._engine(psd_mode='fixed')
        mix_matrix_in = mix_matrix
        mix = np.asarray(mix_matrix)
        if mix.shape != (eng.R, eng.I, self.nbFreqsSigRepr):
            raise ValueError("mix_matrix must be [Rtot, nchannels, F]")
        for j, idx in rank_part_ind.items():
            if list(np.atleast_1d(idx)) != eng.ranks[j]:
                raise NotImplementedError("rank_part_ind differs from retrieve_subsrc_params()")
        eng.A = eng._upload(mix.astype(np.complex128))
        eng.Rss.copy_(eng._upload(np.asarray(hat_Rss, dtype=np.complex128)))
        eng.Rxs.copy_(eng._upload(np.asarray(hat_Rxs, dtype=np.complex128)))
        eng.flags.zero_()
        eng.update_mix()
        eng.check_flags()
        A = eng.A.cpu().numpy()
        if isinstance(mix_matrix_in, np.ndarray) and mix_matrix_in.flags.writeable:
            mix_matrix_in[...] = A if np.iscomplexobj(mix_matrix_in) else A.real
        for j, sc in self.spat_comps.items():
            if sc['frdm_prior'] == 'free':
                if sc['mix_type'] == 'inst':
                    sc['params'] = np.mean(A[eng.ranks[j]], axis=2).T
                else:
                    sc['params'] = np.ascontiguousarray(A[eng.ranks[j]])

    def update_spectral_components(self, hat_W):
        """Spectral M-step for given posterior powers hat_W [nb_spat_comps, F, N]
        (ref: audioModel.py:1469-1978); the factors of `spec_comps` are updated."""
        eng = self._engine(psd_mode='fixed')
        hw = np.asarray(hat_W, dtype=np.float64)
        if hw.shape != (eng.J, self.nbFreqsSigRepr, self.nbFramesSigRepr):
            raise ValueError("hat_W must be [nb_spat_comps, F, N]")
        if eng._sharded():
            raise NotImplementedError("update_spectral_components of a sharded model: use "
                                      "estim_param_a_post_model")
        eng.hatW[:, :, :eng.N] = eng._upload(hw, eng.tdtype)
        eng.compute_powers()
        eng.update_spectral()
        spat_before = {j: sc['params'] for j, sc in self.spat_comps.items()}
        eng.read_model(self.spat_comps, self.spec_comps)
        for j, p in spat_before.items():  # (only the spectral parameters change)
            self.spat_comps[j]['params'] = p

    # ------------------------------------------------------------------ NMF initialisation
    def _mono_power(self):
        """Device plane [F, ldn] float32: mean over the channels of |X_c|^2, the one-channel
        power spectrum the initialisers factorise (ref: audioModel.py:2150-2158)."""
        import torch
        if self._X is None:
            self.comp_transf_Cx()
        if self._comm is not None and self._comm.world > 1:
            raise NotImplementedError("NMF initialisation of a sharded model")
        k = self._k()
        F, N = self.nbFreqsSigRepr, self.nbFramesSigRepr
        ldn = (N + 3) // 4 * 4
        X = self._X if self._X.dtype == torch.float32 else self._X.to(torch.float32)
        out = torch.zeros((F, ldn), dtype=torch.float32, device=k.device)
        k.mono_power(X, out, F, N, ldn)
        return out

    def initialize_all_spec_comps_with_NMF(self, sameInitAll=False, **kwargs):
        """IS-NMF of the one-channel mixture as initial FB / TW of every spectral component
        (ref: audioModel.py:2091-2116)."""
        if sameInitAll:
            return self.initialize_all_spec_comps_with_NMF_same(**kwargs)
        return self.initialize_all_spec_comps_with_NMF_indiv(**kwargs)

    def initialize_all_spec_comps_with_NMF_indiv(self, niter=10, updateFreqBasis=True,
                                                 updateTimeWeight=True, **kwargs):
        """One NMF with the stacked FB / TW of all components as initial point; every component
        receives its slice back (ref: audioModel.py:2118-2183)."""
        from .tools.nmf import NMF_decomp_init
        eps = 1e-10
        nb = [sc['factor'][0]['FB'].shape[1] for sc in self.spec_comps.values()]
        total = int(np.sum(nb))
        FBinit = np.zeros([self.nbFreqsSigRepr, total])
        TWinit = np.zeros([total, self.nbFramesSigRepr])
        for ind, sc in self.spec_comps.items():
            lo = int(np.sum(nb[:ind]))
            FBinit[:, lo:lo + nb[ind]] = sc['factor'][0]['FB']
            TWinit[lo:lo + nb[ind]] = sc['factor'][0]['TW']
        W, H = NMF_decomp_init(SX=self._mono_power(), nbComps=total, niter=niter,
                               verbose=self.verbose, Winit=FBinit, Hinit=TWinit,
                               updateW=updateFreqBasis, updateH=updateTimeWeight,
                               kernels=self._k(), nframes=self.nbFramesSigRepr)
        for ind, sc in self.spec_comps.items():
            lo = int(np.sum(nb[:ind]))
            if updateFreqBasis:
                sc['factor'][0]['FB'] = np.maximum(W[:, lo:lo + nb[ind]], eps)
            if updateTimeWeight:
                sc['factor'][0]['TW'] = np.maximum(H[lo:lo + nb[ind]], eps)
        self.renormalize_parameters()

    def initialize_all_spec_comps_with_NMF_same(self, niter=10, **kwargs):
        """The same W / H (most energetic components first) for every spectral component
        (ref: audioModel.py:2185-2222)."""
        from .tools.nmf import NMF_decomposition
        if not np.all([len(sc['factor']) == 1 for sc in self.spec_comps.values()]):
            raise NotImplementedError("NMF init not implemented for multi factor models.")
        nb = [sc['factor'][0]['FB'].shape[1] for sc in self.spec_comps.values()]
        W, H = NMF_decomposition(SX=self._mono_power(), verbose=self.verbose,
                                 nbComps=int(np.max(nb)), niter=niter, kernels=self._k(),
                                 nframes=self.nbFramesSigRepr)
        order = np.argsort(H.sum(axis=1))[::-1]
        W, H = W[:, order], H[order]
        for sc in self.spec_comps.values():
            ncomp = sc['factor'][0]['FB'].shape[1]
            sc['factor'][0]['FB'][:] = W[:, :ncomp]
            sc['factor'][0]['TW'][:] = H[:ncomp]
        self.renormalize_parameters()

    # ------------------------------------------------------------------ K6
    def separate_spat_comps(self, dir_results=None, suffix=None):
        """One separated (stereo) signal per spatial component (ref: audioModel.py:1063-1086)."""
        spec_comp_ind = {}
        for spat_ind in range(len(self.spat_comps)):
            spec_comp_ind[spat_ind] = []
        for spec_ind, spec_comp in self.spec_comps.items():
            spec_comp_ind[spec_comp['spat_comp_ind']].append(spec_ind)
        self.separate_comps(dir_results=dir_results, spec_comp_ind=spec_comp_ind, suffix=suffix)

    def separate_comps(self, dir_results=None, spec_comp_ind=None, suffix=None):
        """Wiener separation of groups of spectral components, inverse STFT and WAV output
        (ref: audioModel.py:1088-1236).  File names: <dir>/<root>_<n>-<nbSources>[_suffix].wav"""
        pcm = self.separate_comps_pcm(spec_comp_ind)
        if dir_results is None:
            dir_results = '/'.join(self.audioObject.filename.split('/')[:-1])
        if not hasattr(self, "files"):
            self.files = {}
        self.files['spat_comp'] = []
        nbSources = pcm.shape[0]
        fileroot = self.audioObject.filename.split('/')[-1][:-4]
        for n in range(nbSources):
            _suffix = ''
            if suffix is not None and n in suffix:
                _suffix = '_' + suffix[n]
            name = dir_results + '/' + fileroot + '_' + str(n) + '-' + str(nbSources) + \
                _suffix + '.wav'
            self.files['spat_comp'].append(name)
            out = ao.AudioObject(filename=name, mode='w')
            out._data = pcm[n]
            out._maxdata = 1
            out._encoding = 'pcm16'
            out.samplerate = self.audioObject.samplerate
            out._write()

    def _wiener_groups(self, eng, spec_comp_ind, nbSources):
        """Wiener-filtered STFT planes of the output groups.  Like the reference
        (audioModel.py:1141-1166, :1327-1390): the covariance of group n is the sum, over the
        spatial components its spectral components belong to, of R_spat times the power of THOSE
        spectral components; the mixture covariance is the sum over the REQUESTED groups plus the
        noise -- spectral components that are in no group do not enter it."""
        # "virtual sources": one per (group, spatial component) pair
        pairs = []
        for n in range(nbSources):
            for j in sorted(set(self.spec_comps[s]['spat_comp_ind'] for s in spec_comp_ind[n])):
                pairs.append((n, j, [s for s in spec_comp_ind[n]
                                     if self.spec_comps[s]['spat_comp_ind'] == j]))
        full_cover = sorted(s for n in range(nbSources) for s in spec_comp_ind[n]) == \
            sorted(self.spec_comps.keys()) and len(set(j for _, j, _ in pairs)) == len(pairs)
        if full_cover:
            # every spectral component in exactly one group, every spatial component in one group
            group_of_src = [-1] * eng.J
            for n, j, _ in pairs:
                group_of_src[j] = n
            return eng.wiener(group_of_src, nbSources)
        Rv = sum(len(eng.ranks[j]) for _, j, _ in pairs)
        if len(pairs) > 6 or Rv > 16:
            raise NotImplementedError("separate_comps: more than 6 (group, spatial component) "
                                      "pairs or 16 sub-sources in one call")
        eng.compute_powers(with_G=False)
        torch = eng.torch
        Vv = eng._zeros([len(pairs), eng.F, eng.ld])
        rows, src_of_sub, group_of_src = [], [], []
        for v, (n, j, specs) in enumerate(pairs):
            for s in specs:  # power of the selected spectral components of this spatial component
                Vv[v] += eng.component_power(s)
            rows.extend(eng.ranks[j])
            src_of_sub.extend([v] * len(eng.ranks[j]))
            group_of_src.append(n)
        Av = eng.A[torch.tensor(rows, device=eng.dev)].contiguous()
        return eng.wiener_custom(Vv, Av, src_of_sub, group_of_src, nbSources)

    def separate_comps_pcm(self, spec_comp_ind=None):
        """int16 [nbSources, L, 2]: the separated signals of `separate_comps` before they are
        written (device: Wiener filter K6 + inverse STFT with overlap-add)."""
        import torch
        nc = self.audioObject.channels
        if nc < 2 or nc > 4:
            raise NotImplementedError()
        if spec_comp_ind is None:
            spec_comp_ind = {s: [s] for s in range(len(self.spec_comps))}
        nbSources = len(spec_comp_ind)
        eng = self._engine(psd_mode='fixed')
        Y = self._wiener_groups(eng, spec_comp_ind, nbSources)
        if eng._sharded():
            # the inverse STFT needs every frequency of every frame: gather the shards (every
            # rank then inverts the whole signal; separation is a one-off, not the GEM loop)
            if eng.shard == 'freq':
                Yh = eng._gather_f(Y, 1)
            else:
                Yn = eng._gather_n(Y[:, :, :eng.N].contiguous(), 2)
                Yh = np.zeros(Yn.shape[:2] + ((Yn.shape[2] + 31) // 32 * 32,), dtype=Yn.dtype)
                Yh[:, :, :Yn.shape[2]] = Yn
            Y = torch.tensor(Yh).to(eng.dev)
        L = self.audioObject.nframes
        maxdata = float(self.audioObject._ensure_maxdata())
        hop, nfft = self.sig_repr_params['hopsize'], self.sig_repr_params['fsize']
        _, pcm = _stft.istft_planes(self._k(), Y, self.nbFramesSigRepr, self.tft.synthWindow,
                                    self.tft.window, hop, nfft, length=L, maxdata=maxdata)
        # pcm: [L, nc*nbSources] with signal index = nc*source + channel; the re-ordering to
        # [source, L, channel] happens on the device (a strided host copy of the 423 MB of a
        # 10-minute 4-source separation took 0.3 s)
        return pcm.view(L, nbSources, nc).permute(1, 0, 2).contiguous().cpu().numpy()


class MultiChanNMFInst_FASST(FASST):
    """Multichannel NMF, instantaneous mixing (ref: audioModel.py:2296-2420)."""

    def __init__(self, audio, nbComps=3, nbNMFComps=4, spatial_rank=2, **kwargs):
        super(MultiChanNMFInst_FASST, self).__init__(audio=audio, **kwargs)
        self.comp_transf_Cx()
        self.nbComps = nbComps
        self.nbNMFComps = nbNMFComps
        self.rank = np.atleast_1d(spatial_rank)
        if self.rank.size < self.nbComps:
            self.rank = [self.rank[0], ] * self.nbComps
        self._initialize_structures()

    def _initialize_structures(self):
        """Random initial parameters with the reference's np.random call order
        (ref: audioModel.py:2349-2393), then renormalisation on the device."""
        nc = self.audioObject.channels
        self.spat_comps = {}
        self.spec_comps = {}
        for j in range(self.nbComps):
            params = np.random.randn(nc, self.rank[j])
            if nc == 2:  # sources spread evenly over the stereo field
                ang = (j + 1) * np.pi / (2. * (self.nbComps + 1))
                params = np.array([np.sin(ang) + np.random.randn(self.rank[j]) * np.sqrt(0.01),
                                   np.cos(ang) + np.random.randn(self.rank[j]) * np.sqrt(0.01)])
            self.spat_comps[j] = {'time_dep': 'indep', 'mix_type': 'inst',
                                  'frdm_prior': 'free', 'params': params}
            factor = {
                'FB': 0.75 * np.abs(np.random.randn(self.nbFreqsSigRepr, self.nbNMFComps)) + 0.25,
                'FW': np.eye(self.nbNMFComps),
                'TW': 0.75 * np.abs(np.random.randn(self.nbNMFComps, self.nbFramesSigRepr)) + 0.25,
                'TB': [],
                'FB_frdm_prior': 'free', 'FW_frdm_prior': 'fixed',
                'TW_frdm_prior': 'free', 'TB_frdm_prior': [],
                'TW_constr': 'NMF',
            }
            factor['TW'] = self._host_param(factor['TW'])
            self.spec_comps[j] = {'spat_comp_ind': j, 'factor': {0: factor}}
        self.renormalize_parameters()

    def _host_param(self, arr):
        """The large parameter matrices this class creates itself (TW: K x N float64, 13 MB per
        source for a 10-minute mixture) live in page-locked host memory: every public method moves
        them to HBM and back IN PLACE, and the DMA from / into pinned pages runs 2-4 x faster
        than through the driver's staging of pageable memory (4.6 / 2.7 ms against 1.0 / 1.0 ms for
        the four matrices of configs[1], scripts/micro/host_register.py).  They remain plain NumPy
        arrays (the pinned tensor is their base); arrays a user puts in their place are taken as
        they are.  PYFASST_PINNED_PARAMS=0 keeps pageable memory."""
        if arr.nbytes < (1 << 20) or self._kernels is not None and getattr(self._kernels, 'name', '') != 'cuda':
            return arr
        if os.environ.get('PYFASST_PINNED_PARAMS', '1') == '0':
            return arr
        try:
            import torch
            if not torch.cuda.is_available():
                return arr
            t = torch.empty(arr.shape, dtype=torch.float64, pin_memory=True)
            out = t.numpy()
            out[...] = arr
            return out
        except Exception:  # noqa: BLE001 -- page-locking is an optimisation only
            return arr

    def setSpecCompFB(self, compNb, FB, FB_frdm_prior='fixed'):
        """Sets the frequency basis of one spectral component (ref: audioModel.py:2395-2420)."""
        speccomp = self.spec_comps[compNb]['factor'][0]
        if self.nbFreqsSigRepr != FB.shape[0]:
            raise AttributeError("Size of provided FB is not consistent with inner attributes")
        speccomp['FB'] = np.copy(FB)
        ncomp = FB.shape[1]
        speccomp['FW'] = np.eye(ncomp)
        speccomp['TW'] = 0.75 * np.abs(np.random.randn(ncomp, self.nbFramesSigRepr)) + 0.25
        speccomp['FB_frdm_prior'] = FB_frdm_prior


class MultiChanNMFConv(MultiChanNMFInst_FASST):
    """Multichannel NMF, convolutive mixing (ref: audioModel.py:2422-2508)."""

    def __init__(self, audio, nbComps=3, nbNMFComps=4, spatial_rank=2, **kwargs):
        super(MultiChanNMFConv, self).__init__(audio=audio, nbComps=nbComps,
                                               nbNMFComps=nbNMFComps,
                                               spatial_rank=spatial_rank, **kwargs)

    def makeItConvolutive(self):
        """Instantaneous -> convolutive: the mixing vector is copied to every frequency
        (ref: audioModel.py:2488-2508)."""
        nc = self.audioObject.channels
        for nspat, (spat_ind, spat_comp) in enumerate(self.spat_comps.items()):
            if spat_comp['mix_type'] != 'inst':
                warnings.warn("Spatial component %d " % spat_ind +
                              "already not instantaneous, skipping...")
                continue
            spat_comp['mix_type'] = 'conv'
            inst = np.asarray(spat_comp['params'])
            spat_comp['params'] = np.zeros([self.rank[nspat], nc, self.nbFreqsSigRepr],
                                           dtype=complex)
            spat_comp['params'][:] = inst.T[:, :, None]

    def initializeConvParams(self, initMethod='rand'):
        """Random convolutive mixing parameters (ref: audioModel.py:2224-2294, 'rand'
        branch).  The DEMIX initialisation ('demix', the reference's default) is out of
        scope of the accelerated path."""
        nc = self.audioObject.channels
        if 'rand' not in initMethod:
            if initMethod == 'demix':
                raise NotImplementedError("DEMIX initialisation is not part of this package")
            raise ValueError("Init method not implemented.")
        for spat_ind, spat_comp in self.spat_comps.items():
            if spat_comp['mix_type'] != 'inst':
                warnings.warn("Spatial component %d " % spat_ind +
                              "already not instantaneous, overwriting...")
            spat_comp['mix_type'] = 'conv'
        A = (np.random.randn(len(self.spat_comps), self.nbFreqsSigRepr, nc)
             + 1j * np.random.randn(len(self.spat_comps), self.nbFreqsSigRepr, nc))
        for nspat, (spat_ind, spat_comp) in enumerate(self.spat_comps.items()):
            spat_comp['params'] = np.zeros([self.rank[nspat], nc, self.nbFreqsSigRepr],
                                           dtype=complex)
            for r in range(self.rank[nspat]):
                spat_comp['params'][r] = A[spat_ind].T


class multiChanSourceF0Filter(FASST):
    """Multichannel source/filter model (ref: audioModel.py:2551-3014): `nbComps - 1` sources
    with a two-factor spectral component -- glottal-comb dictionary x smooth filters -- and one
    residual NMF component.  The dictionary WF0 is generated on the GPU
    (SeparateLeadStereo.separateLeadFunctions.generate_WF0_TR_chirped) and SHARED by all the
    sources, like the reference's `self.sourceFreqComps` (quirk Q11: the renormalisation rescales
    the shared array once per source).  Estimation runs on GeneralGemEngine.

    `sparsity` (the re-weighting of the source activations after every iteration,
    reweigh_sparsity_constraint :2981-3014) runs on the device inside the estimation loop.
    Not here: `initSpecCompsWithLabelAndFiles` (needs the external gmm-gsmm module) and
    `initializeFreeMats`."""

    def __init__(self, audio, nbComps=3, nbNMFResComps=1, nbFilterComps=20, nbFilterWeigs=[4, ],
                 minF0=39, maxF0=2000, minF0search=80, maxF0search=800, stepnoteF0=16,
                 chirpPerF0=1, spatial_rank=1, sparsity=None, **kwargs):
        from .SeparateLeadStereo import separateLeadFunctions as slf
        super(multiChanSourceF0Filter, self).__init__(audio=audio, **kwargs)
        self.comp_transf_Cx()
        self.sourceParams = {'minF0': minF0, 'maxF0': maxF0, 'stepnoteF0': stepnoteF0,
                             'chirpPerF0': chirpPerF0, 'minF0search': minF0search,
                             'maxF0search': maxF0search}
        self.nbComps = nbComps
        self.nbNMFResComps = nbNMFResComps
        self.nbFilterComps = nbFilterComps
        if len(nbFilterWeigs) < self.nbComps - 1:
            self.nbFilterWeigs = [nbFilterWeigs[0], ] * self.nbComps
        else:
            self.nbFilterWeigs = nbFilterWeigs
        self.spatial_rank = np.atleast_1d(spatial_rank)
        if self.spatial_rank.size < self.nbComps:
            self.spatial_rank = [self.spatial_rank[0], ] * self.nbComps
        # the source dictionary is shared among all the components (:2613-2635)
        self.F0Table, WF0, _ = slf.generate_WF0_TR_chirped(
            transform=self.tft, minF0=minF0, maxF0=maxF0, stepNotes=stepnoteF0, Ot=0.5,
            perF0=chirpPerF0, depthChirpInSemiTone=0.5, loadWF0=True, verbose=self.verbose,
            kernels=self._kernels)
        WF0 = np.array(WF0)
        for n in range(WF0.shape[1]):  # patterns in low-energy bins are set to eps (:2624-2627)
            WF0[WF0[:, n] < WF0[:, n].max() * 1e-4, n] = eps
        self.sourceFreqComps = np.ascontiguousarray(
            np.hstack([WF0[:self.nbFreqsSigRepr], np.vstack(np.ones(self.nbFreqsSigRepr))]))
        self.nbSourceComps = self.sourceFreqComps.shape[1]
        self.sourceFreqWeights = np.eye(self.nbSourceComps)
        self.filterFreqComps = slf.generateHannBasis(
            numberFrequencyBins=self.nbFreqsSigRepr, sizeOfFourier=self.sig_repr_params['fsize'],
            Fs=self.audioObject.samplerate, frequencyScale='linear',
            numberOfBasis=self.nbFilterComps)
        self.sparsity = sparsity
        self._initialize_structures()

    def _initialize_structures(self, seed=None):
        """ref: audioModel.py:2650-2772, same np.random call order."""
        np.random.seed(seed)
        self.rank = self.spatial_rank
        nc = self.audioObject.channels
        rnd = lambda *shape: 0.75 * np.abs(np.random.randn(*shape)) + 0.25
        self.spat_comps, self.spec_comps = {}, {}
        for j in range(self.nbComps - 1):
            params = np.random.randn(nc, self.rank[j])
            if nc == 2:
                ang = (j + 1) * np.pi / (2. * self.nbComps)
                params = np.array([np.sin(ang) + np.random.randn(self.rank[j]) * np.sqrt(0.01),
                                   np.cos(ang) + np.random.randn(self.rank[j]) * np.sqrt(0.01)])
            self.spat_comps[j] = {'time_dep': 'indep', 'mix_type': 'inst',
                                  'frdm_prior': 'free', 'params': params}
            source = {'FB': self.sourceFreqComps, 'FW': self.sourceFreqWeights,
                      'TW': rnd(self.nbSourceComps, self.nbFramesSigRepr), 'TB': [],
                      'FB_frdm_prior': 'fixed', 'FW_frdm_prior': 'fixed',
                      'TW_frdm_prior': 'free', 'TB_frdm_prior': [], 'TW_constr': 'NMF'}
            filt = {'FB': self.filterFreqComps,
                    'FW': rnd(self.nbFilterComps, self.nbFilterWeigs[j]),
                    'TW': rnd(self.nbFilterWeigs[j], self.nbFramesSigRepr), 'TB': [],
                    'FB_frdm_prior': 'fixed', 'FW_frdm_prior': 'free',
                    'TW_frdm_prior': 'free', 'TB_frdm_prior': [], 'TW_constr': 'NMF'}
            self.spec_comps[j] = {'spat_comp_ind': j, 'factor': {0: source, 1: filt}}
        # residual component: single-factor NMF (:2709-2745)
        self.resSpatialRank = self.rank[-1]
        j = self.nbComps - 1
        self.spat_comps[j] = {'time_dep': 'indep', 'mix_type': 'inst', 'frdm_prior': 'free',
                              'params': np.random.randn(nc, self.resSpatialRank)}
        res = {'FB': rnd(self.nbFreqsSigRepr, self.nbNMFResComps),
               'FW': np.eye(self.nbNMFResComps),
               'TW': rnd(self.nbNMFResComps, self.nbFramesSigRepr), 'TB': [],
               'FB_frdm_prior': 'free', 'FW_frdm_prior': 'fixed',
               'TW_frdm_prior': 'free', 'TB_frdm_prior': [], 'TW_constr': 'NMF'}
        self.spec_comps[j] = {'spat_comp_ind': j, 'factor': {0: res}}
        # sparsity: median-filter length of the re-weighting of the source activations, per
        # component or one value for all (ref: audioModel.py:2747-2770)
        sparsity = self.sparsity
        for j in range(self.nbComps):
            if sparsity is None or len(sparsity) not in (1, self.nbComps):
                self.spec_comps[j]['sparsity'] = False
            elif len(sparsity) == self.nbComps:
                self.spec_comps[j]['sparsity'] = sparsity[j]
            else:
                self.spec_comps[j]['sparsity'] = sparsity[0]
        self.renormalize_parameters()

    def setSpecCompFB(self, compNb, FB, FB_frdm_prior='fixed'):
        """ref: audioModel.py:2858-2876"""
        speccomp = self.spec_comps[compNb]['factor'][0]
        if self.nbFreqsSigRepr != FB.shape[0]:
            raise AttributeError("Size of provided FB is not consistent with inner attributes")
        speccomp['FB'] = np.copy(FB)
        ncomp = FB.shape[1]
        speccomp['FW'] = np.eye(ncomp)
        speccomp['TW'] = 0.75 * np.abs(np.random.randn(ncomp, self.nbFramesSigRepr)) + 0.25
        speccomp['FB_frdm_prior'] = FB_frdm_prior

    def makeItConvolutive(self):
        """ref: audioModel.py:2910-2931"""
        nc = self.audioObject.channels
        for nspat, (spat_ind, spat_comp) in enumerate(self.spat_comps.items()):
            if spat_comp['mix_type'] != 'inst':
                warnings.warn("Spatial component %d " % spat_ind +
                              "already not instantaneous, skipping...")
                continue
            spat_comp['mix_type'] = 'conv'
            inst = np.asarray(spat_comp['params'])
            spat_comp['params'] = np.zeros([self.rank[nspat], nc, self.nbFreqsSigRepr],
                                           dtype=complex)
            spat_comp['params'][:] = np.atleast_2d(inst.T)[:, :, None]
