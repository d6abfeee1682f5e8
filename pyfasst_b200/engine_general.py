"""GEM engine for general factor structures -- source/filter models inside FASST.

`GemEngine` (engine.py) covers the structures of MultiChanNMFInst_FASST / MultiChanNMFConv: one
single-factor NMF component per source, fixed FW, K <= a few tens.  This subclass runs the same
E-step and spatial M-step, and restates the *general* branch of
FASST.update_spectral_components (pyfasst/audioModel.py:1469-1727), comp_spat_comp_power
(:430-498) and renormalize_parameters (:1980-2040) for

  * several factors per spectral component (power = product of the factor powers, :486-494),
    with the true `other_fact_power` of the other factors (:1513-1516),
  * several spectral components per spatial component (:476-497),
  * free FW (:1577-1631), fixed or free FB / TW, large dictionaries (the 1093 glottal combs of
    multiChanSourceF0Filter, :2551-2760),
  * parameter arrays shared between components (the reference renormalises the shared
    dictionary object in place once per component, quirk Q11).

Every contraction with an F x N operand is one launch of the tensor-core GEMM
(csrc/gemm_tc.cu, float32 3xTF32); the planes hat_W / P^2 * O and O / P are formed once per
updated matrix by csrc/gemfac.cu.  float32 planes only on the GPU; one GPU (no sharding);
discrete-state TW constraints raise NotImplementedError.  Also here: time-blob factors
(H = TW TB, the TB update :1931-1978 and its renormalisation :2026-2030) and the correlation
penalty `lambdaCorr` (:1484-1703).

The order of operations is the reference's Gauss-Seidel order: components in key order, factors
in key order, FB -> FW -> TW inside a factor; the power P is recomputed from the current
parameters before every update, `other` once per factor (quirks Q1-Q3 of SURVEY.md 7.4).
"""
import numpy as np

from .engine import EPS, GemEngine, _round_up


class GeneralGemEngine(GemEngine):
    def __init__(self, *args, **kwargs):
        super(GeneralGemEngine, self).__init__(*args, **kwargs)
        if self._sharded():
            raise NotImplementedError("general factor structures run on one GPU")
        self._use_streams = False  # the factor chains of one component depend on each other

    # ------------------------------------------------------------------ model
    def _padded(self, arr, rows, cols):
        """Host matrix -> zero-padded device matrix [rows, cols] of the plane type."""
        t = self._zeros([rows, cols])
        a = np.asarray(arr, dtype=np.float64)
        t[:a.shape[0], :a.shape[1]] = self._upload(a, self.tdtype)
        return t

    def _set_spectral(self, spec_comps):
        S = len(spec_comps)
        if sorted(spec_comps.keys()) != list(range(S)):
            raise ValueError("spec_comps keys must be 0..S-1")
        if self.k.device.type == "cuda" and self.tdtype != self.torch.float32:
            raise NotImplementedError("general factor structures: compute_dtype='float32' only "
                                      "(the contractions run on the tf32 tensor-core GEMM)")
        shared = {}  # id(host array) -> device matrix: shared dictionaries stay shared (Q11)
        self.spec = []
        for s in range(S):
            j = spec_comps[s]["spat_comp_ind"]
            if not 0 <= j < self.J:
                raise ValueError("spec_comps[%d]['spat_comp_ind'] = %r" % (s, j))
            facs = []
            keys = sorted(spec_comps[s]["factor"].keys())
            for fi in keys:
                fac = spec_comps[s]["factor"][fi]
                if fac.get("TW_constr", "NMF") != "NMF":
                    raise NotImplementedError("discrete-state TW constraints (GMM/HMM)")
                FB, FW, TW = (np.asarray(fac[m]) for m in ("FB", "FW", "TW"))
                has_tb = len(fac["TB"]) > 0
                TB = np.asarray(fac["TB"]) if has_tb else None
                # with time blobs the activations are H = TW TB: TW [Kw, L], TB [L, N] (:473-476)
                L = TB.shape[0] if has_tb else self.N_total
                if FB.shape[0] != self.F_total or TW.shape[1] != L or \
                        FW.shape != (FB.shape[1], TW.shape[0]) or \
                        (has_tb and TB.shape[1] != self.N_total):
                    raise ValueError("inconsistent factor shapes FB%s FW%s TW%s TB%s"
                                     % (FB.shape, FW.shape, TW.shape, np.shape(fac["TB"])))
                Kb, Kw = FW.shape
                Kb4, Kw4, L4 = _round_up(Kb, 4), _round_up(Kw, 4), _round_up(L, 4)
                ent = {"key": fi, "Kb": Kb, "Kw": Kw, "Kb4": Kb4, "Kw4": Kw4,
                       "has_tb": has_tb, "L": L, "L4": L4,
                       "FB_free": fac["FB_frdm_prior"] == "free",
                       "FW_free": fac["FW_frdm_prior"] == "free",
                       "TW_free": fac["TW_frdm_prior"] == "free",
                       "TB_free": has_tb and fac["TB_frdm_prior"] == "free"}
                mats = [("FB", fac["FB"], self.F, Kb4), ("FW", fac["FW"], Kb4, Kw4),
                        ("TW", fac["TW"], Kw4, L4 if has_tb else self.ld)]
                if has_tb:
                    mats.append(("TB", fac["TB"], L4, self.ld))
                for name, arr, rows, cols in mats:
                    key = id(arr)
                    if key not in shared:
                        shared[key] = self._padded(arr, rows, cols)
                    ent[name] = shared[key]
                    ent[name + "_host"] = arr
                ent["W"] = self._zeros([self.F, Kw4])    # FB FW
                # the activations H [Kw, N]: TW itself, or TW TB with time blobs
                ent["H"] = self._zeros([Kw4, self.ld]) if has_tb else ent["TW"]
                ent["G"] = self._zeros([Kb4, self.ld])   # FW H (FB update)
                ent["P"] = self._zeros([self.F, self.ld])  # power of the factor
                facs.append(ent)
            self.spec.append({"j": j, "fac": facs, "sparsity": spec_comps[s].get("sparsity"),
                              # power of the component = product of its factor powers
                              "C": facs[0]["P"] if len(facs) == 1 else self._zeros([self.F, self.ld])})
        self.by_src = [[sp for sp in self.spec if sp["j"] == j] for j in range(self.J)]
        if any(len(b) == 0 for b in self.by_src):
            raise ValueError("every spatial component needs at least one spectral component")

    def _alloc_spectral(self):
        torch = self.torch
        F, ld = self.F, self.ld
        f64 = torch.float64
        facs = [fc for sp in self.spec for fc in sp["fac"]]
        Kb4 = max(fc["Kb4"] for fc in facs)
        Kw4 = max(fc["Kw4"] for fc in facs)
        self.other = self._zeros([F, ld])        # other_fact_power of the factor being updated
        self.planes = self._zeros([F, 2 * ld])   # (hat_W / P^2 * O | O / P)
        self.tnd = self._zeros([2 * F * max(Kb4, Kw4)])   # planes contracted over the frames
        self.knd = self._zeros([Kw4, 2 * ld])            # planes contracted over the frequencies
        self.fwnd = self._zeros([2, Kb4, Kw4])
        L4 = max([fc["L4"] for fc in facs if fc["has_tb"]] or [4])
        if any(fc["has_tb"] for fc in facs):
            self.tbnd = self._zeros([2 * F * L4])                 # planes TB^T (TW update)
            self.twnd = self._zeros([2, Kw4, L4])
            self.wl = self._zeros([F, L4])                        # FB FW TW (TB update)
            self.lnd = self._zeros([L4, 2 * ld])
            self.lvec = self._zeros([L4], f64)
        if self.lambdaCorr > 0:
            self.ptot = self._zeros([F, ld])
            self.pminus = self._zeros([F, ld])
        nb = self.k.gemm_splitk_workspace_bytes(2 * F, max(Kb4, Kw4, max([fc["L4"] for fc in facs if fc["has_tb"]] or [4])), ld)
        self.gemm_ws = self._zeros([(nb + 3) // 4], torch.float32)
        n = len(facs)
        self.colmax = self._zeros([n, Kb4], f64)
        self.wcol = self._zeros([n, Kb4], f64)
        self.w2 = self._zeros([n, Kw4], f64)
        self._totals2 = self._zeros([1, n], f64)
        self.totals = self._totals2[0]           # sum of every TW after its rescaling
        self.gcount = self._f64(np.array([fc["Kw"] * fc["L"] for fc in facs], dtype=np.float64))
        self.gvec = self._zeros([Kw4], f64)
        self.sparse_work = self._zeros([2 * self.N], f64)
        # sparsity re-weighting runs inside estim_param_a_post_model only (audioModel.py:2933-2979)
        self.sparsity_enabled = False

    # ------------------------------------------------------------------ powers
    def _gemm(self, A, B, C, M, N, K, transA=False, transB=False, splitk=False):
        self.k.gemm_view(A, B, C, M, N, K, transA=transA, transB=transB,
                         workspace=self.gemm_ws if splitk else None)

    def _refresh_factor(self, fc):
        """W = FB FW and the factor's power P = W TW from the current matrices."""
        if fc["Kb4"] * fc["Kw4"] <= 4096:
            self.k.small_matmul(fc["FB"], fc["FW"], fc["W"])
        else:  # the weights of a large dictionary: tensor-core GEMM
            self._gemm(fc["FB"], fc["FW"], fc["W"], self.F, fc["Kw4"], fc["Kb4"])
        if fc["has_tb"]:  # H = TW TB
            self._gemm(fc["TW"], fc["TB"], fc["H"], fc["Kw4"], self.ld, fc["L4"])
        if fc["Kw"] <= 32:
            self.k.spec_power(fc["W"][:, :fc["Kw"]], fc["H"][:fc["Kw"]], fc["P"], self.N, False)
        else:
            self._gemm(fc["W"], fc["H"], fc["P"], self.F, self.ld, fc["Kw4"])

    def _refresh_comp(self, sp):
        """Power of one spectral component: product over its factors (audioModel.py:486-494)."""
        facs = sp["fac"]
        if len(facs) == 1:
            return
        self.k.mul_planes(facs[0]["P"], facs[1]["P"], sp["C"], self.N)
        for fc in facs[2:]:
            self.k.mul_planes(sp["C"], fc["P"], sp["C"], self.N)

    def _refresh_src(self, j):
        """V_j = sum of the powers of the spectral components of source j (:476-497)."""
        comps = self.by_src[j]
        self.k.mul_planes(comps[0]["C"], None, self.V[j], self.N)
        for sp in comps[1:]:
            self.k.mul_planes(sp["C"], None, self.V[j], self.N, accumulate=True)

    def compute_powers(self, with_G=True):
        for sp in self.spec:
            for fc in sp["fac"]:
                self._refresh_factor(fc)
            self._refresh_comp(sp)
        for j in range(self.J):
            self._refresh_src(j)

    def component_power(self, s):
        return self.spec[s]["C"]

    # ------------------------------------------------------------------ spectral M-step
    def _other(self, sp, fi):
        """other_fact_power, computed once per factor (audioModel.py:1511-1516).  Q1: for a
        single-factor component the empty list of other factors means ALL factors, i.e. its
        own (not yet updated) power."""
        facs = sp["fac"]
        if len(facs) == 1:
            self.k.mul_planes(facs[0]["P"], None, self.other, self.N)
            return
        rest = [fc for i, fc in enumerate(facs) if i != fi]
        self.k.mul_planes(rest[0]["P"], rest[1]["P"] if len(rest) > 1 else None, self.other, self.N)
        for fc in rest[2:]:
            self.k.mul_planes(self.other, fc["P"], self.other, self.N)

    def _ratio_planes(self, hatW, P):
        if self.lambdaCorr > 0:
            self.k.gem_ratio_planes(hatW, P, self.other, self.planes, self.N, self.ptot,
                                    self.pminus, self.lambdaCorr)
        else:
            self.k.gem_ratio_planes(hatW, P, self.other, self.planes, self.N)

    def update_spectral(self):
        k, F, N, ld = self.k, self.F, self.N, self.ld
        for sp in self.spec:
            j = sp["j"]
            if self.lambdaCorr > 0:
                # the powers the correlation penalty is built from, once per spectral component
                # with the parameters as they are now (audioModel.py:1484-1508).  The reference
                # clamps Pminus at eps only when ALL its entries are >= 0 (:1502-1508, a debug
                # leftover): Ptot >= max(V_j, eps) always holds here, so the clamp applies.
                self.compute_powers()
                k.corr_planes(self.V, j, self.ptot, self.pminus, N, True)
            for fi, fc in enumerate(sp["fac"]):
                if not (fc["FB_free"] or fc["FW_free"] or fc["TW_free"] or fc["TB_free"]):
                    continue
                Kb, Kw, Kb4, Kw4 = fc["Kb"], fc["Kw"], fc["Kb4"], fc["Kw4"]
                # a dictionary shared with a component updated earlier may have changed
                self._refresh_factor(fc)
                self._refresh_comp(sp)
                self._other(sp, fi)
                if fc["FB_free"]:
                    # Q3: the power of ALL the spectral components of the source (:1521-1523)
                    self._refresh_src(j)
                    self._ratio_planes(self.hatW[j], self.V[j])
                    if Kb4 * Kw4 <= 4096:                                # (FW H), :1531-1540
                        k.small_matmul(fc["FW"], fc["H"], fc["G"])
                    else:
                        self._gemm(fc["FW"], fc["H"], fc["G"], Kb4, ld, Kw4)
                    T = self.tnd[:2 * F * Kb4].view(2 * F, Kb4)
                    self._gemm(self.planes.view(2 * F, ld), fc["G"], T, 2 * F, Kb4, ld,
                               transB=True, splitk=True)
                    Tv = T.view(F, 2 * Kb4)
                    k.mult_update_same(fc["FB"], Tv[:, :Kb4], Tv[:, Kb4:], F, Kb, self.omega)
                    self._refresh_factor(fc)
                    self._refresh_comp(sp)
                if fc["FW_free"]:
                    # FB^T [(planes) H^T]  (:1577-1631)
                    self._ratio_planes(self.hatW[j], sp["C"])
                    T = self.tnd[:2 * F * Kw4].view(2 * F, Kw4)
                    self._gemm(self.planes.view(2 * F, ld), fc["H"], T, 2 * F, Kw4, ld,
                               transB=True, splitk=True)
                    Tv = T.view(F, 2 * Kw4)
                    for h in range(2):
                        self._gemm(fc["FB"], Tv[:, h * Kw4:(h + 1) * Kw4],
                                   self.fwnd[h, :Kb4, :Kw4], Kb4, Kw4, F, transA=True)
                    k.mult_update_same(fc["FW"], self.fwnd[0, :Kb4, :Kw4], self.fwnd[1, :Kb4, :Kw4],
                                       Kb, Kw, self.omega)
                    self._refresh_factor(fc)
                    self._refresh_comp(sp)
                if fc["TW_free"]:
                    self._ratio_planes(self.hatW[j], sp["C"])
                    if fc["has_tb"]:
                        # (FB FW)^T [(planes) TB^T]  (:1668-1689)
                        L, L4 = fc["L"], fc["L4"]
                        T = self.tbnd[:2 * F * L4].view(2 * F, L4)
                        self._gemm(self.planes.view(2 * F, ld), fc["TB"], T, 2 * F, L4, ld,
                                   transB=True, splitk=True)
                        Tv = T.view(F, 2 * L4)
                        for h in range(2):
                            self._gemm(fc["W"], Tv[:, h * L4:(h + 1) * L4],
                                       self.twnd[h, :Kw4, :L4], Kw4, L4, F, transA=True)
                        k.mult_update_same(fc["TW"], self.twnd[0, :Kw4, :L4], self.twnd[1, :Kw4, :L4],
                                           Kw, L, self.omega)
                    else:
                        # (FB FW)^T (planes)  (:1690-1727)
                        C = self.knd[:Kw4]
                        self._gemm(fc["W"], self.planes, C, Kw4, 2 * ld, F, transA=True)
                        k.mult_update_same(fc["TW"], C[:, :ld], C[:, ld:], Kw, N, self.omega)
                    self._refresh_factor(fc)
                    self._refresh_comp(sp)
                if fc["TB_free"]:
                    # (FB FW TW)^T (planes)  (:1931-1978)
                    L, L4 = fc["L"], fc["L4"]
                    self._ratio_planes(self.hatW[j], sp["C"])
                    self._gemm(fc["W"], fc["TW"], self.wl[:, :L4], F, L4, Kw4)
                    C = self.lnd[:L4]
                    self._gemm(self.wl[:, :L4], self.planes, C, L4, 2 * ld, F, transA=True)
                    k.mult_update_same(fc["TB"], C[:, :ld], C[:, ld:], L, N, self.omega)
                    self._refresh_factor(fc)
                    self._refresh_comp(sp)

    # ------------------------------------------------------------------ renormalisation
    def renormalize(self):
        """renormalize_parameters (audioModel.py:1980-2040): the energy of the mixing parameters
        goes into FB of the first factor, the column maxima of FB into FW, the column means of
        FW into TW, and the mean of TW into FB of the next factor (the last factor keeps it).
        Shared matrices are rescaled in place once per component that refers to them (Q11)."""
        k, torch = self.k, self.torch
        k.spat_energy(self.A, self.src_of_sub, self.J, self.sums)
        k.spat_scale(self.A, self.src_of_sub, self.sums, self.counts)
        i = 0
        for sp in self.spec:
            nfac = len(sp["fac"])
            for fi, fc in enumerate(sp["fac"]):
                Kb, Kw = fc["Kb"], fc["Kw"]
                FB, FW, TW = fc["FB"][:, :Kb], fc["FW"][:Kb, :Kw], fc["TW"][:Kw]
                if fc["has_tb"]:
                    TW = fc["TW"][:Kw, :fc["L"]]
                if fi == 0:
                    k.fb_scale_colmax(FB, self.sums, self.counts, sp["j"], self.colmax[i])
                else:  # global_energy = mean of the previous factor's TW (:2031)
                    k.fb_scale_colmax(FB, self.totals[i - 1:i], self.gcount[i - 1:i], 0,
                                      self.colmax[i])
                k.fw_renorm(FW, self.colmax[i], self.wcol[i], self.w2[i])
                k.scale_matrix(FB, self.F, Kb, self.wcol[i], False, True)
                self.totals[i:i + 1].zero_()
                ncol = fc["L"] if fc["has_tb"] else self.N
                k.scale_matrix(TW, Kw, ncol, self.w2[i], True, False, self.totals[i:i + 1])
                self._redraw_hook(i, fc)
                if fc["has_tb"]:
                    # TB /= its row means, TW *= them (:2026-2030); sum(TW) is taken again
                    L = fc["L"]
                    TB = fc["TB"][:L]
                    k.row_sums(TB, L, self.N, self.lvec)
                    lv = self.lvec[:L]
                    lv.div_(float(self.N_total))
                    lv[lv == 0] = 1.0
                    k.scale_matrix(TB, L, self.N, self.lvec, True, True)
                    self.totals[i:i + 1].zero_()
                    k.scale_matrix(TW, Kw, L, self.lvec, False, False, self.totals[i:i + 1])
                if fi < nfac - 1:  # TW /= its mean (:2032-2033); the sum itself is kept for
                    # the next factor and the restart check
                    torch.div(self.totals[i:i + 1], self.gcount[i:i + 1], out=self.gvec[:1])
                    self.gvec[:Kw] = self.gvec[:1].expand(Kw).clone()
                    k.scale_matrix(TW, Kw, ncol, self.gvec, True, True)
                i += 1
        # a TW that vanished would be re-drawn at random by the reference (:2023-2025)
        self._check_totals()

    def _check_totals(self):
        tot = self.totals.clone()
        self.k.check_totals(tot, EPS, self.flags, self.iter_dev, self.first_vanish)

    def redraw_vanished_TW(self):
        pass  # (done inside the factor loop of renormalize, see _redraw_hook)

    def _redraw_hook(self, i, fc):
        """Host-synchronous renormalisation (renormalize_parameters() called by hand): a TW whose
        sum fell below eps is re-drawn at random right here, like the reference (:2023-2025)."""
        if not getattr(self, "sync_redraw", False):
            return
        if float(self.totals[i].cpu().item()) < EPS:
            Kw = fc["Kw"]
            ncol = fc["L"] if fc["has_tb"] else self.N
            Z = np.random.randn(Kw, ncol) ** 2 * (1e3 * EPS)
            fc["TW"][:Kw, :ncol] = self._upload(Z, self.tdtype)
            self.totals[i] = float(Z.sum())
            self.redrawn = getattr(self, "redrawn", 0) + 1

    # ------------------------------------------------------------------ sparsity
    def gem_iteration(self, n_iter_total, logliks, mark=None):
        super(GeneralGemEngine, self).gem_iteration(n_iter_total, logliks, mark)
        if not self.sparsity_enabled or not any(sp["sparsity"] for sp in self.spec):
            return
        # sigma: exp(log K^2 + (log 9 - log K^2) / max(iter_num - 1, 1) * i)  (:2937-2977)
        log0 = float(np.log(max(sp["fac"][0]["Kw"] for sp in self.spec) ** 2))
        slope = (float(np.log(9.0)) - log0) / max(n_iter_total - 1.0, 1.0)
        for sp in self.spec:
            fc = sp["fac"][0]
            if sp["sparsity"] and fc["Kw"] > 2:
                self.k.sparsity_reweigh(fc["TW"], fc["Kw"], self.N, int(sp["sparsity"]), log0,
                                        slope, self.iter_dev, self.sparse_work)

    # ------------------------------------------------------------------ results
    def read_model(self, spat_comps, spec_comps, gather=True):
        A = self.A.cpu().numpy()
        for j in range(self.J):
            if self.mix_type == "inst":
                spat_comps[j]["params"] = np.ascontiguousarray(A[self.ranks[j], :, 0].T)
            else:
                spat_comps[j]["params"] = np.ascontiguousarray(A[self.ranks[j]])
        f64 = self.torch.float64
        done = {}
        for s, sp in enumerate(self.spec):
            for fc in sp["fac"]:
                fac = spec_comps[s]["factor"][fc["key"]]
                mats = [("FB", self.F, fc["Kb"]), ("FW", fc["Kb"], fc["Kw"]),
                        ("TW", fc["Kw"], fc["L"] if fc["has_tb"] else self.N)]
                if fc["has_tb"]:
                    mats.append(("TB", fc["L"], self.N))
                for name, r, c in mats:
                    host = fc[name + "_host"]
                    if id(host) not in done:
                        done[id(host)] = fc[name][:r, :c].to(f64).cpu().numpy()
                    new = done[id(host)]
                    if isinstance(host, np.ndarray) and host.dtype == np.float64 and \
                            host.flags.writeable:
                        host[...] = new      # in place: shared arrays stay shared objects
                        fac[name] = host
                    else:
                        fac[name] = new
