"""NumPy stand-ins for the SIMM kernels of the C ABI (TEST INFRASTRUCTURE ONLY).

Same role as tests/fake_kernels.py: every method is the specification of the CUDA kernel of
the same name in pyfasst_b200/csrc/simm.cu / gemm_tc.cu, working on torch CPU tensors (views
allowed).  The CPU tests run the SIMM host orchestration (pyfasst_b200/simm_engine.py) on these
and compare it with the oracle and the reference's golden vectors; the `-m gpu` tests compare
each CUDA kernel with its stand-in.  The product never imports this file.
"""
import numpy as np
import torch

EPS = np.float32(1e-20)


def _np(t):
    return None if t is None else t.numpy()


class FakeSimmKernels(object):
    name = "fake"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def launch_count(self):
        return self.launches

    # ---- dense products ---------------------------------------------------------------
    def gemm_view(self, A, B, C, M, N, K, transA=False, transB=False, workspace=None):
        self.launches += 1
        a, b = _np(A), _np(B)
        a = a[:K, :M].T if transA else a[:M, :K]
        b = b[:N, :K].T if transB else b[:K, :N]
        c = _np(C)
        c[:M, :N] = (a.astype(np.float64) @ b.astype(np.float64)).astype(c.dtype)
        if workspace is not None:
            c[:M, N:] = 0

    def gemm_splitk_workspace_bytes(self, M, N, K):
        return 16

    def simm_reduce_workspace_bytes(self):
        return 16

    def spec_power(self, W, H, V, N, accumulate):
        self.launches += 1
        w, h, v = _np(W), _np(H), _np(V)
        K = w.shape[1]
        p = (w.astype(np.float64) @ h[:K, :N].astype(np.float64)).astype(v.dtype)
        if accumulate:
            v[:, :N] += p
        else:
            v[:, :N] = p

    def small_matmul(self, A, B, C):
        self.launches += 1
        a, b, c = _np(A), _np(B), _np(C)
        c[:a.shape[0], :b.shape[1]] = (a.astype(np.float64) @ b.astype(np.float64)).astype(c.dtype)

    # ---- planes ------------------------------------------------------------------------
    def _hat(self, SM, SF0, SPHI, a2, nch, N, ldn):
        """hat_c = max(a2_c SF0 SPHI + SM_c, eps) as a [F, nch * ldn] array (padding = 1)."""
        m, s0, sp, a = _np(SM), _np(SF0), _np(SPHI), _np(a2)
        h = np.ones_like(m)
        lead = s0[:, :N] * sp[:, :N]
        for c in range(nch):
            h[:, c * ldn:c * ldn + N] = np.maximum(a[c] * lead + m[:, c * ldn:c * ldn + N], EPS)
        return h

    def simm_lead_terms(self, SM, SF0, SPHI, SX, a2, other_is_sf0, out, nch, F, N, ldn):
        self.launches += 1
        h = self._hat(SM, SF0, SPHI, a2, nch, N, ldn)
        o, x, a, w = _np(SF0 if other_is_sf0 else SPHI), _np(SX), _np(a2), _np(out)
        num = np.zeros((F, N), np.float32)
        den = np.zeros((F, N), np.float32)
        for c in range(nch):
            ih = 1 / np.maximum(h[:, c * ldn:c * ldn + N], EPS)
            cc = a[c] * o[:, :N] * ih
            den += cc
            num += cc * x[:, c * ldn:c * ldn + N] * ih
        w[:, :2 * ldn] = 0
        w[:, :N] = num
        w[:, ldn:ldn + N] = den

    def simm_acc_terms(self, SM, SF0, SPHI, SX, a2, out, nch, sq_clamp, F, N, ldn):
        self.launches += 1
        h, x, w = self._hat(SM, SF0, SPHI, a2, nch, N, ldn), _np(SX), _np(out)
        w[:, :2 * nch * ldn] = 0
        for c in range(nch):
            hc, xc = h[:, c * ldn:c * ldn + N], x[:, c * ldn:c * ldn + N]
            iv = 1 / np.maximum(hc, EPS)
            t = xc / np.maximum(hc * hc, EPS) if sq_clamp else iv * xc * iv
            w[:, c * ldn:c * ldn + N] = t
            w[:, (nch + c) * ldn:(nch + c) * ldn + N] = iv

    def simm_hat(self, SM, SF0, SPHI, a2, hat, nch, F, N, ldn):
        self.launches += 1
        _np(hat)[:, :nch * ldn] = self._hat(SM, SF0, SPHI, a2, nch, N, ldn)

    def simm_is_divergence(self, SX, SM, SF0, SPHI, a2, nch, F, N, ldn, workspace, out):
        self.launches += 1
        x, h = _np(SX), self._hat(SM, SF0, SPHI, a2, nch, N, ldn)
        tot = 0.0
        for c in range(nch):
            r = x[:, c * ldn:c * ldn + N] / h[:, c * ldn:c * ldn + N]
            tot += np.sum((r - 1 - np.log(r)).astype(np.float64))
        _np(out)[0] = tot

    def simm_alpha_update(self, SX, SM, SF0, SPHI, F, N, ldn, omega, workspace, alpha, a2):
        self.launches += 1
        x, al, a2n = _np(SX), _np(alpha), _np(a2)
        h = self._hat(SM, SF0, SPHI, a2, 2, N, ldn)
        lead = _np(SF0)[:, :N] * _np(SPHI)[:, :N]
        new = []
        for c in range(2):
            ih = 1 / np.maximum(h[:, c * ldn:c * ldn + N], EPS)
            d = lead * ih
            num = np.sum((d * x[:, c * ldn:c * ldn + N] * ih).astype(np.float64))
            den = np.sum(d.astype(np.float64))
            new.append(max(al[c] * (num / den) ** (omega * 0.1), 1e-20))
        r = new[0] / max(new[0] + new[1], 0.001)
        al[0], al[1] = r, 1 - r
        a2n[0], a2n[1] = r * r, (1 - r) * (1 - r)

    # ---- small matrices ------------------------------------------------------------------
    def simm_update_rows(self, theta, C, nch, ldn, w, omega, floor_value, rows, N):
        self.launches += 1
        th, c = _np(theta), _np(C)
        num = np.zeros((rows, N), np.float32)
        den = np.zeros((rows, N), np.float32)
        for ch in range(nch):
            wc = np.float32(1) if w is None else _np(w)[ch, :rows, None]
            num += wc * c[:rows, ch * ldn:ch * ldn + N]
            den += wc * c[:rows, (nch + ch) * ldn:(nch + ch) * ldn + N]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = num / np.maximum(den, EPS)
            g = ratio if omega == 1.0 else ratio ** np.float32(omega)
            t = th[:rows, :N] * g
        if floor_value > 0:
            t = np.maximum(t, np.float32(floor_value))
        th[:rows, :N] = t

    def simm_hphi_normalise(self, HPHI, K, rowscale, N, s_out):
        self.launches += 1
        h = _np(HPHI)
        if rowscale is not None:
            h[:K, :N] *= _np(rowscale)[:K, None]
        s = h[:K, :N].sum(axis=0)
        pos = s > 0
        h[:K, :N][:, pos] /= s[pos]
        _np(s_out)[:N] = s

    def simm_scale_columns(self, P, rows, N, s):
        self.launches += 1
        p = _np(P)
        p[:rows, :N] *= _np(s)[:N]

    def simm_scale_rows(self, P, rows, N, s):
        self.launches += 1
        p = _np(P)
        p[:rows, :N] *= _np(s)[:rows, None]

    def simm_hgamma_update(self, HGAMMA, WGAMMA, tn, td, F, P, K, omega, s_out):
        self.launches += 1
        hg, wg = _np(HGAMMA), _np(WGAMMA)[:F, :P].astype(np.float64)
        num = wg.T @ _np(tn)[:F, :K].astype(np.float64)
        den = wg.T @ _np(td)[:F, :K].astype(np.float64)
        ratio = (num / np.maximum(den, 1e-20)).astype(np.float32)
        hg[:P, :K] *= ratio if omega == 1.0 else ratio ** np.float32(omega)
        s = hg[:P, :K].sum(axis=0)
        pos = s > 0
        hg[:P, :K][:, pos] /= s[pos]
        _np(s_out)[:K] = s

    def simm_wm_update(self, WM, R, D, nch, b2, clamp_den, omega, F, s_out):
        self.launches += 1
        wm, d = _np(WM), _np(D)
        num = np.zeros((F, R), np.float32)
        den = np.zeros((F, R), np.float32)
        for c in range(nch):
            wc = np.float32(1) if b2 is None else _np(b2)[c, None, :R]
            num += wc * d[c, :, :R]
            den += wc * d[nch + c, :, :R]
        if clamp_den:
            den = np.maximum(den, EPS)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = num / den
            wm[:, :R] *= ratio if omega == 1.0 else ratio ** np.float32(omega)
        s = wm[:, :R].astype(np.float64).sum(axis=0).astype(np.float32)
        pos = s > 0
        wm[:, :R][:, pos] /= s[pos]
        _np(s_out)[:R] = s

    def simm_beta_update(self, WM, R, D, F, omega, beta, b2):
        self.launches += 1
        wm, d, be, b2n = _np(WM)[:, :R].astype(np.float64), _np(D), _np(beta), _np(b2)
        dg = [np.sum(wm * d[q, :, :R].astype(np.float64), axis=0) for q in range(4)]
        bR = be[0, :R] * (dg[0] / dg[2]) ** (omega * 0.1)
        bL = be[1, :R] * (dg[1] / dg[3]) ** (omega * 0.1)
        bR = bR / np.maximum(bR + bL, 1e-20)
        be[0, :R], be[1, :R] = bR, 1 - bR
        b2n[0, :R], b2n[1, :R] = bR ** 2, (1 - bR) ** 2

    def simm_power(self, X, SX, nch, F, N, ldn):
        self.launches += 1
        x, sx = _np(X), _np(SX)
        sx[:, :nch * ldn] = 0
        for c in range(nch):
            sx[:, c * ldn:c * ldn + N] = x[2 * c, :, :N] ** 2 + x[2 * c + 1, :, :N] ** 2

    def simm_masks(self, SM, SF0, SPHI, a2, X, Y, eps_hat, nch, F, N, ldn):
        self.launches += 1
        m, a, x, y = _np(SM), _np(a2), _np(X), _np(Y)
        lead = _np(SF0)[:, :N] * _np(SPHI)[:, :N]
        for c in range(nch):
            lv, sm = a[c] * lead, m[:, c * ldn:c * ldn + N]
            ih = 1 / np.maximum(lv + sm, np.float32(eps_hat))
            for part in (0, 1):
                y[2 * c + part, :, :N] = lv * ih * x[2 * c + part, :, :N]
                y[2 * (nch + c) + part, :, :N] = sm * ih * x[2 * c + part, :, :N]

    # ---- glottal-source F0 dictionary -----------------------------------------------------------
    def wf0_combs(self, f1, f2, npart, fs, Ot, Lsig, t_begin, window, nfft, rows):
        """Specification of wf0_comb_kernel (csrc/wf0.cu): power spectrum of one windowed frame
        of the KLGLOTT88 waveform per column, float64."""
        self.launches += 1
        f1, f2 = np.asarray(f1, dtype=np.float64), np.asarray(f2, dtype=np.float64)
        window = np.asarray(window, dtype=np.float64)
        wlen = window.size
        out = np.zeros((f1.size, rows))
        t = t_begin + np.arange(wlen)
        inside = (t >= 0) & (t < Lsig)
        tau = t / float(fs)
        for c in range(f1.size):
            P = int(npart[c])
            frame = np.zeros(wlen)
            if P > 0:
                h = np.arange(1, P + 1)
                z = 1j * 2.0 * np.pi * h * Ot
                E = np.exp(-z)
                amp = (f1[c] + f2[c]) / 2.0 * 27 / 4 * (E + 2 * (1 + 2 * E) / z
                                                          - 6 * (1 - E) / z ** 2) / z
                cyc = f1[c] * tau + (f2[c] - f1[c]) * tau ** 2 / (2.0 * Lsig / fs)
                x = np.real(np.sum(amp[:, None] * np.exp(2j * np.pi * np.outer(h, cyc)), axis=0))
                frame = np.where(inside, x, 0.0) * window
            spec = np.fft.fft(frame, nfft)
            out[c] = np.abs(spec[np.arange(rows) % nfft]) ** 2
        return torch.tensor(out)

    # ---- Viterbi --------------------------------------------------------------------------------
    def viterbi(self, log_density, log_prior, log_trans):
        self.launches += 2
        dens, prior, trans = _np(log_density), _np(log_prior), _np(log_trans)
        S, N = dens.shape
        cum = prior + dens[:, 0]
        ante = np.zeros((N, S), dtype=np.int64)
        for n in range(1, N):
            cand = cum[:, None] + trans          # [from, to]
            ante[n] = np.argmax(cand, axis=0)    # first maximum, like the strict `>` scan
            cum = cand[ante[n], np.arange(S)] + dens[:, n]
        path = np.zeros(N, dtype=np.int64)
        path[-1] = np.argmax(cum)
        for n in range(N - 2, -1, -1):
            path[n] = ante[n + 1, path[n + 1]]
        return torch.from_numpy(path)

    # ---- IS-NMF initialisers ------------------------------------------------------------------
    def nmf_is_terms(self, hat, SX, out, eps, F, N, ldn):
        self.launches += 1
        h, x, w = _np(hat)[:, :N], _np(SX)[:, :N], _np(out)
        e = np.float32(eps)
        w[:, :2 * ldn] = 0
        w[:, :N] = x / np.maximum(h * h, e)
        w[:, ldn:ldn + N] = 1 / np.maximum(h, e)

    def nmf_update_rows(self, H, C, ldn, eps, rows, N):
        self.launches += 1
        h, c = _np(H), _np(C)
        h[:rows, :N] *= c[:rows, :N] / np.maximum(c[:rows, ldn:ldn + N], np.float32(eps))

    def nmf_w_update(self, W, K, D, eps, F, s_out):
        self.launches += 1
        w, d = _np(W), _np(D)
        w[:, :K] *= d[0, :, :K] / np.maximum(d[1, :, :K], np.float32(eps))
        s = w[:, :K].astype(np.float64).sum(axis=0).astype(np.float32)
        s[s == 0] = 1
        w[:, :K] /= s
        _np(s_out)[:K] = s

    def mono_power(self, X, out, F, N, ldn):
        self.launches += 1
        x, o = _np(X), _np(out)
        o[:, :ldn] = 0
        o[:, :N] = (x[:, :, :N].astype(np.float32) ** 2).sum(axis=0) / np.float32(x.shape[0] // 2)

    def simm_wm_scaled(self, WM, R, b2, nch, F, WMs):
        self.launches += 1
        wm, out = _np(WM), _np(WMs)
        out[:] = 0
        for c in range(nch):
            out[c, :, :R] = wm[:, :R] * (1 if b2 is None else _np(b2)[c, None, :R])
