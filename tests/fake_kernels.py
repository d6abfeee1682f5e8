"""NumPy stand-ins for the kernels of the C ABI (TEST INFRASTRUCTURE ONLY).

`FakeKernels` has the method surface of pyfasst_b200._lib.CudaKernels but works on
torch CPU tensors with NumPy.  Each method is the *specification* of the CUDA kernel of
the same name: the `-m gpu` tests compare every CUDA kernel with its stand-in, and the
CPU tests run the GEM engine's orchestration (pyfasst_b200/engine.py) on the stand-ins
and compare the result with the oracle / the reference's golden vectors -- so the host
logic is verified without a GPU.  The product never imports this file: GemEngine is
always constructed with CudaKernels outside tests/.

Plane arithmetic is done in the dtype of the plane tensors (float32 or float64) and
reductions in float64, like the kernels.
"""
import numpy as np
import torch

EPS = 1e-10


def _np(t):
    return t.numpy() if t is not None else None


class FakeKernels(object):
    name = "fake"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def dtype_code(self, t):
        return {torch.float32: 0, torch.float64: 1}[t.dtype]

    def launch_count(self):
        return self.launches

    # ---- K1 / K6 -------------------------------------------------------------------
    def stft(self, pcm, window, hop, nfft, X, N, psd_sum, pcm_div=1.0, sample0=0, L_total=None,
             frame0=0):
        self.launches += 1
        x, w = _np(pcm), _np(window)
        if x.dtype != np.float64:
            x = x.T  # interleaved [L, nch] -> planar
        x = x.astype(np.float64) / pcm_div
        nch, L = x.shape
        L_total = L if L_total is None else L_total
        wlen = w.size
        Xo = _np(X)
        Xo[:] = 0
        ntot = int(np.ceil(L_total / float(hop)) + 2)
        for c in range(nch):
            buf = np.zeros((ntot - 1) * hop + wlen + nfft)
            buf[wlen // 2 + sample0: wlen // 2 + sample0 + L] = x[c]
            idx = hop * (frame0 + np.arange(N))[:, None] + np.arange(wlen)[None, :]
            S = np.fft.rfft(w[None, :] * buf[idx], nfft, axis=1).T
            Xo[2 * c, :, :N] = S.real
            Xo[2 * c + 1, :, :N] = S.imag
        if psd_sum is not None:
            _np(psd_sum)[:] = (Xo[:, :, :N].astype(np.float64) ** 2).sum(axis=(0, 2))

    def overlap_norm(self, prod, hop, N):
        self.launches += 1
        prod = np.asarray(prod, dtype=np.float64)
        norm = np.zeros(hop * (N - 1) + prod.size)
        for n in range(N):
            norm[n * hop:n * hop + prod.size] += prod
        return torch.tensor(norm)

    def pcm_peak(self, pcm, peak):
        self.launches += 1
        x = _np(pcm)
        if np.issubdtype(x.dtype, np.signedinteger):
            x = x[x != np.iinfo(x.dtype).min]
        _np(peak)[0] = np.abs(x.astype(np.float64)).max() if x.size else 0.0

    def istft(self, Y, N, synth, norm, hop, nfft, out, pcm, maxdata, drop=None, pcm_round=False):
        self.launches += 1
        Yn, ws, nrm = _np(Y), _np(synth), _np(norm)
        nsig = Yn.shape[0] // 2
        wlen = ws.size
        total = (N - 1) * hop + wlen
        o = _np(out)
        Lout = o.shape[1]
        for s in range(nsig):
            S = Yn[2 * s, :, :N].astype(np.float64) + 1j * Yn[2 * s + 1, :, :N].astype(np.float64)
            frames = np.fft.irfft(S.T, nfft, axis=1)[:, :wlen] * ws[None, :]
            data = np.zeros(total)
            for n in range(N):
                data[n * hop:n * hop + wlen] += frames[n]
            data = (data / nrm)[wlen // 2 if drop is None else drop:]
            m = min(Lout, data.size)
            o[s, :m] = data[:m]
        if pcm is not None:
            _np(pcm)[:] = np.int16(np.round(o.T * maxdata) if pcm_round else o.T * maxdata)

    def _spat(self, A, src_of_sub, J):
        """R_j = sum_r a_r a_r^H per frequency: [J, F] arrays r00, r11, r01."""
        A = _np(A)
        F = A.shape[2]
        r00, r11 = np.zeros([J, F]), np.zeros([J, F])
        r01 = np.zeros([J, F], dtype=complex)
        for r, j in enumerate(src_of_sub):
            r00[j] += np.abs(A[r, 0]) ** 2
            r11[j] += np.abs(A[r, 1]) ** 2
            r01[j] += A[r, 0] * np.conj(A[r, 1])
        return r00, r11, r01

    def _sigma_inv_y(self, X, V, A, src_of_sub, noise, N):
        """Per-bin algebra in float64 whatever the storage type (like the kernels, which
        keep Sigma^-1 and y in double: see estep_stereo_kernel)."""
        J = V.shape[0]
        t = np.float64
        r00, r11, r01 = self._spat(A, src_of_sub, J)
        Xn, Vn = _np(X)[:, :, :N].astype(t), _np(V)[:, :, :N].astype(t)
        s2 = _np(noise).astype(t)[:, None]
        col = lambda a: a.astype(t)[:, None]
        s00 = s2 + sum(Vn[j] * col(r00[j]) for j in range(J))
        s11 = s2 + sum(Vn[j] * col(r11[j]) for j in range(J))
        s01r = sum(Vn[j] * col(r01[j].real) for j in range(J))
        s01i = sum(Vn[j] * col(r01[j].imag) for j in range(J))
        det = s2 * (s00 + (s11 - s2))
        for j in range(J):
            for k in range(j, J):
                if j == k:
                    d = r00[j] * r11[j] - np.abs(r01[j]) ** 2
                else:
                    d = r00[j] * r11[k] + r11[j] * r00[k] - 2 * np.real(r01[j] * np.conj(r01[k]))
                det = det + Vn[j] * Vn[k] * col(np.maximum(d, 0.0))
        det = np.maximum(det, t(EPS))
        idet = t(1) / det
        i00, i11, i01r, i01i = s11 * idet, s00 * idet, -s01r * idet, -s01i * idet
        x0r, x0i, x1r, x1i = Xn
        y0r = i00 * x0r + i01r * x1r - i01i * x1i
        y0i = i00 * x0i + i01r * x1i + i01i * x1r
        y1r = i01r * x0r + i01i * x0i + i11 * x1r
        y1i = i01r * x0i - i01i * x0r + i11 * x1i
        return (r00, r11, r01), det, (i00, i11, i01r, i01i), (y0r, y0i, y1r, y1i), col

    def wiener_stereo(self, X, V, A, src_of_sub, noise, group_of_src, ngroups, N, Y, workspace):
        self.launches += 2
        J = V.shape[0]
        (r00, r11, r01), det, inv, (y0r, y0i, y1r, y1i), col = self._sigma_inv_y(
            X, V, A, src_of_sub, noise, N)
        Vn, Yn = _np(V)[:, :, :N].astype(np.float64), _np(Y)
        Yn[:] = 0
        for g in range(ngroups):
            g00 = sum(Vn[j] * col(r00[j]) for j in range(J) if group_of_src[j] == g)
            g11 = sum(Vn[j] * col(r11[j]) for j in range(J) if group_of_src[j] == g)
            g01r = sum(Vn[j] * col(r01[j].real) for j in range(J) if group_of_src[j] == g)
            g01i = sum(Vn[j] * col(r01[j].imag) for j in range(J) if group_of_src[j] == g)
            Yn[4 * g + 0, :, :N] = g00 * y0r + g01r * y1r - g01i * y1i
            Yn[4 * g + 1, :, :N] = g00 * y0i + g01r * y1i + g01i * y1r
            Yn[4 * g + 2, :, :N] = g01r * y0r + g01i * y0i + g11 * y1r
            Yn[4 * g + 3, :, :N] = g01r * y0i - g01i * y0r + g11 * y1i

    # ---- K2 / K6 for I = 2..4 channels ---------------------------------------------------
    def _multi_core(self, X, V, A, src_of_sub, noise, N):
        """Sigma^-1 (clamped, Q5 on the generic determinant) and y = Sigma^-1 x in float64."""
        Xn = _np(X)[:, :, :N].astype(np.float64)
        I = Xn.shape[0] // 2
        x = np.transpose(Xn[0::2] + 1j * Xn[1::2], (1, 2, 0))       # [F, N, I]
        Vn = _np(V)[:, :, :N].astype(np.float64)                    # [J, F, N]
        J = Vn.shape[0]
        An = _np(A)                                                  # [R, I, F]
        Rj = np.zeros([J, An.shape[2], I, I], dtype=complex)
        for r, j in enumerate(src_of_sub):
            a = An[r].T                                              # [F, I]
            Rj[j] += a[:, :, None] * np.conj(a[:, None, :])
        Sig = np.einsum("jfn,jfab->fnab", Vn, Rj) + \
            _np(noise)[:, None, None, None] * np.eye(I)[None, None]
        det = np.real(np.linalg.det(Sig))
        detc = np.maximum(det, EPS)
        Sinv = np.linalg.inv(Sig) * (det / detc)[..., None, None]
        y = np.einsum("fnab,fnb->fna", Sinv, x)
        return x, Vn, Rj, Sinv, y, detc

    def estep_multi_workspace_bytes(self, I, J, F, N):
        return 64

    def estep_multi(self, X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace,
                    N_norm=0):
        self.launches += 3
        Nn = N_norm if N_norm > 0 else N
        x, Vn, Rj, Sinv, y, detc = self._multi_core(X, V, A, src_of_sub, noise, N)
        J, F = Vn.shape[0], Vn.shape[1]
        R = A.shape[0]
        quad = np.real(np.sum(np.conj(x) * y, axis=2))
        _np(ll_f)[:] = (np.log(detc) + np.log(np.pi) + quad).sum(1)
        M = y[..., :, None] * np.conj(y[..., None, :]) - Sinv        # [F, N, I, I]
        count = np.bincount(src_of_sub, minlength=J)
        hw = _np(hatW)
        hw[:] = 0
        for j in range(J):
            q = np.real(np.einsum("fnab,fba->fn", M, Rj[j]))
            hw[j, :, :N] = np.abs(Vn[j] + Vn[j] * Vn[j] * q / count[j])
        # moments: factors and sums in float64 whatever the plane type (the general-I kernel
        # keeps its shared-memory records in float64)
        U = x[..., :, None] * np.conj(y[..., None, :])
        S = {}
        for j in range(J):
            for k in range(j, J):
                S[j, k] = S[k, j] = np.einsum("fn,fnab->fab", Vn[j] * Vn[k], M)
        T = [np.einsum("fn,fnab->fab", Vn[j], U) for j in range(J)]
        sv = Vn.sum(2)
        An = _np(A)
        hRss, hRxs = _np(Rss), _np(Rxs)
        for r1 in range(R):
            j1 = src_of_sub[r1]
            a1 = An[r1].T                                            # [F, I]
            hRxs[:, :, r1] = np.einsum("fab,fb->fa", T[j1], a1) / Nn
            for r2 in range(r1, R):
                j2 = src_of_sub[r2]
                a2 = An[r2].T
                h = np.einsum("fa,fab,fb->f", np.conj(a1), S[j1, j2], a2) / Nn
                if r1 == r2:
                    h = np.real(h) + sv[j1] / Nn
                hRss[:, r1, r2] = h
                hRss[:, r2, r1] = np.conj(h)

    def wiener_multi(self, X, V, A, src_of_sub, noise, group_of_src, ngroups, N, Y, workspace):
        self.launches += 2
        x, Vn, Rj, Sinv, y, detc = self._multi_core(X, V, A, src_of_sub, noise, N)
        I = x.shape[2]
        Yn = _np(Y)
        Yn[:] = 0
        for g in range(ngroups):
            Sg = sum(np.einsum("fn,fab->fnab", Vn[j], Rj[j]) for j in range(Vn.shape[0])
                     if group_of_src[j] == g)
            out = np.einsum("fnab,fnb->fna", Sg, y)                  # [F, N, I]
            for i in range(I):
                Yn[2 * I * g + 2 * i, :, :N] = out[..., i].real
                Yn[2 * I * g + 2 * i + 1, :, :N] = out[..., i].imag

    # ---- K2 ---------------------------------------------------------------------------
    def estep_workspace_bytes(self, J, F, N, dtype_code):
        return 64

    def estep_stereo_inst(self, X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace,
                          N_norm=0):
        """Real mixing vectors: the real parts of the statistics, zero imaginary parts."""
        assert float(np.abs(_np(A).imag).max()) == 0.0, "estep_stereo_inst needs real mixing vectors"
        self.estep_stereo(X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace, N_norm)
        _np(Rss).imag[...] = 0.0
        _np(Rxs).imag[...] = 0.0

    def estep_stereo(self, X, V, A, src_of_sub, noise, N, hatW, Rss, Rxs, ll_f, workspace,
                     N_norm=0):
        self.launches += 3
        Nn = N_norm if N_norm > 0 else N
        J, F, ld = V.shape
        R = A.shape[0]
        t = np.float64
        (r00, r11, r01), det, (i00, i11, i01r, i01i), (y0r, y0i, y1r, y1i), col = \
            self._sigma_inv_y(X, V, A, src_of_sub, noise, N)
        Xn, Vn = _np(X)[:, :, :N].astype(t), _np(V)[:, :, :N].astype(t)
        x0r, x0i, x1r, x1i = Xn
        quad = x0r * y0r + x0i * y0i + x1r * y1r + x1i * y1i
        _np(ll_f)[:] = (np.log(det) + np.log(np.pi) + quad).sum(1)
        m00 = y0r * y0r + y0i * y0i - i00
        m11 = y1r * y1r + y1i * y1i - i11
        m01r = y0r * y1r + y0i * y1i - i01r
        m01i = y0i * y1r - y0r * y1i - i01i
        count = np.bincount(src_of_sub, minlength=J)
        hw = _np(hatW)
        hw[:] = 0
        for j in range(J):
            q = (col(r00[j]) * m00 + col(r11[j]) * m11
                 + t(2) * (col(r01[j].real) * m01r + col(r01[j].imag) * m01i))
            hw[j, :, :N] = np.abs(Vn[j] + Vn[j] * Vn[j] * (q * t(1.0 / count[j])))
        f64 = np.float64
        S = np.zeros([J, J, F, 2, 2], dtype=complex)
        for j in range(J):
            for k in range(j, J):
                p = Vn[j] * Vn[k]
                S[j, k, :, 0, 0] = (p * m00).astype(f64).sum(1)
                S[j, k, :, 1, 1] = (p * m11).astype(f64).sum(1)
                S[j, k, :, 0, 1] = (p * m01r).astype(f64).sum(1) + 1j * (p * m01i).astype(f64).sum(1)
                S[j, k, :, 1, 0] = np.conj(S[j, k, :, 0, 1])
                S[k, j] = S[j, k]
        x = [x0r + 1j * x0i, x1r + 1j * x1i]
        y = [y0r + 1j * y0i, y1r + 1j * y1i]
        T = np.zeros([J, F, 2, 2], dtype=complex)
        for j in range(J):
            for c in range(2):
                for c2 in range(2):
                    T[j, :, c, c2] = (Vn[j] * x[c] * np.conj(y[c2])).astype(complex).sum(1)
        sv = Vn.astype(f64).sum(2)
        Af = np.transpose(_np(A), (2, 1, 0))  # [F, 2, R]
        hRss, hRxs = _np(Rss), _np(Rxs)
        for r1 in range(R):
            j1 = src_of_sub[r1]
            for r2 in range(R):
                hRss[:, r1, r2] = np.einsum("fi,fij,fj->f", np.conj(Af[:, :, r1]),
                                            S[j1, src_of_sub[r2]], Af[:, :, r2]) / Nn
            hRss[:, r1, r1] = np.real(hRss[:, r1, r1]) + sv[j1] / Nn
            hRxs[:, :, r1] = np.einsum("fij,fj->fi", T[j1], Af[:, :, r1]) / Nn
        hRss[:] = 0.5 * (hRss + np.conj(np.transpose(hRss, (0, 2, 1))))

    # ---- K3 ---------------------------------------------------------------------------
    def mix_inst_stats(self, Rss, Rxs, A, upd, oth, stats):
        self.launches += 1
        hRss, hRxs, An = _np(Rss), _np(Rxs), _np(A)
        F = An.shape[2]
        rxs = hRxs[:, :, upd].copy()
        if len(oth):
            for f in range(F):
                rxs[f] -= np.dot(An[oth, :, f].T, hRss[f][np.ix_(oth, upd)])
        rxs = np.real(rxs.sum(0))  # [I, Ku]
        rss = np.real(hRss[:, np.vstack(upd), upd].sum(0))  # [Ku, Ku]
        _np(stats)[:] = np.concatenate([rxs.ravel(), rss.ravel()])

    def mix_inst_solve(self, stats, F_total, upd, A, flags):
        self.launches += 1
        st = _np(stats)
        Ku = len(upd)
        I = _np(A).shape[1]
        rxs = st[:I * Ku].reshape(I, Ku) / F_total
        rss = st[I * Ku:I * Ku + Ku * Ku].reshape(Ku, Ku) / F_total
        try:
            sol = np.linalg.solve(rss.T, rxs.T)
        except np.linalg.LinAlgError:
            _np(flags)[0] |= 1
            return
        _np(A)[upd] = sol[:, :, None]

    def mix_conv_solve(self, Rss, Rxs, A, flags):
        self.launches += 1
        hRss, hRxs, An = _np(Rss), _np(Rxs), _np(A)
        for f in range(An.shape[2]):
            try:
                An[:, :, f] = np.linalg.solve(hRss[f].T, hRxs[f].T)
            except np.linalg.LinAlgError:
                _np(flags)[0] |= 1

    # ---- K4 ---------------------------------------------------------------------------
    def spec_power(self, W, H, V, N, accumulate):
        self.launches += 1
        out = np.dot(_np(W), _np(H)[:, :N])
        Vn = _np(V)
        if accumulate:
            Vn[:, :N] += out
        else:
            Vn[:, :N] = out
        Vn[:, N:] = 0

    def small_matmul(self, A, B, C):
        self.launches += 1
        _np(C)[:] = np.dot(_np(A).astype(np.float64), _np(B).astype(np.float64))

    def fb_plan(self, F, K, N, dtype_code):
        return N, 1

    def fb_contract(self, hatW, P, O, G, N, num_partial, den_partial, chunk, nsplit):
        self.launches += 1
        t = _np(hatW).dtype.type
        hw, p, o = _np(hatW)[:, :N], _np(P)[:, :N], _np(O)[:, :N]
        g = _np(G)[:, :N]
        p, o = np.maximum(p, t(EPS)), np.maximum(o, t(EPS))
        e2 = o / p
        e1 = hw / p / p * o
        _np(num_partial)[:] = 0
        _np(den_partial)[:] = 0
        _np(num_partial)[0] = np.dot(e1.astype(np.float64), g.T.astype(np.float64))
        _np(den_partial)[0] = np.dot(e2.astype(np.float64), g.T.astype(np.float64))

    def tw_plan(self, F, K, N, dtype_code):
        return F, 1

    def tw_contract(self, hatW, O, W, H, N, num_partial, den_partial, fchunk, fsplit,
                    scratch=None):
        self.launches += 1
        t = _np(hatW).dtype.type
        hw, o = _np(hatW)[:, :N], np.maximum(_np(O)[:, :N], t(EPS))
        Wn, Hn = _np(W), _np(H)[:, :N]
        p = np.maximum(np.dot(Wn, Hn), t(EPS))
        e2 = o / p
        e1 = o * (hw / p / p)
        _np(num_partial)[:] = 0
        _np(den_partial)[:] = 0
        _np(num_partial)[0, :, :N] = np.dot(Wn.T.astype(np.float64), e1.astype(np.float64))
        _np(den_partial)[0, :, :N] = np.dot(Wn.T.astype(np.float64), e2.astype(np.float64))

    def tw_pack_chunks(self, num_partial, den_partial, out, world):
        self.launches += 1
        o = _np(out)
        nsplit, K, ld = num_partial.shape
        c = ld // world
        for q, part in enumerate((num_partial, den_partial)):
            tot = _np(part).sum(0)                                   # [K, ld]
            o[:, q, :K] = tot.reshape(K, world, c).transpose(1, 0, 2)

    def sum_splits(self, parts, out):
        self.launches += 1
        p = _np(parts)
        _np(out).reshape(-1)[:] = p.reshape(p.shape[0], -1).sum(0)

    def mult_update(self, theta, num, den, rows, cols, omega):
        self.launches += 1
        th = _np(theta)
        ratio = _np(num)[:rows, :cols] / np.maximum(_np(den)[:rows, :cols], EPS)
        th[:rows, :cols] = th[:rows, :cols].astype(np.float64) * ratio ** omega

    def mult_update_splits(self, theta, num_partial, den_partial, rows, cols, omega):
        self.launches += 1
        th = _np(theta)
        num = _np(num_partial).sum(0)[:rows, :cols]
        den = _np(den_partial).sum(0)[:rows, :cols]
        th[:rows, :cols] = th[:rows, :cols].astype(np.float64) * (num / np.maximum(den, EPS)) ** omega

    # ---- K5 ---------------------------------------------------------------------------
    def spat_energy(self, A, src_of_sub, J, sums):
        self.launches += 1
        An = _np(A)
        s = np.zeros(J)
        for r, j in enumerate(src_of_sub):
            s[j] += (np.abs(An[r]) ** 2).sum()
        _np(sums)[:] = s

    def spat_scale(self, A, src_of_sub, sums, counts):
        self.launches += 1
        An = _np(A)
        e = _np(sums) / _np(counts)
        for r, j in enumerate(src_of_sub):
            An[r] /= np.sqrt(e[j])

    def fb_scale_colmax(self, FB, sums, counts, j, colmax):
        self.launches += 1
        fb = _np(FB)
        g = _np(sums)[j] / _np(counts)[j]
        fb[:] = fb.astype(np.float64) * g
        _np(colmax)[:fb.shape[1]] = fb.astype(np.float64).max(axis=0)

    def fw_renorm(self, FW, colmax, w, w2):
        self.launches += 1
        fw = _np(FW)
        Kb, Kw = fw.shape
        wv = _np(colmax)[:Kb].copy()
        wv[wv == 0] = 1.0
        _np(w)[:Kb] = wv
        fw[:] = fw.astype(np.float64) * wv[:, None]
        w2v = fw.astype(np.float64).mean(axis=0)
        w2v[w2v == 0] = 1.0
        _np(w2)[:Kw] = w2v
        fw[:] = fw.astype(np.float64) / w2v

    def scale_matrix(self, M, rows, cols, s, by_row, divide, total=None):
        self.launches += 1
        m = _np(M)
        sv = _np(s)
        sc = sv[:rows, None] if by_row else sv[None, :cols]
        v = m[:rows, :cols].astype(np.float64)
        v = v / sc if divide else v * sc
        m[:rows, :cols] = v
        if total is not None:
            _np(total)[0] += m[:rows, :cols].astype(np.float64).sum()

    # ---- general factor structures (csrc/gemfac.cu) ---------------------------------------
    def gem_ratio_planes(self, hatW, P, O, out, N, Ptot=None, Pminus=None, lam=0.0):
        self.launches += 1
        F, ld = hatW.shape
        p = np.maximum(_np(P)[:, :N].astype(np.float64), EPS)
        o = np.maximum(_np(O)[:, :N].astype(np.float64), EPS)
        on = _np(out)
        on[:] = 0
        if Ptot is None:
            on[:, :N] = _np(hatW)[:, :N] / p ** 2 * o
            on[:, ld:ld + N] = o / p
        else:  # correlation penalty (audioModel.py:1544-1567)
            pt = _np(Ptot)[:, :N].astype(np.float64)
            c = lam * _np(Pminus)[:, :N] / np.maximum(pt ** 2, EPS)
            on[:, ld:ld + N] = o * (1. / p + c)
            on[:, :N] = (_np(hatW)[:, :N] / p ** 2 + c * 2 * (p / pt)) * o

    def corr_planes(self, V, own, Ptot, Pminus, N, clamp=True):
        self.launches += 1
        v = _np(V)[:, :, :N].astype(np.float64)
        pt = np.maximum(v.sum(axis=0), EPS)
        pm = pt - np.maximum(v[own], EPS)
        if clamp:
            pm = np.maximum(pm, EPS)
        for t, a in ((Ptot, pt), (Pminus, pm)):
            tn = _np(t)
            tn[:] = 0
            tn[:, :N] = a

    def row_sums(self, M, rows, cols, out):
        self.launches += 1
        _np(out)[:rows] = _np(M)[:rows, :cols].astype(np.float64).sum(axis=1)

    def apply_filter(self, X, W, Y, N):
        self.launches += 1
        Xn, Wn, Yn = _np(X), _np(W), _np(Y)
        nc = Xn.shape[0] // 2
        Yn[:] = 0
        Xc = [Xn[2 * c, :, :N].astype(np.float64) + 1j * Xn[2 * c + 1, :, :N] for c in range(nc)]
        for c1 in range(nc):
            acc = 0
            for c2 in range(nc):
                w = Wn[c1, c2][:, :N] if Wn.ndim == 4 else Wn[c1, c2][:, None]
                acc = acc + w * Xc[c2]
            Yn[2 * c1, :, :N], Yn[2 * c1 + 1, :, :N] = acc.real, acc.imag

    def mul_planes(self, a, b, out, N, accumulate=False):
        self.launches += 1
        v = _np(a)[:, :N].astype(np.float64)
        if b is not None:
            v = v * _np(b)[:, :N]
        on = _np(out)
        if accumulate:
            v = v + on[:, :N]
        on[:, :N] = v
        on[:, N:] = 0

    def mult_update_same(self, theta, num, den, rows, cols, omega):
        self.launches += 1
        th = _np(theta)
        ratio = _np(num)[:rows, :cols].astype(np.float64) / np.maximum(
            _np(den)[:rows, :cols].astype(np.float64), EPS)
        th[:rows, :cols] = th[:rows, :cols].astype(np.float64) * ratio ** omega

    def sparsity_reweigh(self, TW, K, N, length, log_sigma0, slope, iter_dev, work):
        """Specification of pf_sparsity_reweigh (audioModel.py:2981-3014)."""
        self.launches += 3
        tw = _np(TW)
        v = tw[:K, :N].astype(np.float64)
        sigma = np.exp(log_sigma0 + slope * (int(_np(iter_dev)[0]) - 1))
        w = np.arange(K - 1, 0, -1) ** 2
        mu = np.dot(np.arange(K - 1) * w, v[:-1]) / np.dot(w, np.maximum(v[:-1], EPS))
        muf = np.zeros_like(mu)
        for n in range(N):
            win = mu[max(n - length, 0):min(n + length, N - 1)]
            muf[n] = np.median(win) if win.size else np.nan
            if np.isnan(muf[n]):
                muf[n] = mu[n]
        mask = np.exp(-0.5 * ((np.arange(K)[:, None] - muf) ** 2) / sigma)
        mask[-1] = mask.max(axis=0)
        pos = mask[-1] > 0
        mask[:, pos] /= mask[-1][pos]
        tw[:K, :N] = v * mask

    def check_totals(self, totals, eps, flags, iter_dev=None, first_iter=None):
        self.launches += 1
        tt = _np(totals)
        if (tt < eps).any():
            _np(flags)[0] |= 2
            if first_iter is not None:
                it = int(_np(iter_dev)[0]) if iter_dev is not None else 0
                _np(first_iter)[0] = min(int(_np(first_iter)[0]), it)
        tt[:] = 0

    # ---- glue ---------------------------------------------------------------------------
    def noise_anneal(self, sqrt0, sqrt1, iter_dev, n_iter, noise):
        self.launches += 1
        i = float(_np(iter_dev)[0])
        _np(noise)[:] = ((_np(sqrt0) * (n_iter - i) + _np(sqrt1) * i) / n_iter) ** 2

    def ll_reduce(self, ll_f, ll_sum):
        self.launches += 1
        _np(ll_sum)[0] = _np(ll_f).sum()

    def ll_store(self, ll_sum, bins, logliks, iter_dev, advance):
        self.launches += 1
        i = int(_np(iter_dev)[0])
        _np(logliks)[i] = -_np(ll_sum)[0] / bins
        if advance:
            _np(iter_dev)[0] = i + 1
