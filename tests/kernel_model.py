"""numpy model of the *algebra* the CUDA E-step kernel uses (test-only).

The kernel does not walk Rtot^2 planes like the reference
(audioModel.py:698-731); it accumulates, per frequency, source-pair moments of
M = y y^H - Sigma^-1 (y = Sigma^-1 x) and contracts them with the mixing
vectors once per frequency.  This file states that algebra in numpy so that
tests can check it against the oracle on the CPU (any dtype), independently of
the CUDA implementation.
"""
import numpy as np

EPS = 1e-10


def spat_moments(A_sub, ranks, dtype=np.float64):
    """Per-frequency R_j = sum_{r in j} a_r a_r^H as (a, b, c, d) = (R00, R11,
    Re R01, Im R01) and the pair coefficients D_jj' of the positive determinant
    expansion.  A_sub[R, 2, F] complex; ranks = list of index arrays."""
    J = len(ranks)
    F = A_sub.shape[2]
    Rj = np.zeros([J, 4, F])
    for j, idx in enumerate(ranks):
        a = A_sub[idx]
        Rj[j, 0] = (np.abs(a[:, 0]) ** 2).sum(0)
        Rj[j, 1] = (np.abs(a[:, 1]) ** 2).sum(0)
        r01 = (a[:, 0] * np.conj(a[:, 1])).sum(0)
        Rj[j, 2], Rj[j, 3] = r01.real, r01.imag
    D = np.zeros([J, J, F])
    for j in range(J):
        for k in range(j, J):
            if j == k:
                D[j, k] = Rj[j, 0] * Rj[j, 1] - Rj[j, 2] ** 2 - Rj[j, 3] ** 2
            else:
                D[j, k] = (Rj[j, 0] * Rj[k, 1] + Rj[j, 1] * Rj[k, 0]
                           - 2 * (Rj[j, 2] * Rj[k, 2] + Rj[j, 3] * Rj[k, 3]))
    D = np.maximum(D, 0.0)
    return Rj.astype(dtype), D.astype(dtype)


def estep_moments(X, V, A_sub, ranks, noise, dtype=np.float64):
    """X[2,F,N] complex, V[J,F,N], A_sub[R,2,F], noise[F].
    Returns hat_Rxs[F,2,R], hat_Rss[F,R,R], hat_W[J,F,N], loglik."""
    J = len(ranks)
    R = A_sub.shape[0]
    F, N = V.shape[1:]
    Rj, D = spat_moments(A_sub, ranks, dtype)
    t = dtype
    x0r, x0i = X[0].real.astype(t), X[0].imag.astype(t)
    x1r, x1i = X[1].real.astype(t), X[1].imag.astype(t)
    v = V.astype(t)
    s2 = noise.astype(t)[:, None]
    col = lambda a: a[:, None]
    s00 = s2 + sum(v[j] * col(Rj[j, 0]) for j in range(J))
    s11 = s2 + sum(v[j] * col(Rj[j, 1]) for j in range(J))
    s01r = sum(v[j] * col(Rj[j, 2]) for j in range(J))
    s01i = sum(v[j] * col(Rj[j, 3]) for j in range(J))
    det = s2 * (s00 + (s11 - s2))
    for j in range(J):
        for k in range(j, J):
            det = det + v[j] * v[k] * col(D[j, k])
    det = np.maximum(det, t(EPS))
    idet = t(1) / det
    i00, i11 = s11 * idet, s00 * idet
    i01r, i01i = -s01r * idet, -s01i * idet
    y0r = i00 * x0r + i01r * x1r - i01i * x1i
    y0i = i00 * x0i + i01r * x1i + i01i * x1r
    y1r = i01r * x0r + i01i * x0i + i11 * x1r
    y1i = i01r * x0i - i01i * x0r + i11 * x1i
    quad = x0r * y0r + x0i * y0i + x1r * y1r + x1i * y1i
    ll = -(np.log(det * t(np.pi)) + quad).astype(np.float64).mean()
    m00 = y0r * y0r + y0i * y0i - i00
    m11 = y1r * y1r + y1i * y1i - i11
    m01r = y0r * y1r + y0i * y1i - i01r
    m01i = y0i * y1r - y0r * y1i - i01i
    hat_W = np.zeros([J, F, N], dtype=t)
    for j, idx in enumerate(ranks):
        q = (col(Rj[j, 0]) * m00 + col(Rj[j, 1]) * m11
             + 2 * (col(Rj[j, 2]) * m01r + col(Rj[j, 3]) * m01i))
        hat_W[j] = np.abs(v[j] + v[j] * v[j] * q * t(1.0 / len(idx)))
    # per-frequency moments (float64 accumulation over n)
    f64 = np.float64
    S = np.zeros([J, J, F, 2, 2], dtype=complex)
    for j in range(J):
        for k in range(j, J):
            p = (v[j] * v[k]).astype(f64)
            S[j, k, :, 0, 0] = (p * m00).sum(1)
            S[j, k, :, 1, 1] = (p * m11).sum(1)
            S[j, k, :, 0, 1] = (p * m01r).sum(1) + 1j * (p * m01i).sum(1)
            S[j, k, :, 1, 0] = np.conj(S[j, k, :, 0, 1])
            S[k, j] = S[j, k]
    # U = x y^H
    x = [x0r + 1j * x0i, x1r + 1j * x1i]
    y = [y0r + 1j * y0i, y1r + 1j * y1i]
    T = np.zeros([J, F, 2, 2], dtype=complex)
    for j in range(J):
        for c in range(2):
            for c2 in range(2):
                T[j, :, c, c2] = (v[j] * x[c] * np.conj(y[c2])).astype(
                    complex).sum(1)
    sv = v.astype(f64).sum(2)  # [J,F]
    src_of = np.zeros(R, dtype=int)
    for j, idx in enumerate(ranks):
        src_of[idx] = j
    Af = np.transpose(A_sub, (2, 1, 0))  # [F,2,R]
    hat_Rss = np.zeros([F, R, R], dtype=complex)
    hat_Rxs = np.zeros([F, 2, R], dtype=complex)
    for r1 in range(R):
        for r2 in range(R):
            Sm = S[src_of[r1], src_of[r2]]
            hat_Rss[:, r1, r2] = np.einsum("fi,fij,fj->f", np.conj(Af[:, :, r1]),
                                           Sm, Af[:, :, r2]) / N
        hat_Rss[:, r1, r1] += sv[src_of[r1]] / N
        hat_Rxs[:, :, r1] = np.einsum("fij,fj->fi", T[src_of[r1]],
                                      Af[:, :, r1]) / N
    hat_Rss = 0.5 * (hat_Rss + np.conj(np.transpose(hat_Rss, (0, 2, 1))))
    return hat_Rxs, hat_Rss, hat_W, ll
