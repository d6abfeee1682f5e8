"""Window functions and small helpers (host side).

Mirrors pyfasst/tools/utils.py:17-72 (same names and values; the reference's
known-answer tests pyfasst_tests/pyfasst/tools/test_utils.py:11-63 are re-run
against these in tests/test_api_cpu.py).
"""
import numpy as np


def db(val):
    """10 log10(val)  (ref: tools/utils.py:17-23)"""
    return 10 * np.log10(val)


def ident(energy):
    """identity (ref: tools/utils.py:25-28)"""
    return energy


def nextpow2(i):
    """Smallest power of two >= i, at least 2 (ref: tools/utils.py:30-41)."""
    n = 2
    while n < i:
        n *= 2
    return n


def sinebell(lengthWindow):
    """sin(pi t / L), t = 0..L-1 (ref: tools/utils.py:43-57)."""
    return np.sin(np.pi * np.arange(lengthWindow) / (1.0 * lengthWindow))


def hann(args):
    """numpy.hanning (ref: tools/utils.py:59-65)."""
    return np.hanning(args)


def sqrt_blackmanharris(M):
    """Root of the 4-term Blackman-Harris window (ref: tools/utils.py:67-72,
    scipy.signal.blackmanharris: symmetric, a = .35875, .48829, .14128, .01168)."""
    if M < 1:
        return np.array([])
    if M == 1:
        return np.ones(1)
    n = np.arange(M)
    fac = 2.0 * np.pi * n / (M - 1)
    w = 0.35875 - 0.48829 * np.cos(fac) + 0.14128 * np.cos(2 * fac) - 0.01168 * np.cos(3 * fac)
    return np.sqrt(np.maximum(w, 0.0))
