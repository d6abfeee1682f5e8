"""SIMM / Stereo_SIMM: (stereo) Smoothed Instantaneous Mixture Model parameter estimation,
drop-in for pyfasst/SeparateLeadStereo/SIMM/SIMM.py (`SIMM` :46-395, `Stereo_SIMM` :397-943):
same names, argument order, defaults, return tuples and error behaviour; NumPy in / NumPy out.
The multiplicative-update loops run on the GPU (pyfasst_b200/simm_engine.py, csrc/simm.cu,
csrc/gemm_tc.cu); there is no CPU fallback.

Arguments that only drive plotting / GUI feedback in the reference (displayEvolution, makeMovie,
imageCanvas, progressBar, F0Table, chirpPerF0, stepNotes, lambdaHF0, alphaHF0) are accepted and
ignored: they do not enter the arithmetic (SIMM.py:152-193 only builds tick labels from them).
"""
import numpy as np
from numpy.random import randn

from ...simm_engine import SimmEngine

__all__ = ["SIMM", "Stereo_SIMM"]


def _initial(given, shape, name, verbose_dims=True):
    """The reference's handling of an optional initial matrix (SIMM.py:198-248): random
    |N(0,1)| when missing or mis-shaped (global NumPy RNG, drawn in the reference's order)."""
    if given is None:
        return np.abs(randn(*shape))
    arr = np.array(given, dtype=np.float64)
    if arr.shape != tuple(shape):
        print("Wrong dimensions for given %s, \n" % name)
        print("random initialization used instead")
        return np.abs(randn(*shape))
    return np.array(arr, copy=True, order="C")


def _kernels(kernels):
    if kernels is not None:
        return kernels
    from ...tftransforms.stft import default_kernels
    return default_kernels()  # raises without a GPU / the built library: no CPU fallback


def SIMM(SX, WF0, WGAMMA, numberOfFilters=4, numberOfAccompanimentSpectralShapes=10,
         HGAMMA0=None, HPHI0=None, HF00=None, WM0=None, HM0=None, numberOfIterations=1000,
         updateRulePower=1.0, stepNotes=4, lambdaHF0=0.00, alphaHF0=0.99, displayEvolution=False,
         verbose=True, makeMovie=False, imageCanvas=None, progressBar=None, F0Table=None,
         chirpPerF0=1, kernels=None):
    """HGAMMA, HPHI, HF0, HM, WM, recoError = SIMM(SX, WF0, WGAMMA, ...)  (SIMM.py:46-395).

    SX ~ (WF0 HF0) * (WGAMMA HGAMMA HPHI) + WM HM, Itakura-Saito multiplicative updates in the
    order HF0, HPHI, HM, HGAMMA, WM.  Returns False when WF0 does not have F rows (:195-196)."""
    K, R = numberOfFilters, numberOfAccompanimentSpectralShapes
    SX = np.asarray(SX)
    F, N = SX.shape
    Fwf0, NF0 = np.shape(WF0)
    _, P = np.shape(WGAMMA)
    if Fwf0 != F:
        return False
    HGAMMA0 = _initial(HGAMMA0, (P, K), "HGAMMA0")
    HPHI0 = _initial(HPHI0, (K, N), "HPHI0")
    HF00 = _initial(HF00, (NF0, N), "HF00")
    HM0 = _initial(HM0, (R, N), "HM0")
    WM0 = _initial(WM0, (F, R), "WM0")
    eng = SimmEngine(_kernels(kernels), [SX], WF0, WGAMMA, HGAMMA0, HPHI0, HF00, WM0, HM0,
                     omega=updateRulePower, n_iter=numberOfIterations)
    for n in range(numberOfIterations):
        if verbose:
            print("iteration ", n, " over ", numberOfIterations)
        eng.iterate()
    r = eng.results()
    # the mono function never fills recoError (SIMM.py:274, :395)
    return r["HGAMMA"], r["HPHI"], r["HF0"], r["HM"], r["WM"], np.zeros_like(r["recoError"])


def Stereo_SIMM(SXR, SXL, WF0, WGAMMA, numberOfFilters=4, numberOfAccompanimentSpectralShapes=10,
                HGAMMA0=None, HPHI0=None, HF00=None, WM0=None, HM0=None,
                numberOfIterations=1000, updateRulePower=1.0, stepNotes=4, lambdaHF0=0.00,
                alphaHF0=0.99, displayEvolution=False, verbose=True, updateHGAMMA=True,
                computeError=False, kernels=None):
    """alphaR, alphaL, HGAMMA, HPHI, HF0, betaR, betaL, HM, WM, recoError =
    Stereo_SIMM(SXR, SXL, WF0, WGAMMA, ...)  (SIMM.py:397-943); betaR / betaL are returned as
    diagonal matrices (:943).  Raises ValueError when the two spectrograms differ in shape
    (:509-514); returns False when WF0 does not have F rows (:520-521)."""
    K, R = numberOfFilters, numberOfAccompanimentSpectralShapes
    SXR, SXL = np.asarray(SXR), np.asarray(SXL)
    F, N = SXR.shape
    if (F, N) != SXL.shape:
        print("The input STFT matrices do not have the same dimension.\n")
        print("Please check what happened...")
        raise ValueError("Dimension of STFT matrices must be the same.")
    Fwf0, NF0 = np.shape(WF0)
    _, P = np.shape(WGAMMA)
    if Fwf0 != F:
        return False
    HGAMMA0 = _initial(HGAMMA0, (P, K), "HGAMMA0")
    HPHI0 = _initial(HPHI0, (K, N), "HPHI0")
    HF00 = _initial(HF00, (NF0, N), "HF00")
    HM0 = _initial(HM0, (R, N), "HM0")
    WM0 = _initial(WM0, (F, R), "WM0")
    betaR = np.random.rand(R)  # :581
    eng = SimmEngine(_kernels(kernels), [SXR, SXL], WF0, WGAMMA, HGAMMA0, HPHI0, HF00, WM0, HM0,
                     betaR=betaR, omega=updateRulePower, update_hgamma=updateHGAMMA,
                     compute_error=computeError, n_iter=numberOfIterations)
    if computeError and verbose:
        print("Reconstruction error at beginning: ", float(eng.reco[0]))
    for n in range(numberOfIterations):
        if verbose:
            print("iteration ", n, " over ", numberOfIterations)
        eng.iterate()
    r = eng.results()
    return (r["alphaR"], r["alphaL"], r["HGAMMA"], r["HPHI"], r["HF0"], np.diag(r["betaR"]),
            np.diag(r["betaL"]), r["HM"], r["WM"], r["recoError"])
