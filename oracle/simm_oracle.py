"""CPU oracle for the SIMM / Stereo_SIMM multiplicative-update loops (lead / accompaniment
source-filter model of SeparateLeadStereo).

TEST INFRASTRUCTURE ONLY -- imported by tests/ and bench.py's CPU arm, never by the product.

float64 NumPy restatement of pyfasst/SeparateLeadStereo/SIMM/SIMM.py: `SIMM` (:46-395) and
`Stereo_SIMM` (:397-943).  Parity status: PINNED by tests/golden/simm.npz, produced by running
the reference itself (oracle/make_golden.py, run_simm) -- see tests/test_simm_cpu.py.

Model: SX ~ (WF0 HF0) * (WGAMMA HGAMMA HPHI) + WM HM   (mono)
       SXc ~ alpha_c^2 (WF0 HF0) * (WPHI HPHI) + (WM beta_c^2) HM,  c in {R, L}   (stereo)
All updates are Itakura-Saito multiplicative rules  theta *= (num / den)^omega; the order
HF0, HPHI, HM, HGAMMA, WM (, alpha, beta) and the renormalisations between them follow the
reference exactly; quirks kept on purpose are marked QUIRK.
"""
import numpy as np

EPS = 10 ** (-20)  # ref: SIMM.py:150, :497  (NB: 1e-20, not audioModel's 1e-10)


def is_distortion(X, Y):
    """Itakura-Saito divergence (ref: SIMM.py:35-44)."""
    ratio = X / Y
    return np.sum(-np.log(ratio) + ratio - 1)


def simm(SX, WF0, WGAMMA, HGAMMA0, HPHI0, HF00, WM0, HM0, numberOfIterations=1000,
         updateRulePower=1.0):
    """Mono model.  ref: SIMM.py:46-395 (update loop :303-393).  All initial matrices must be
    given (the reference draws the missing ones from the unseeded global RNG).
    Returns (HGAMMA, HPHI, HF0, HM, WM)."""
    om = updateRulePower
    HGAMMA, HPHI, HF0 = HGAMMA0.copy(), HPHI0.copy(), HF00.copy()
    WM, HM = WM0.copy(), HM0.copy()
    WPHI = WGAMMA @ HGAMMA
    SF0, SPHI, SM = WF0 @ HF0, WPHI @ HPHI, WM @ HM
    hat = SF0 * SPHI + SM  # QUIRK: not clamped before the first update (:268)
    WF0T = WF0.T
    for _ in range(numberOfIterations):
        # HF0 (:303-315)
        den = SPHI / np.maximum(hat, EPS)
        num = den * SX / np.maximum(hat, EPS)
        HF0 *= ((WF0T @ num) / np.maximum(WF0T @ den, EPS)) ** om
        SF0 = WF0 @ HF0
        hat = np.maximum(SF0 * SPHI + SM, EPS)
        # HPHI (:319-331), columns normalised to sum one, the gain goes to HF0
        den = SF0 / np.maximum(hat, EPS)
        num = den * SX / np.maximum(hat, EPS)
        HPHI *= ((WPHI.T @ num) / np.maximum(WPHI.T @ den, EPS)) ** om
        s = HPHI.sum(axis=0)
        HPHI[:, s > 0] /= s[s > 0]
        HF0 *= s
        SF0, SPHI = WF0 @ HF0, WPHI @ HPHI
        hat = np.maximum(SF0 * SPHI + SM, EPS)
        # HM (:335-347)
        den = 1 / np.maximum(hat, EPS)
        num = den * SX / np.maximum(hat, EPS)
        HM *= ((WM.T @ num) / np.maximum(WM.T @ den, EPS)) ** om
        HM = np.maximum(HM, EPS)
        SM = WM @ HM
        hat = np.maximum(SF0 * SPHI + SM, EPS)
        # HGAMMA (:351-372)
        den = SF0 / np.maximum(hat, EPS)
        num = den * SX / np.maximum(hat, EPS)
        HGAMMA *= ((WGAMMA.T @ (num @ HPHI.T)) /
                   np.maximum(WGAMMA.T @ (den @ HPHI.T), EPS)) ** om
        s = HGAMMA.sum(axis=0)
        HGAMMA[:, s > 0] /= s[s > 0]
        HPHI *= s[:, None]
        s = HPHI.sum(axis=0)
        HPHI[:, s > 0] /= s[s > 0]
        HF0 *= s
        WPHI = WGAMMA @ HGAMMA
        SF0, SPHI = WF0 @ HF0, WPHI @ HPHI
        hat = np.maximum(SF0 * SPHI + SM, EPS)
        # WM (:376-393)
        den = 1 / np.maximum(hat, EPS)
        num = den * SX / np.maximum(hat, EPS)
        WM *= ((num @ HM.T) / np.maximum(den @ HM.T, EPS)) ** om
        s = WM.sum(axis=0)
        WM[:, s > 0] /= s[s > 0]
        HM *= s  # QUIRK: broadcasts sumWM[R] against HM[R, N] columns -> only valid for R == 1 (:388)
        SM = WM @ HM
        hat = np.maximum(SF0 * SPHI + SM, EPS)
    return HGAMMA, HPHI, HF0, HM, WM


def stereo_simm(SXR, SXL, WF0, WGAMMA, HGAMMA0, HPHI0, HF00, WM0, HM0, betaR0,
                numberOfIterations=1000, updateRulePower=1.0, updateHGAMMA=True,
                computeError=False):
    """Stereo model with lead panning alpha and accompaniment pannings beta[R].
    ref: SIMM.py:397-943 (update loop :613-941).  `betaR0` replaces the reference's
    np.random.rand(R) (:581).  Returns (alphaR, alphaL, HGAMMA, HPHI, HF0, diag(betaR),
    diag(betaL), HM, WM, recoError)."""
    om = updateRulePower
    F, N = SXR.shape
    if SXL.shape != (F, N):
        raise ValueError("Dimension of STFT matrices must be the same.")
    NF0 = WF0.shape[1]
    HGAMMA, HPHI, HF0 = HGAMMA0.copy(), HPHI0.copy(), HF00.copy()
    WM, HM = WM0.copy(), HM0.copy()
    aR, aL = 0.5, 0.5
    bR = np.array(betaR0, dtype=np.float64).copy()
    bL = 1 - bR
    WPHI = WGAMMA @ HGAMMA
    SF0, SPHI = WF0 @ HF0, WPHI @ HPHI
    WF0T = WF0.T

    def hats(clamp=True):
        lead = SF0 * SPHI
        hR = (WM * bR ** 2) @ HM + aR ** 2 * lead
        hL = aL ** 2 * lead + (WM * bL ** 2) @ HM
        if clamp:
            hR, hL = np.maximum(hR, EPS), np.maximum(hL, EPS)
        return hR, hL

    hR, hL = hats(clamp=False)  # QUIRK: unclamped before the first update (:587-592)
    recoError = np.zeros([numberOfIterations * 5 * 2 + NF0 * 2 + 1])
    if computeError:
        recoError[0] = is_distortion(SXR, hR) + is_distortion(SXL, hL)
    counter = 1

    def lead_terms(other):
        """num / den planes of the lead-side updates for the factor multiplying `other`."""
        cR = aR ** 2 * other / np.maximum(hR, EPS)
        cL = aL ** 2 * other / np.maximum(hL, EPS)
        num = cR * SXR / np.maximum(hR, EPS) + cL * SXL / np.maximum(hL, EPS)
        return num, cL + cR

    for _ in range(numberOfIterations):
        # HF0 (:622-664)
        num, den = lead_terms(SPHI)
        HF0 *= ((WF0T @ num) / np.maximum(WF0T @ den, EPS)) ** om
        SF0 = WF0 @ HF0
        hR, hL = hats()
        if computeError:
            recoError[counter] = is_distortion(SXR, hR) + is_distortion(SXL, hL)
        counter += 1
        # HPHI (:685-730)
        num, den = lead_terms(SF0)
        HPHI *= ((WPHI.T @ num) / np.maximum(WPHI.T @ den, EPS)) ** om
        s = HPHI.sum(axis=0)
        HPHI[:, s > 0] = HPHI[:, s > 0] / s[s > 0]
        HF0 *= s
        SF0, SPHI = WF0 @ HF0, WPHI @ HPHI
        hR, hL = hats()
        if computeError:
            recoError[counter] = is_distortion(SXR, hR) + is_distortion(SXL, hL)
        counter += 1
        # HM (:741-763)
        WR, WL = WM * bR ** 2, WM * bL ** 2
        HM *= ((WR.T @ (SXR / np.maximum(hR ** 2, EPS)) + WL.T @ (SXL / np.maximum(hL ** 2, EPS))) /
               np.maximum(WR.T @ (1 / np.maximum(hR, EPS)) + WL.T @ (1 / np.maximum(hL, EPS)),
                          EPS)) ** om
        hR, hL = hats()
        counter += 1
        # HGAMMA (:776-819)
        if updateHGAMMA:
            num, den = lead_terms(SF0)
            HGAMMA *= ((WGAMMA.T @ (num @ HPHI.T)) /
                       np.maximum(WGAMMA.T @ (den @ HPHI.T), EPS)) ** om
            s = HGAMMA.sum(axis=0)
            HGAMMA[:, s > 0] /= s[s > 0]
            HPHI *= s[:, None]
            s = HPHI.sum(axis=0)
            HPHI[:, s > 0] /= s[s > 0]
            HF0 *= s
            WPHI = WGAMMA @ HGAMMA
            SF0, SPHI = WF0 @ HF0, WPHI @ HPHI
            hR, hL = hats()
            counter += 1
        # WM (:829-866)  QUIRK (Q12): the denominator is not clamped
        TR, TL = SXR / np.maximum(hR ** 2, EPS), SXL / np.maximum(hL ** 2, EPS)
        IR, IL = 1 / np.maximum(hR, EPS), 1 / np.maximum(hL, EPS)
        HR, HL = HM.T * bR ** 2, HM.T * bL ** 2
        WM = WM * ((TR @ HR + TL @ HL) / (IR @ HR + IL @ HL)) ** om
        s = WM.sum(axis=0)
        WM[:, s > 0] /= s[s > 0]
        HM *= s[:, None]
        hR, hL = hats()
        counter += 1
        # alpha (:869-896): exponent omega / 10, then alphaR / (alphaR + alphaL)
        lead = SF0 * SPHI
        d = lead / np.maximum(hR, EPS)
        aR_new = np.maximum(aR * (np.sum(d * SXR / np.maximum(hR, EPS)) / np.sum(d)) ** (om * .1), EPS)
        d = lead / np.maximum(hL, EPS)
        aL_new = np.maximum(aL * (np.sum(d * SXL / np.maximum(hL, EPS)) / np.sum(d)) ** (om * .1), EPS)
        aR = aR_new / np.maximum(aR_new + aL_new, .001)
        aL = 1 - aR
        hR, hL = hats()
        counter += 1
        # beta (:909-941): only the diagonal of the R x R products is used
        TR, TL = SXR / np.maximum(hR ** 2, EPS), SXL / np.maximum(hL ** 2, EPS)
        IR, IL = 1 / np.maximum(hR, EPS), 1 / np.maximum(hL, EPS)
        dg = lambda T: np.einsum("fr,fn,rn->r", WM, T, HM)
        bR = bR * (dg(TR) / dg(IR)) ** (om * .1)
        bL = bL * (dg(TL) / dg(IL)) ** (om * .1)
        bR = bR / np.maximum(bR + bL, EPS)
        bL = 1 - bR
        hR, hL = hats()
        counter += 1
    return aR, aL, HGAMMA, HPHI, HF0, np.diag(bR), np.diag(bL), HM, WM, recoError
