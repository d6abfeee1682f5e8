"""Parity at BASELINE.json's FULL size (configs[1]: 10-min stereo 44.1 kHz mixture, STFT
2048 / hop 512, F = 1025, N = 51 682, 4 sources x K = 32, rank 2) through properties that do
not need the oracle to walk 53 M bins:

  * the E-step and the Wiener filter are independent per frequency (audioModel.py:613-764,
    :1327-1467): a handful of frequency rows of the full-size launch, ALL frames, against the
    NumPy specification of the kernel (tests/fake_kernels.py, itself pinned to the oracle by
    tests/test_kernel_model.py / test_oracle_golden.py);
  * the sufficient statistics are sums over frames: the full launch equals the sum of two
    launches over the two halves of the frames, and the per-bin outputs are bit-identical;
  * STFT -> inverse STFT reconstructs the 26.46 M-sample PCM (stft.py:71-131);
  * without annealing and with the plain multiplicative updates the GEM log-likelihood does
    not decrease (SURVEY.md 8c iii), and the separated images add up to the mixture minus the
    noise posterior (conservation of the Wiener filter).

Everything goes through the C ABI (ctypes) or the public API, as the product does."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from tests.fake_kernels import FakeKernels

pytestmark = pytest.mark.gpu

FS, WLEN, HOP = 44100, 2048, 512
F_FULL = WLEN // 2 + 1
L_FULL = 600 * FS
N_FULL = int(np.ceil(L_FULL / float(HOP)) + 2)  # 51 682 (stft.py:47)
J, RANK = 4, 2
ROWS = [0, 1, 300, 777, 1024]


@pytest.fixture(scope="module")
def ck():
    from pyfasst_b200._lib import CudaKernels
    return CudaKernels()


@pytest.fixture(scope="module")
def fk():
    return FakeKernels()


def rel(a, b):
    a, b = np.asarray(a, dtype=np.complex128), np.asarray(b, dtype=np.complex128)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def device_problem(F, N, seed=7):
    """Model-consistent E-step inputs of the full size, drawn on the device (float32 planes):
    x = sum_r a_r s_r + noise with s_r ~ CN(0, v_j)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    ld = (N + 31) // 32 * 32
    R = J * RANK
    src = [j for j in range(J) for _ in range(RANK)]
    randn = lambda *s: torch.randn(*s, generator=g, device="cuda", dtype=torch.float32)
    V = torch.zeros((J, F, ld), device="cuda", dtype=torch.float32)
    # sources that are active in turns, 40 dB of dynamics: the conditioning real data has
    V[:, :, :N] = (randn(J, F, N).abs() + 0.05) * 10.0 ** (2.0 * torch.rand(
        (J, 1, N), generator=g, device="cuda") - 1.0)
    A = (torch.randn((R, 2, F), generator=g, device="cuda", dtype=torch.float64)
         + 1j * torch.randn((R, 2, F), generator=g, device="cuda", dtype=torch.float64))
    noise = torch.rand(F, generator=g, device="cuda", dtype=torch.float64) * 0.01 + 1e-3
    X = torch.zeros((4, F, ld), device="cuda", dtype=torch.float32)
    sq = noise.sqrt().to(torch.float32)[:, None] * (0.5 ** 0.5)
    for c in range(2):
        X[2 * c, :, :N] = sq * randn(F, N)
        X[2 * c + 1, :, :N] = sq * randn(F, N)
    for r in range(R):
        amp = (V[src[r], :, :N] * 0.5).sqrt()
        sr, si = amp * randn(F, N), amp * randn(F, N)
        for c in range(2):
            ar = A[r, c].real.to(torch.float32)[:, None]
            ai = A[r, c].imag.to(torch.float32)[:, None]
            X[2 * c, :, :N] += ar * sr - ai * si
            X[2 * c + 1, :, :N] += ar * si + ai * sr
        del sr, si, amp
    return ld, R, src, X, V, A, noise


def run_estep(k, X, V, A, src, noise, N, N_norm=0):
    Jn, F, ld = V.shape
    R = A.shape[0]
    dev = X.device
    hatW = torch.zeros((Jn, F, ld), dtype=V.dtype, device=dev)
    Rss = torch.zeros((F, R, R), dtype=torch.complex128, device=dev)
    Rxs = torch.zeros((F, 2, R), dtype=torch.complex128, device=dev)
    ll = torch.zeros(F, dtype=torch.float64, device=dev)
    ws = torch.zeros((k.estep_workspace_bytes(Jn, F, N, k.dtype_code(V)) + 7) // 8,
                     dtype=torch.float64, device=dev)
    k.estep_stereo(X, V, A, src, noise, N, hatW, Rss, Rxs, ll, ws, N_norm)
    return hatW, Rss, Rxs, ll


@pytest.fixture(scope="module")
def full_problem():
    return device_problem(F_FULL, N_FULL)


def test_estep_full_size_rows_against_specification(ck, fk, full_problem):
    ld, R, src, X, V, A, noise = full_problem
    N = N_FULL
    hatW, Rss, Rxs, ll = run_estep(ck, X, V, A, src, noise, N)
    torch.cuda.synchronize()
    rows = torch.tensor(ROWS, device="cuda")
    sub = lambda t, ax: t.index_select(ax, rows).contiguous().cpu()
    hw0, rss0, rxs0, ll0 = run_estep(fk, sub(X, 1), sub(V, 1), sub(A, 2), src, sub(noise, 0), N)
    hw1 = sub(hatW, 1).numpy()
    assert rel(hw1[:, :, :N], hw0.numpy()[:, :, :N]) < 1e-6
    assert (hw1[:, :, N:] == 0).all()
    # float32 moment sums over 51 682 frames (see test_kernels_gpu.test_estep_stereo)
    assert rel(sub(Rss, 0).numpy(), rss0.numpy()) < 1e-4
    assert rel(sub(Rxs, 0).numpy(), rxs0.numpy()) < 1e-4
    assert_allclose(sub(ll, 0).numpy(), ll0.numpy(), rtol=1e-6, atol=1e-6 * N)
    rss = Rss.cpu().numpy()
    assert_allclose(rss, np.conj(np.transpose(rss, (0, 2, 1))), atol=1e-14 * np.abs(rss).max())
    assert np.isfinite(hatW.sum().item()) and np.isfinite(ll.sum().item())


def test_estep_full_size_is_additive_over_frames(ck, full_problem):
    """Statistics of the whole mixture = sum of the statistics of its two halves (both
    normalised by the total frame count); per-bin outputs do not depend on the split."""
    ld, R, src, X, V, A, noise = full_problem
    N = N_FULL
    hatW, Rss, Rxs, ll = run_estep(ck, X, V, A, src, noise, N)
    N1 = 25856  # a multiple of 32: both halves start on an aligned frame
    parts = []
    for lo, hi in ((0, N1), (N1, N)):
        n = hi - lo
        ldp = (n + 31) // 32 * 32
        Xp = torch.zeros((4, F_FULL, ldp), device="cuda", dtype=torch.float32)
        Vp = torch.zeros((J, F_FULL, ldp), device="cuda", dtype=torch.float32)
        Xp[:, :, :n] = X[:, :, lo:hi]
        Vp[:, :, :n] = V[:, :, lo:hi]
        parts.append((lo, hi) + run_estep(ck, Xp, Vp, A, src, noise, n, N_norm=N))
        del Xp, Vp
    for lo, hi, hw, _, _, _ in parts:
        assert torch.equal(hw[:, :, :hi - lo], hatW[:, :, lo:hi]), "per-bin algebra must not move"
    rss = (parts[0][3] + parts[1][3]).cpu().numpy()
    rxs = (parts[0][4] + parts[1][4]).cpu().numpy()
    llh = (parts[0][5] + parts[1][5]).cpu().numpy()
    assert rel(rss, Rss.cpu().numpy()) < 2e-6
    assert rel(rxs, Rxs.cpu().numpy()) < 2e-6
    assert_allclose(llh, ll.cpu().numpy(), rtol=1e-6, atol=1e-6 * N)


def test_wiener_full_size_rows_and_conservation(ck, fk, full_problem):
    ld, R, src, X, V, A, noise = full_problem
    N = N_FULL
    groups = list(range(J))
    Y = torch.zeros((4 * J, F_FULL, ld), device="cuda", dtype=torch.float32)
    ws = torch.zeros(1 << 18, dtype=torch.float64, device="cuda")  # per-frequency coefficients
    ck.wiener_stereo(X, V, A, src, noise, groups, J, N, Y, ws)
    rows = torch.tensor(ROWS, device="cuda")
    sub = lambda t, ax: t.index_select(ax, rows).contiguous().cpu()
    Y0 = torch.zeros((4 * J, len(ROWS), ld), dtype=torch.float32)
    fk.wiener_stereo(sub(X, 1), sub(V, 1), sub(A, 2), src, sub(noise, 0), groups, J, N, Y0,
                     torch.zeros(1 << 18, dtype=torch.float64))
    Y1 = sub(Y, 1).numpy()
    assert rel(Y1[:, :, :N], Y0.numpy()[:, :, :N]) < 1e-6
    assert (Y1[:, :, N:] == 0).all()
    # sum_j Sigma_j Sigma_x^-1 x = x - s2 Sigma_x^-1 x: the images add up to the mixture up to
    # the noise posterior, whose power is below the noise floor times the bin count
    tot = Y.view(J, 4, F_FULL, ld).sum(0)
    resid = (X - tot)[:, :, :N].double().pow(2).sum(dim=(0, 2))  # per frequency
    assert (resid <= 2.0 * noise * N * 1.05).all()
    assert (resid > 0).all()


def test_stft_istft_full_size_round_trip(ck):
    g = torch.Generator(device="cuda").manual_seed(3)
    L, nch = L_FULL, 2
    pcm16 = torch.randint(-20000, 20000, (L, nch), generator=g, device="cuda",
                          dtype=torch.int32).to(torch.int16)
    window = torch.tensor(np.hanning(WLEN), device="cuda")
    N = N_FULL
    ld = (N + 31) // 32 * 32
    X = torch.zeros((2 * nch, F_FULL, ld), dtype=torch.float32, device="cuda")
    psd = torch.zeros(F_FULL, dtype=torch.float64, device="cuda")
    maxdata = 1.1 * 20000.0
    ck.stft(pcm16, window, HOP, WLEN, X, N, psd, pcm_div=maxdata)
    # overlap-added window product, as istft divides by it (stft.py:118-127)
    total = (N - 1) * HOP + WLEN
    w2 = window ** 2
    norm = torch.zeros(total, dtype=torch.float64, device="cuda")
    idx = (torch.arange(N, device="cuda")[:, None] * HOP
           + torch.arange(WLEN, device="cuda")[None, :]).reshape(-1)
    norm.index_add_(0, idx, w2.repeat(N))
    norm[norm == 0] = 1.0
    out = torch.zeros((nch, L), dtype=torch.float64, device="cuda")
    back = torch.zeros((L, nch), dtype=torch.int16, device="cuda")
    ck.istft(X, N, window, norm, HOP, WLEN, out, back, maxdata, pcm_round=True)
    ref = pcm16.to(torch.float64).t() / maxdata
    assert (out - ref).abs().max().item() < 1e-4
    assert (back.to(torch.int32) - pcm16.to(torch.int32)).abs().max().item() <= 1
    # Parseval on the mean spectrum: the window's energy times the signal power
    assert np.isfinite(psd.sum().item()) and psd.min().item() > 0


def test_gem_full_size_loglik_and_separation():
    """configs[1] through the public API: no annealing, plain multiplicative updates."""
    import pyfasst_b200.audioModel as am
    import pyfasst_b200.audioObject as ao
    from bench import synth_mix
    pcm = synth_mix(600.0)
    audio = ao.AudioObject("synthetic_mix.wav")
    audio._samplerate = FS
    audio._set_raw(pcm)
    np.random.seed(0)
    iters = 8
    m = am.MultiChanNMFInst_FASST(audio=audio, nbComps=J, nbNMFComps=32, spatial_rank=RANK,
                                  wlen=WLEN, hopsize=HOP, iter_num=iters, sim_ann_opt='no_ann',
                                  ann_PSD_lim=[None, None], compute_dtype="float32")
    assert (m.nbFreqsSigRepr, m.nbFramesSigRepr) == (F_FULL, N_FULL)
    ll = np.asarray(m.estim_param_a_post_model())
    assert ll.shape == (iters,) and np.isfinite(ll).all()
    # GEM with fixed noise: the likelihood cannot decrease (float32 planes: 1e-6 slack)
    assert (np.diff(ll) >= -1e-6 * np.abs(ll[:-1])).all(), ll
    assert ll[-1] > ll[0]
    for s in range(J):
        fac = m.spec_comps[s]['factor'][0]
        assert fac['FB'].shape == (F_FULL, 32) and fac['TW'].shape == (32, N_FULL)
        assert np.isfinite(fac['FB']).all() and np.isfinite(fac['TW']).all()
        assert (fac['FB'] >= 0).all() and (fac['TW'] >= 0).all()
    sep = m.separate_comps_pcm()
    assert sep.shape == (J, L_FULL, 2) and sep.dtype == np.int16
    # the separated images add up to the mixture (minus the small noise posterior)
    err = sep.astype(np.int32).sum(0) - pcm.astype(np.int32)
    assert np.sqrt(np.mean(err.astype(np.float64) ** 2)) < 0.02 * np.sqrt(
        np.mean(pcm.astype(np.float64) ** 2))
