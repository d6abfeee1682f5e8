"""multiChanSourceF0Filter on the CPU: two-factor source/filter components
with a shared glottal dictionary + one residual NMF component, against golden vectors made by
running the reference's own class (tests/golden/fasst_sourcefilter.npz, oracle/make_golden.py:
run_sourcefilter):
  * the oracle's general update_spectral_components / renormalize_parameters,
  * the host side (pyfasst_b200.audioModel.multiChanSourceF0Filter + GeneralGemEngine) on the
    NumPy specification of the kernels in float64."""
import os
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import fasst_oracle as fo
import pyfasst_b200.audioModel as am
from tests.fake_kernels import FakeKernels
from tests.fake_simm_kernels import FakeSimmKernels

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WAV = os.path.join(GOLDEN, "mix_inst.wav")
KW = dict(nbComps=3, nbNMFResComps=2, nbFilterComps=6, nbFilterWeigs=[3, ], minF0=100, maxF0=400,
          stepnoteF0=1, chirpPerF0=1, spatial_rank=1, sparsity=None, wlen=256, hopsize=64,
          verbose=0, ann_PSD_lim=[None, None])


class AllFakeKernels(FakeSimmKernels, FakeKernels):
    """FASST stand-ins + the GEMM / dictionary stand-ins of the SIMM set."""

    def __init__(self):
        FakeSimmKernels.__init__(self)
        FakeKernels.__init__(self)


def load():
    return np.load(os.path.join(GOLDEN, "fasst_sourcefilter.npz"))


def structure_from(g, prefix):
    """spat_comps / spec_comps of the golden snapshot, the dictionary shared like the
    reference's (one array object for both sources)."""
    shared = np.array(g["%s_FB0_0" % prefix])
    assert np.array_equal(shared, g["%s_FB1_0" % prefix])
    spat, spec = {}, {}
    for j in range(3):
        spat[j] = {'time_dep': 'indep', 'mix_type': 'inst', 'frdm_prior': 'free',
                   'params': np.array(g["%s_A%d" % (prefix, j)])}
        facs = {}
        for fi in range(2 if j < 2 else 1):
            facs[fi] = {'FB': shared if (j < 2 and fi == 0) else np.array(g["%s_FB%d_%d" % (prefix, j, fi)]),
                        'FW': np.array(g["%s_FW%d_%d" % (prefix, j, fi)]),
                        'TW': np.array(g["%s_TW%d_%d" % (prefix, j, fi)]), 'TB': [],
                        'FB_frdm_prior': 'free' if j == 2 else 'fixed',
                        'FW_frdm_prior': 'free' if (j < 2 and fi == 1) else 'fixed',
                        'TW_frdm_prior': 'free', 'TB_frdm_prior': [], 'TW_constr': 'NMF'}
        spec[j] = {'spat_comp_ind': j, 'factor': facs}
    return spat, spec


def check_snapshot(model, g, prefix, tol):
    for j in range(3):
        a, b = np.asarray(model.spat_comps[j]['params']), g["%s_A%d" % (prefix, j)]
        assert np.abs(a - b).max() / np.abs(b).max() < tol, ("A", j)
        for fi, fac in model.spec_comps[j]['factor'].items():
            for m in ("FB", "FW", "TW"):
                b = g["%s_%s%d_%d" % (prefix, m, j, fi)]
                assert np.abs(fac[m] - b).max() / np.abs(b).max() < tol, (m, j, fi)


def test_oracle_general_updates_against_reference():
    g = load()
    for iters, key, snap in ((1, "ll_it1", "it1"), (5, "logliks", "final")):
        np.random.seed(0)
        m = fo.OracleFASST(WAV, nbComps=3, nbNMFComps=2, spatial_rank=1, wlen=256, hopsize=64,
                           iter_num=iters)
        m.spat_comps, m.spec_comps = structure_from(g, "init")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ll = m.estim_param_a_post_model()
        assert_allclose(ll, g[key], rtol=1e-9)
        check_snapshot(m, g, snap, 1e-9)
        assert m.spec_comps[0]['factor'][0]['FB'] is m.spec_comps[1]['factor'][0]['FB']


def build(kernels, iters, dtype="float64"):
    m = am.multiChanSourceF0Filter(audio=WAV, iter_num=iters, kernels=kernels,
                                   compute_dtype=dtype, **KW)
    m._initialize_structures(seed=5)
    return m


def check_model(kernels, dtype, tol_init, tol_it1, tol_ll, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # the dictionary cache is written to the working directory
    g = load()
    m = build(kernels, 1, dtype)
    assert_allclose(m.F0Table, g["F0Table"], rtol=1e-12)
    assert m.spec_comps[0]['factor'][0]['FB'] is m.spec_comps[1]['factor'][0]['FB']
    check_snapshot(m, g, "init", tol_init)
    ll = m.estim_param_a_post_model()
    assert_allclose(ll, g["ll_it1"], rtol=tol_ll)
    check_snapshot(m, g, "it1", tol_it1)
    assert m.spec_comps[0]['factor'][0]['FB'] is m.spec_comps[1]['factor'][0]['FB']
    m = build(kernels, 5, dtype)
    ll = m.estim_param_a_post_model()
    assert_allclose(ll, g["logliks"], rtol=tol_ll)
    return m


def test_model_on_kernel_spec_against_reference(tmp_path, monkeypatch):
    m = check_model(AllFakeKernels(), "float64", 1e-9, 1e-8, 1e-9, tmp_path, monkeypatch)
    check_snapshot(m, load(), "final", 1e-7)


def check_sparse_model(kernels, dtype, tol_ll, tol_par, tmp_path, monkeypatch):
    """The estimation loop with the sparsity re-weighting of the source activations
    (sparsity=[2]: a median filter of length 2 for every component), 4 iterations."""
    monkeypatch.chdir(tmp_path)
    g = load()
    kw = dict(KW, sparsity=[2, ])
    m = am.multiChanSourceF0Filter(audio=WAV, iter_num=4, kernels=kernels, compute_dtype=dtype,
                                   **kw)
    m._initialize_structures(seed=5)
    assert [m.spec_comps[j]['sparsity'] for j in range(3)] == [2, 2, 2]
    ll = m.estim_param_a_post_model()
    assert_allclose(ll, g["sparse_logliks"], rtol=tol_ll)
    check_snapshot(m, g, "sparse", tol_par)
    # one GEM_iteration() does not re-weigh (only the estimation loop does, :2933-2979)
    before = m.spec_comps[0]['factor'][0]['TW'].copy()
    m.GEM_iteration()
    assert np.isfinite(m.spec_comps[0]['factor'][0]['TW']).all() and before.shape


def test_sparse_model_on_kernel_spec_against_reference(tmp_path, monkeypatch):
    check_sparse_model(AllFakeKernels(), "float64", 1e-9, 1e-7, tmp_path, monkeypatch)


def test_sparsity_argument_forms(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    m = am.multiChanSourceF0Filter(audio=WAV, kernels=AllFakeKernels(), compute_dtype="float64",
                                   **dict(KW, sparsity=[3, 0, 5]))
    assert [m.spec_comps[j]['sparsity'] for j in range(3)] == [3, 0, 5]
    m = am.multiChanSourceF0Filter(audio=WAV, kernels=AllFakeKernels(), compute_dtype="float64",
                                   **dict(KW, sparsity=[3, 4]))   # neither 1 nor nbComps entries
    assert [m.spec_comps[j]['sparsity'] for j in range(3)] == [False, False, False]


def check_separation_and_powers(kernels, dtype, tol_pow, max_lsb, tmp_path, monkeypatch):
    """comp_spat_comp_power (with factor selection) and the Wiener separation of the source/filter
    model against the oracle, both loaded with the reference's final parameters."""
    monkeypatch.chdir(tmp_path)
    g = load()
    m = build(kernels, 1, dtype)
    m.spat_comps, m.spec_comps = structure_from(g, "final")
    np.random.seed(0)
    o = fo.OracleFASST(WAV, nbComps=3, nbNMFComps=2, spatial_rank=1, wlen=256, hopsize=64, iter_num=1)
    o.spat_comps, o.spec_comps = structure_from(g, "final")
    for j, specs, facs in ((0, [], []), (1, [1], [0]), (1, [1], [1]), (2, [], [])):
        a = m.comp_spat_comp_power(j, specs, facs)
        b = o.comp_spat_comp_power(j, specs, facs)
        assert np.abs(a - b).max() / np.abs(b).max() < tol_pow, (j, specs, facs)
    # a component of another source contributes nothing (audioModel.py:476-478)
    assert np.abs(m.comp_spat_comp_power(0, [1])).max() == 0
    noise = np.array(g["final_TW0_0"]).mean() * 0 + 1e-3 * np.ones(129)
    m.noise['PSD'], o.noise['PSD'] = noise.copy(), noise.copy()
    pcm = m.separate_comps_pcm()
    ref = o.separate_signals()
    for n in range(3):
        want = fo.pcm_from_float(ref[n], o.maxdata)
        diff = np.abs(pcm[n].astype(int) - want.astype(int))
        assert diff.max() <= max_lsb, (n, diff.max())


def test_separation_and_powers_on_kernel_spec(tmp_path, monkeypatch):
    check_separation_and_powers(AllFakeKernels(), "float64", 1e-10, 1, tmp_path, monkeypatch)


def sparse_structure(g, prefix):
    spat, spec = structure_from(g, prefix)
    for j in range(3):
        spec[j]['sparsity'] = 2   # sparsity=[2] applies to every component (audioModel.py:2766-2768)
    return spat, spec


def test_oracle_sparsity_reweighting_against_reference():
    g = load()
    for L in (1, 2, 5):
        assert_allclose(fo.median_filter(g["median_in"], length=L), g["median_%d" % L], rtol=0, atol=0)
    np.random.seed(0)
    m = fo.OracleFASST(WAV, nbComps=3, nbNMFComps=2, spatial_rank=1, wlen=256, hopsize=64,
                       iter_num=4)
    m.spat_comps, m.spec_comps = sparse_structure(g, "init")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ll = m.estim_param_a_post_model()
    assert_allclose(ll, g["sparse_logliks"], rtol=1e-9)
    check_snapshot(m, g, "sparse", 1e-9)
