#!/usr/bin/env python
"""Test infrastructure only -- NOT part of the product path.

Generates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (the py2 sources
under /root/reference, made importable by oracle/_py2shim.py) on small seeded
inputs.  These vectors pin oracle/ (the numpy restatement) and, through it, the
CUDA path.  Run in the authoring container only (it needs /root/reference):

    python oracle/make_golden.py

The GPU box never runs this; it only reads the committed .npz fixtures.
"""
import os
import sys
import warnings

import numpy as np
import scipy.io.wavfile as wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import _py2shim  # noqa: E402

warnings.simplefilter("ignore")


def synth_mix(seed, fs=8000, dur=0.6, nsrc=3, conv=False):
    """Small stereo int16 mixture: AR(2)-coloured, gated noise sources."""
    rng = np.random.default_rng(seed)
    L = int(fs * dur)
    mix = np.zeros((L, 2))
    for j in range(nsrc):
        e = rng.standard_normal(L)
        s = np.zeros(L)
        a1, a2 = 1.6 - 0.5 * j, -0.8 + 0.1 * j
        for t in range(2, L):
            s[t] = a1 * s[t - 1] + a2 * s[t - 2] + e[t]
        gate = (np.sin(2 * np.pi * (1.5 + j) * np.arange(L) / fs + j) > -0.3)
        s = s * gate / np.abs(s).max()
        if conv:
            for c in range(2):
                h = rng.standard_normal(8) * np.exp(-np.arange(8) / 2.0)
                mix[:, c] += np.convolve(s, h)[:L]
        else:
            th = (j + 1) * np.pi / (2.0 * (nsrc + 1))
            mix[:, 0] += np.sin(th) * s
            mix[:, 1] += np.cos(th) * s
    mix += 0.003 * rng.standard_normal(mix.shape)
    mix = 0.9 * mix / np.abs(mix).max()
    return fs, np.int16(np.round(mix * 32767))


def snapshot(model, prefix, out):
    for j, sc in model.spat_comps.items():
        out["%s_A%d" % (prefix, j)] = np.array(sc["params"])
    for k, sp in model.spec_comps.items():
        fac = sp["factor"][0]
        out["%s_FB%d" % (prefix, k)] = np.array(fac["FB"])
        out["%s_FW%d" % (prefix, k)] = np.array(fac["FW"])
        out["%s_TW%d" % (prefix, k)] = np.array(fac["TW"])


def run_fasst(ref, name, wav, conv, rank, iters, nbcomps=3, K=4, seed=0):
    am = ref["audioModel"]
    out = {}
    np.random.seed(seed)
    cls = am.MultiChanNMFConv if conv else am.MultiChanNMFInst_FASST
    model = cls(audio=wav, nbComps=nbcomps, nbNMFComps=K, spatial_rank=rank,
                wlen=256, hopsize=64, iter_num=iters, verbose=0,
                ann_PSD_lim=[None, None])
    if conv:
        model.makeItConvolutive()
    if name == "fasst_inst_r1":
        out["Cx"] = model.Cx
    out["ann0"] = np.array(model.noise["ann_PSD_lim"][0])
    out["ann1"] = np.array(model.noise["ann_PSD_lim"][1])
    snapshot(model, "init", out)
    # one E-step on the initial parameters (audioModel.py:580-764)
    model.noise["PSD"] = model.noise["ann_PSD_lim"][0]
    scp, mm, rpi = model.retrieve_subsrc_params()
    hRxx, hRxs, hRss, hWs, ll = model.compute_suff_stat(scp, mm)
    out["e0_hat_Rxs"], out["e0_hat_Rss"] = hRxs, hRss
    out["e0_hat_Ws"], out["e0_loglik"] = hWs, np.real(ll)
    # full GEM (audioModel.py:330-382); keep the state after iteration 1 too
    model.iter_num = 1
    ll1 = model.estim_param_a_post_model()
    snapshot(model, "it1", out)
    out["ll_it1"] = np.real(ll1)
    # restart from the same init for the full trajectory
    np.random.seed(seed)
    model = cls(audio=wav, nbComps=nbcomps, nbNMFComps=K, spatial_rank=rank,
                wlen=256, hopsize=64, iter_num=iters, verbose=0,
                ann_PSD_lim=[None, None])
    if conv:
        model.makeItConvolutive()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lls = model.estim_param_a_post_model()
    out["logliks"] = np.real(lls)
    snapshot(model, "final", out)
    out["noise_PSD_final"] = np.array(model.noise["PSD"])
    # separation (audioModel.py:1063-1236)
    outdir = "/tmp/pyfasst_golden_out_%s" % name
    os.makedirs(outdir, exist_ok=True)
    model.separate_spat_comps(dir_results=outdir)
    for n, f in enumerate(model.files["spat_comp"]):
        fs, y = wavfile.read(f)
        out["sep%d" % n] = y
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "logliks", out["logliks"])


def run_stft(ref):
    st = ref["stft"]
    rng = np.random.default_rng(7)
    x = rng.standard_normal(3001) * np.hanning(3001)
    tf = st.STFT(linFTLen=256, atomHopFactor=0.25, fs=8000)
    tf.computeTransform(x)
    X = np.array(tf.transfo)
    y = tf.invertTransform()
    tf2 = st.STFT(linFTLen=2048, atomHopFactor=0.25, fs=44100)
    x2 = rng.standard_normal(5000)
    tf2.computeTransform(x2)
    X2 = np.array(tf2.transfo)
    y2 = tf2.invertTransform()
    np.savez_compressed(os.path.join(GOLD, "stft.npz"), x=x, X=X, y=y,
                        freqs=tf.freq_stamps, times=tf.time_stamps,
                        x2=x2, X2=X2, y2=y2)
    print("stft", X.shape, np.abs(y - x).max(), X2.shape)


def run_inv(ref):
    sst = ref["signalTools"]
    rng = np.random.default_rng(11)
    d = np.abs(rng.standard_normal((2, 6, 5))) + 0.1
    o = 0.3 * (rng.standard_normal((6, 5)) + 1j * rng.standard_normal((6, 5)))
    d[:, 0, 0] = 0.0  # exercises the determinant clamp (signalTools.py:186-188)
    o[0, 0] = 0.0
    d[:, 0, 1] = 1e-6
    o[0, 1] = 1e-6
    i_d, i_o, det = sst.inv_herm_mat_2d(d, o)
    np.savez_compressed(os.path.join(GOLD, "inv2d.npz"), d=d, o=o, inv_d=i_d,
                        inv_o=i_o, det=det)


def run_simm(ref):
    simm = ref["SIMM"]
    rng = np.random.default_rng(3)
    F, N, NF0, P, K, R = 65, 40, 24, 6, 3, 5
    WF0 = np.abs(rng.standard_normal((F, NF0))) ** 2
    WF0 /= WF0.max(axis=0)
    WGAMMA = np.abs(rng.standard_normal((F, P)))
    SXR = np.abs(rng.standard_normal((F, N))) ** 2 + 1e-3
    SXL = np.abs(rng.standard_normal((F, N))) ** 2 + 1e-3
    init = dict(HGAMMA0=np.abs(rng.standard_normal((P, K))),
                HPHI0=np.abs(rng.standard_normal((K, N))),
                HF00=np.abs(rng.standard_normal((NF0, N))),
                WM0=np.abs(rng.standard_normal((F, R))),
                HM0=np.abs(rng.standard_normal((R, N))))
    out = dict(WF0=WF0, WGAMMA=WGAMMA, SXR=SXR, SXL=SXL, **init)
    # mono SIMM needs R == 1 (SIMM.py:388 broadcasts sumWM[R] against HM[R,N])
    init1 = dict(init)
    init1["WM0"] = init["WM0"][:, :1]
    init1["HM0"] = init["HM0"][:1]
    res = simm.SIMM(0.5 * (SXR + SXL), WF0, WGAMMA, numberOfFilters=K,
                    numberOfAccompanimentSpectralShapes=1,
                    numberOfIterations=4, verbose=False, **init1)
    for nm, a in zip(("HGAMMA", "HPHI", "HF0", "HM", "WM"), res):
        out["mono_" + nm] = a
    np.random.seed(5)  # Stereo_SIMM draws betaR from np.random (SIMM.py:581)
    res = simm.Stereo_SIMM(SXR, SXL, WF0, WGAMMA, numberOfFilters=K,
                           numberOfAccompanimentSpectralShapes=R,
                           numberOfIterations=4, verbose=False,
                           computeError=True, **init)
    for nm, a in zip(("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR",
                      "betaL", "HM", "WM", "recoError"), res):
        out["st_" + nm] = np.asarray(a)
    np.savez_compressed(os.path.join(GOLD, "simm.npz"), **out)
    print("simm alphaR", out["st_alphaR"], "err", out["st_recoError"][:3])


def run_lead_sep(ref):
    """The SIMM front / back end: slf.stft, slf.istft and SeparateLeadProcess.
    writeSeparatedSignals (SeparateLeadStereoTF.py:1762-1871, executed from its own source
    against a stand-in `self`: the module itself needs the compiled Viterbi extension)."""
    slf, simm = ref["slf"], ref["SIMM"]
    rng = np.random.default_rng(21)
    fs, L, wlen, hop = 8000, 3000, 256, 32
    t = np.arange(L) / fs
    sig = np.stack([np.sin(2 * np.pi * 440 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t))
                    + 0.3 * rng.standard_normal(L),
                    0.6 * np.sin(2 * np.pi * 440 * t) + 0.4 * rng.standard_normal(L)], axis=1)
    pcm = np.int16(np.round(0.8 * 32767 * sig / np.abs(sig).max()))
    wavfile.write(os.path.join(GOLD, "mix_lead.wav"), fs, pcm)
    scale = 1.2 * np.abs(pcm).max()
    data = np.double(pcm) / scale
    win = slf.sinebell(wlen)
    XR, Fr, Nt = slf.stft(data[:, 0], window=win, hopsize=hop, nfft=wlen, fs=fs)
    XL, _, _ = slf.stft(data[:, 1], window=win, hopsize=hop, nfft=wlen, fs=fs)
    yR = slf.istft(XR, window=win, hopsize=hop, nfft=wlen)
    F, N = XR.shape
    NF0, P, K, R = 20, 8, 3, 4
    WF0 = np.abs(rng.standard_normal((F, NF0))) ** 2
    WF0 /= WF0.max(axis=0)
    WGAMMA = np.abs(rng.standard_normal((F, P)))
    init = dict(HGAMMA0=np.abs(rng.standard_normal((P, K))),
                HPHI0=np.abs(rng.standard_normal((K, N))),
                HF00=np.abs(rng.standard_normal((NF0, N))),
                WM0=np.abs(rng.standard_normal((F, R))),
                HM0=np.abs(rng.standard_normal((R, N))))
    np.random.seed(9)
    res = simm.Stereo_SIMM(np.abs(XR) ** 2, np.abs(XL) ** 2, WF0, WGAMMA, numberOfFilters=K,
                           numberOfAccompanimentSpectralShapes=R, numberOfIterations=3,
                           verbose=False, **init)
    names = ("alphaR", "alphaL", "HGAMMA", "HPHI", "HF0", "betaR", "betaL", "HM", "WM")
    params = dict(zip(names, res[:9]))
    params.update(WF0=WF0, WGAMMA=WGAMMA)
    written = {}

    class Wav(object):
        @staticmethod
        def write(name, rate, arr):
            written[name] = np.array(arr)

    class Self(object):
        pass
    obj = Self()
    obj.SIMMParams = params
    obj.stftParams = dict(windowSizeInSamples=wlen, hopsize=hop, NFT=wlen)
    obj.XR, obj.XL = XR, XL
    obj.tfrepresentation = "stft"
    obj.scaleData, obj.dataType, obj.fs = scale, pcm.dtype, fs
    obj.files = dict(voc_output_file="out_lead.wav", mus_output_file="out_acc.wav")
    ns = dict(np=np, slf=slf, wav=Wav, eps=10 ** -9, knownTransfos=["stft"])
    exec(ref["writeSeparatedSignals"], ns)
    ns["writeSeparatedSignals"](obj, suffix=".wav")
    hann = ref["filter"].generateHannBasis(numberFrequencyBins=1025, sizeOfFourier=2048, Fs=44100,
                                           frequencyScale='linear', numberOfBasis=30, overlap=.75)
    hann2 = ref["filter"].generateHannBasis(129, 256, 8000, numberOfBasis=8)
    out = dict(pcm=pcm, XR=XR, XL=XL, F=Fr, N=Nt, yR=yR, voc=written["out_lead.wav"],
               hann_1025_30=hann, hann_129_8=hann2,
               mus=written["out_acc.wav"], **init)
    out.update({"p_" + k: np.asarray(v) for k, v in params.items()})
    np.savez_compressed(os.path.join(GOLD, "lead_sep.npz"), **out)
    print("lead_sep", XR.shape, written["out_lead.wav"].shape, np.abs(written["out_lead.wav"]).max())


def run_nmf(ref):
    """tools/nmf.py: NMF_decomposition and NMF_decomp_init (the step before GEM)."""
    nmf = ref["nmf"]
    rng = np.random.default_rng(8)
    F, N, K = 65, 120, 6
    Wt = np.abs(rng.standard_normal((F, K))) ** 2
    Ht = np.abs(rng.standard_normal((K, N))) ** 2
    SX = Wt @ Ht * (1 + 0.1 * rng.standard_normal((F, N))) ** 2 + 1e-6
    np.random.seed(12)
    W0 = np.random.randn(F, K) ** 2          # the draws NMF_decomposition makes (nmf.py:34-35)
    H0 = np.random.randn(K, N) ** 2
    np.random.seed(12)
    W1, H1 = nmf.NMF_decomposition(SX, nbComps=K, niter=5)
    Wi = np.abs(rng.standard_normal((F, K))) + 0.1
    Hi = np.abs(rng.standard_normal((K, N))) + 0.1
    W2, H2 = nmf.NMF_decomp_init(SX, nbComps=K, niter=5, Winit=Wi, Hinit=Hi)
    W3, H3 = nmf.NMF_decomp_init(SX, nbComps=K, niter=5, Winit=Wi, Hinit=Hi.T, updateW=False)
    np.savez_compressed(os.path.join(GOLD, "nmf.npz"), SX=SX, W0=W0, H0=H0, W1=W1, H1=H1, Wi=Wi,
                        Hi=Hi, W2=W2, H2=H2, W3=W3, H3=H3)
    print("nmf", W1.shape, H1.shape, float(np.abs(W3 - Wi).max()))


def run_viterbi():
    """The reference's Cython tracker (compiled by oracle/build_ref.py) on small HMMs, ties
    included (integer log-probabilities) and on a melody-like problem built as runViterbi does
    (SeparateLeadStereoTF.py:1183-1215)."""
    from oracle import build_ref
    trk = build_ref.load()
    rng = np.random.default_rng(17)
    out = {}
    for tag, (S, N, ties) in {"a": (7, 40, False), "b": (33, 300, False), "c": (12, 200, True),
                              "d": (1, 5, False)}.items():
        dens = rng.standard_normal((S + 1, N))
        prior = np.log(np.ones(S + 1) / (S + 1))
        trans = np.log(rng.random((S + 1, S + 1)) + 1e-3)
        if ties:  # exactly equal candidates: the smallest predecessor index must win
            dens, trans = np.round(dens), np.round(trans)
        out.update({tag + "_dens": dens, tag + "_prior": prior, tag + "_trans": trans,
                    tag + "_path": trk.viterbiTracking(S, N, dens, prior, trans)})
    # melody-like: banded Toeplitz transitions, log of a sparse non-negative HF0 (with -inf)
    NF0, N, step = 60, 400, 4
    t = np.exp(-np.floor(np.arange(NF0) / step))
    cut = min(NF0, 2 * 5 * step)
    t[cut:] = t[cut - 1]
    T = np.zeros([NF0 + 1, NF0 + 1])
    b = np.arange(NF0)
    T[:NF0, :NF0] = t[np.abs(b[None, :] - b[:, None])]
    T[:NF0, NF0], T[NF0, :NF0], T[NF0, NF0] = t[cut - 1] * 1e-90, t[cut - 1] * 1e-80, t[cut - 1] * 1e-100
    T = T / T.sum(axis=1)[:, None]
    HF0 = np.abs(rng.standard_normal((NF0, N))) ** 4
    HF0[rng.random((NF0, N)) < 0.2] = 0.0
    with np.errstate(divide="ignore"):
        logH = np.zeros([NF0 + 1, N])
        logH[:NF0] = np.log(HF0)
    logH[NF0] = -100
    prior = np.log(np.ones(NF0 + 1) / (NF0 + 1.0))
    out.update(m_dens=logH, m_prior=prior, m_trans=np.log(T),
               m_path=trk.viterbiTracking(NF0, N, logH, prior, np.log(T)))
    np.savez_compressed(os.path.join(GOLD, "viterbi.npz"), **out)
    print("viterbi", {k: v[:8] for k, v in out.items() if k.endswith("path")})


def run_melody(ref):
    """SeparateLeadProcess.runViterbi and initiateHF0WithIndexBestPath
    (SeparateLeadStereoTF.py:1150-1368), executed from their own source against a stand-in
    `self`, with the reference's Cython tracker compiled by oracle/build_ref.py."""
    import tempfile
    from oracle import build_ref
    trk = build_ref.load()
    rng = np.random.default_rng(23)
    NF0, N, step, fs, hop = 61, 250, 4, 8000, 32
    F0Table = 100.0 * 2 ** (np.arange(NF0) / (12.0 * step))
    HF0 = np.abs(rng.standard_normal((NF0, N))) ** 6
    line = np.clip((30 + 20 * np.sin(np.arange(N) / 25.0)).astype(int), 0, NF0 - 1)
    HF0[line, np.arange(N)] += 50.0
    HF0[rng.random((NF0, N)) < 0.15] = 0.0
    HF0[:, 100:110] = 0.0      # silent frames: the log floor (:1205)
    HF0[0, 200:205] = 1e4      # the decoded line touches state 0 -> "no melody" (:1311, :1364)
    out = {}
    for tag, search in (("full", (None, None)), ("band", (130.0, 200.0))):
        class Self(object):
            pass
        obj = Self()
        obj.SIMMParams = dict(HF0=HF0.copy(), NF0=NF0, chirpPerF0=1, minF0=100, maxF0=float(
            F0Table[-1]) + 1, stepNotes=step, F0Table=F0Table)
        obj.trackingParams = dict(minF0search=search[0] or 100, maxF0search=search[1] or 1e9)
        obj.stftParams = dict(hopsize=float(hop))
        obj.fs, obj.N, obj.verbose, obj.scopeAllowedHF0 = fs, N, False, 4.0
        obj.computeNFrames = lambda: N
        tmp = tempfile.mkdtemp()
        obj.files = dict(pitch_output_file=os.path.join(tmp, "pitches.txt"))
        ns = dict(np=_py2shim.OldNumpy(), viterbiTrackingArray=trk.viterbiTracking, eps=10 ** -9)
        exec(ref["runViterbi"], ns)
        exec(ref["initiateHF0WithIndexBestPath"], ns)
        ns["runViterbi"](obj)
        ns["initiateHF0WithIndexBestPath"](obj)
        out.update({tag + "_path": obj.indexBestPath, tag + "_freq": obj.freqMelody,
                    tag + "_HF00": obj.SIMMParams["HF00"],
                    tag + "_pitches": np.loadtxt(obj.files["pitch_output_file"])})
    np.savez_compressed(os.path.join(GOLD, "melody.npz"), HF0=HF0, F0Table=F0Table, NF0=NF0,
                        N=N, stepNotes=step, fs=fs, hopsize=hop, **out)
    print("melody", out["full_path"][:10], out["band_path"][:10], out["full_HF00"].sum())


def run_wf0(ref):
    """The glottal-source F0 dictionary: generate_WF0_TR_chirped with the
    reference's own STFT object (what SeparateLeadProcess.computeWF0 runs for
    tfrepresentation='stft', SeparateLeadStereoTF.py:661-684), with and without chirps, and the
    older generate_WF0_chirped (separateLeadFunctions.py:237-345); the .npz cache names the
    reference writes are recorded too."""
    import tempfile
    slf, st, ut = ref["slf"], ref["stft"], ref["utils"]
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        fs, nft = 8000, 256
        tr = st.STFT(linFTLen=nft, atomHopFactor=0.25, winFunc=ut.sqrt_blackmanharris, fs=fs)
        t1, w1, _ = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=2, Ot=0.5,
                                                perF0=1, depthChirpInSemiTone=0.5, loadWF0=False)
        t2, w2, _ = slf.generate_WF0_TR_chirped(tr, minF0=100, maxF0=800, stepNotes=1, Ot=0.5,
                                                perF0=3, depthChirpInSemiTone=0.5, loadWF0=False)
        t3, w3 = slf.generate_WF0_chirped(100, 800, fs, Nfft=256, stepNotes=1, lengthWindow=256,
                                          Ot=0.5, perF0=2, depthChirpInSemiTone=.15,
                                          loadWF0=False, analysisWindow='sinebell')
        # a hop that is not ftlen / 4 and a window shorter than the transform would need another
        # transform class; the reference's STFT ties window length to linFTLen
        tr2 = st.STFT(linFTLen=512, atomHopFactor=0.125, winFunc=np.hanning, fs=16000)
        t4, w4, _ = slf.generate_WF0_TR_chirped(tr2, minF0=60, maxF0=500, stepNotes=1, Ot=0.25,
                                                perF0=2, depthChirpInSemiTone=0.5, loadWF0=False)
        names = sorted(os.listdir("."))
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "wf0.npz"), t1=t1, w1=w1, t2=t2, w2=w2, t3=t3, w3=w3,
                        t4=t4, w4=w4, cache_names=np.array(names))
    print("wf0", w1.shape, w2.shape, w3.shape, w4.shape, names)


def run_sourcefilter(ref, wav):
    """The reference's multiChanSourceF0Filter (audioModel.py:2551-3014) run as it is: two
    sources with a two-factor (glottal dictionary x smooth filters) spectral component sharing
    one dictionary object (quirk Q11), one residual NMF component.  Its
    SeparateLeadStereoTF import is a stub in the shim; only `SLS.slf` is needed."""
    import tempfile
    am = ref["audioModel"]
    am.SLS.slf = ref["slf"]
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    out = {}
    try:
        def build(iters):
            m = am.multiChanSourceF0Filter(
                audio=wav, nbComps=3, nbNMFResComps=2, nbFilterComps=6, nbFilterWeigs=[3, ],
                minF0=100, maxF0=400, stepnoteF0=1, chirpPerF0=1, spatial_rank=1, sparsity=None,
                wlen=256, hopsize=64, iter_num=iters, verbose=0, ann_PSD_lim=[None, None])
            # _initialize_structures(seed=None) reseeds from the OS: redo it with a fixed seed
            m._initialize_structures(seed=5)
            return m
        model = build(1)
        out["F0Table"] = model.F0Table
        out["sourceFreqComps0"] = np.array(model.sourceFreqComps)
        snapshot_sf(model, "init", out)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out["ll_it1"] = np.real(model.estim_param_a_post_model())
        snapshot_sf(model, "it1", out)
        model = build(5)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out["logliks"] = np.real(model.estim_param_a_post_model())
        snapshot_sf(model, "final", out)
        out["shared_final"] = np.array(model.sourceFreqComps)
        assert model.spec_comps[0]["factor"][0]["FB"] is model.spec_comps[1]["factor"][0]["FB"]
        # the same with the sparsity re-weighting of the source activations (:2933-3014):
        # median filter of length 2 for every component (a one-element list applies to all)
        def build_sparse(iters):
            m = am.multiChanSourceF0Filter(
                audio=wav, nbComps=3, nbNMFResComps=2, nbFilterComps=6, nbFilterWeigs=[3, ],
                minF0=100, maxF0=400, stepnoteF0=1, chirpPerF0=1, spatial_rank=1, sparsity=[2, ],
                wlen=256, hopsize=64, iter_num=iters, verbose=0, ann_PSD_lim=[None, None])
            m._initialize_structures(seed=5)
            return m
        model = build_sparse(4)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out["sparse_logliks"] = np.real(model.estim_param_a_post_model())
        snapshot_sf(model, "sparse", out)
        x = np.random.RandomState(3).rand(40)
        out["median_in"] = x
        for L in (1, 2, 5):
            out["median_%d" % L] = am.st.medianFilter(x, length=L)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "fasst_sourcefilter.npz"), **out)
    print("sourcefilter logliks", out["logliks"], "NF0+1 =", out["sourceFreqComps0"].shape)


def snapshot_sf(model, prefix, out):
    for j, sc in model.spat_comps.items():
        out["%s_A%d" % (prefix, j)] = np.array(sc["params"])
    for k, sp in model.spec_comps.items():
        for fi, fac in sp["factor"].items():
            for m in ("FB", "FW", "TW"):
                out["%s_%s%d_%d" % (prefix, m, k, fi)] = np.array(fac[m])


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = _py2shim.load()
    fs, mix = synth_mix(1)
    wav = os.path.join(GOLD, "mix_inst.wav")
    wavfile.write(wav, fs, mix)
    fs, mixc = synth_mix(2, conv=True)
    wavc = os.path.join(GOLD, "mix_conv.wav")
    wavfile.write(wavc, fs, mixc)
    run_inv(ref)
    run_stft(ref)
    run_fasst(ref, "fasst_inst_r1", wav, conv=False, rank=1, iters=6)
    run_fasst(ref, "fasst_inst_r2", wav, conv=False, rank=2, iters=6)
    run_fasst(ref, "fasst_conv_r1", wavc, conv=True, rank=1, iters=6)
    run_fasst(ref, "fasst_conv_r2", wavc, conv=True, rank=2, iters=6, nbcomps=2)
    run_simm(ref)
    run_lead_sep(ref)
    run_nmf(ref)
    run_viterbi()
    run_melody(ref)
    run_wf0(ref)
    run_sourcefilter(ref, wav)


if __name__ == "__main__":
    main()
