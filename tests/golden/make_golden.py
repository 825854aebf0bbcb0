#!/usr/bin/env python
"""Generate the golden fixtures in ``tests/golden/`` FROM THE REFERENCE ITSELF.

Runs only in the build container (needs ``/root/reference``): the reference's Python
layers are imported by ``oracle/ref_loader.py`` (in-memory py3 patches, compiled
``general`` module from the reference's own .pyx/.c), inputs are produced from fixed
seeds, and inputs + reference outputs are stored as small ``.npz`` files.  The fixtures
travel to the GPU box; the reference does not.

    python tests/golden/make_golden.py            # regenerate everything (~2 min)

Every array named ``ref_*`` is an output of the reference's code; everything else is input.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")

from oracle import build_ref, ref_loader  # noqa: E402

DT_LONG = 1765.55929      # Kepler long cadence, scripts/flare_detection_efficiency.py:273
DT_SHORT = 58.84876


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024.))


def ref_flare(bf, ts, t0, amp, tg, te):
    return bf.Flare(ts, amp=1.).model({"t0": t0, "amp": amp, "taugauss": tg, "tauexp": te})


def gold_scalars(bf):
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-50, 50, 40), [-np.inf, -np.inf, 0., 3., np.inf]])
    ys = np.concatenate([rng.uniform(-50, 50, 40), [-np.inf, 2., -np.inf, 3., 1.]])
    lp = np.array([bf.logplus(x, y) for x, y in zip(xs, ys)])
    lm = np.array([bf.logminus(x, y) for x, y in zip(xs, ys)])
    lts_in, lts_t, lts_out = [], [], []
    for n in (2, 3, 5, 10, 17):
        t = np.sort(rng.uniform(0, 100, n))
        lx = rng.uniform(-30, 30, n)
        if n == 5:
            lx[1] = -np.inf
        if n == 10:
            lx[:] = -np.inf
        lts_in.append(lx), lts_t.append(t), lts_out.append(bf.logtrapz(lx, t))
    save("scalars", x=xs, y=ys, ref_logplus=lp, ref_logminus=lm,
         lt_sizes=np.array([len(a) for a in lts_in]), lt_lx=np.concatenate(lts_in), lt_t=np.concatenate(lts_t),
         ref_logtrapz=np.array(lts_out),
         kat_logplus_1_2=np.array(bf.logplus(1, 2)),
         kat_logtrapz=np.array(bf.logtrapz(np.array([0, -1, -np.inf]), np.array([0., 1, 2]))))


def gold_elimination():
    """log_marg_amp_full_C / log_marg_amp_except_final_C called directly through ctypes on
    the reference's own compiled C (oracle/_ref/liblogmarg.so)."""
    built = build_ref.build()
    L = C.CDLL(built["liblogmarg"])
    dp = C.POINTER(C.c_double)
    L.log_marg_amp_full_C.restype = C.c_double
    L.log_marg_amp_full_C.argtypes = [C.c_int, dp, dp, C.c_double, C.c_uint]
    L.log_marg_amp_except_final_C.restype = C.c_double
    L.log_marg_amp_except_final_C.argtypes = [C.c_int, dp, dp, C.c_double]
    rng = np.random.default_rng(7)
    NM = 8
    n = 400
    Ms = np.zeros((n, NM, NM))
    Ds = np.zeros((n, NM))
    nms = np.zeros(n, dtype=np.int32)
    sig = np.zeros(n)
    out = np.zeros((n, 3))
    for t in range(n):
        nm = int(rng.integers(2, NM + 1))
        A = rng.standard_normal((nm + 3, nm))
        if t % 5 == 0:   # Hilbert-like, as the monomial Gram matrices on the path are
            x = np.linspace(0, 1, 55)
            A = np.array([x ** p for p in range(nm - 1)] + [np.exp(-x * rng.uniform(1, 20))]).T
        M = np.triu(A.T @ A)
        D = rng.standard_normal(nm) * rng.uniform(0.1, 30)
        if t % 5 == 0:
            D = A.T @ (rng.standard_normal(55) + 3 * A[:, -1])
        s = 1.0 if t % 2 else rng.uniform(0.5, 2)
        if t == 17:
            D[1] = np.inf
        if t == 18:
            M[0, 0] = np.nan
        Mc = np.ascontiguousarray(M)
        nms[t], sig[t] = nm, s
        Ms[t, :nm, :nm], Ds[t, :nm] = M, D
        out[t, 0] = L.log_marg_amp_full_C(nm, Mc.ctypes.data_as(dp), D.ctypes.data_as(dp), s, 0)
        out[t, 1] = L.log_marg_amp_full_C(nm, Mc.ctypes.data_as(dp), D.ctypes.data_as(dp), s, 1)
        out[t, 2] = L.log_marg_amp_except_final_C(nm, Mc.ctypes.data_as(dp), D.ctypes.data_as(dp), s)
    save("elimination", M=Ms, D=Ds, nm=nms, sigma=sig, ref_full_lhr0=out[:, 0], ref_full_lhr1=out[:, 1],
         ref_except_final=out[:, 2])


def _force_sk(sk):
    import bayesflare.stats.bayes as bb
    orig = bb.estimate_noise_ps
    bb.estimate_noise_ps = lambda lc, estfrac=0.5: (sk, sk * sk, None)
    return bb, orig


def gold_kat200(bf):
    """Deterministic N=200 curve with *varying* per-sample errors (general-variance path),
    noise sigma forced to 0.9 so the fixture is independent of the noise estimator.
    Every hypothesis type, all samples including the truncated-window edges."""
    N = 200
    i = np.arange(N)
    clc = np.sin(0.37 * i) + 0.5 * np.cos(0.011 * i ** 2) + 3 * np.exp(-np.maximum(i - 100, 0) / 2.) * (i >= 100)
    cle = 0.1 + 0.05 * np.sin(0.05 * i) ** 2
    ts = i * DT_LONG
    lc = ref_loader.RefLightcurve(ts, clc, cle)
    bb, orig = _force_sk(0.9)
    out = dict(cts=ts, clc=clc, cle=cle, sk=np.array(0.9))
    try:
        dt = ts[1] - ts[0]
        cases = {
            "flare": (bf.Flare, {}, {"t0": (np.inf,), "taugauss": (0, 5400, 4), "tauexp": (1800, 10800, 4), "amp": (1.,)}, True),
            "impulse": (bf.Impulse, {}, {"t0": (0, 54 * dt, 55), "amp": (1.,)}, False),
            "impulse1": (bf.Impulse, {}, {"t0": (np.inf,), "amp": (1.,)}, False),
            "expdecay": (bf.Expdecay, {}, {"t0": (np.inf,), "tauexp": (0.0, 900, 3), "amp": (1.,)}, True),
            "expdecay_rev": (bf.Expdecay, {"reverse": True}, {"t0": (np.inf,), "tauexp": (0.0, 900, 3), "amp": (1.,)}, True),
            "step": (bf.Step, {}, {"t0": (np.inf,), "amp": (1.,)}, False),
            "transit": (bf.Transit, {}, {"t0": (np.inf,), "sigmag": (1800, 7200, 3), "tauf": (0, 7200, 4), "amp": (1.,)}, True),
            "flare_full": (bf.Flare, {}, {"t0": (np.inf,), "taugauss": (0, 5400, 3), "tauexp": (1800, 10800, 3), "amp": (1.,)}, False),
        }
        for name, (cls, kw, pr, hr) in cases.items():
            M = cls(lc.cts, amp=1, paramranges=dict(pr), **kw)
            B = bf.Bayes(lc, M)
            B.bayes_factors_marg_poly_bgd(halfrange=hr)
            out["ref_lnB_" + name] = B.lnBmargAmp
            out["ref_marg_" + name] = B.marginalise_full().lnBmargAmp
            if name == "flare":
                out["ref_bgonly"] = B.bayes_factors_marg_poly_bgd_only()
                out["ref_noise_ev"] = np.array(B.noise_ev)
        # a different window / polynomial order
        M = bf.Flare(lc.cts, amp=1, paramranges=dict(cases["flare"][2]))
        B = bf.Bayes(lc, M)
        B.bayes_factors_marg_poly_bgd(bglen=21, bgorder=2)
        out["ref_lnB_flare_w21_o2"] = B.lnBmargAmp
        out["ref_bgonly_w21_o2"] = B.bayes_factors_marg_poly_bgd_only(bglen=21, bgorder=2)
    finally:
        bb.estimate_noise_ps = orig
    save("kat200", **out)


def sim_curve(bf, N, seed, dt=DT_LONG, flares=(), sinamp=0.0, nstd=1.0):
    rng = np.random.RandomState(seed)
    ts = np.arange(N) * dt
    clc = nstd * rng.randn(N)
    if sinamp:
        clc = clc + sinamp * np.sin(2 * np.pi * ts / (3.3 * 86400.) + 0.7)
    for (idx, amp, tg, te) in flares:
        clc = clc + ref_flare(bf, ts, ts[idx], amp, tg, te)
    clc = clc - np.median(clc)      # Lightcurve.dcoffset, data/data.py:271-277
    return ts, clc


def gold_detector(bf):
    out = {}
    # (a) N=1000, both noise estimators, default detector
    ts, clc = sim_curve(bf, 1000, 101, flares=[(400, 12., 1000., 3000.), (700, 6., 500., 6000.)], sinamp=2.0)
    lc = ref_loader.RefLightcurve(ts, clc)
    out.update(a_cts=ts, a_clc=clc)
    for meth in ("powerspectrum", "tailveto"):
        det = bf.OddsRatioDetector(lc, noiseestmethod=meth)
        lnO, tso = det.oddsratio()
        out["ref_a_lnO_" + meth] = np.array(lnO)
        fl, nfl, mx = det.thresholder(lnO, 16.5, expand=1)
        out["ref_a_segs_" + meth] = np.array(fl, dtype=np.int64).reshape(-1, 2)
        out["ref_a_max_" + meth] = np.array(mx, dtype=float).reshape(-1, 2)
    # noise sigma as the reference computes it (bayes.py:235-246)
    from copy import deepcopy
    tmp = deepcopy(lc)
    tmp.detrend(method="savitzkygolay", nbins=55, order=4)
    out["ref_a_sk_ps"] = np.array(bf.estimate_noise_ps(tmp, estfrac=0.5)[0])
    out["ref_a_sk_tv"] = np.array(bf.estimate_noise_tv(tmp.clc, sigma=1.0)[0])
    out["ref_a_sgfit"] = bf.savitzky_golay(lc.clc, 55, 4)
    # (b) the scripts' detector (55-point impulse grid, tailveto), N=1639
    # scripts/flare_detection_efficiency.py:495-507
    N = int(np.ceil(2893536. / DT_LONG))
    ts, clc = sim_curve(bf, N, 202, flares=[(800, 9., 2000., 4000.)])
    lc = ref_loader.RefLightcurve(ts, clc)
    det = bf.OddsRatioDetector(lc, bglen=55, bgorder=4, noiseestmethod="tailveto", tvsigma=1.0,
                               flareparams={"taugauss": (0, 1.5 * 3600, 10), "tauexp": (0.5 * 3600, 3 * 3600, 10)},
                               noisepoly=True, noiseimpulse=True,
                               noiseimpulseparams={"t0": (0, 54 * DT_LONG, 55)},
                               noiseexpdecay=True, noiseexpdecayparams={"tauexp": (0.0, 0.25 * 3600, 3)},
                               noiseexpdecaywithreverse=True, ignoreedges=True)
    lnO, tso = det.oddsratio()
    out.update(b_cts=ts, b_clc=clc, ref_b_lnO=np.array(lnO))
    fl, nfl, mx = det.thresholder(lnO, 10.0, expand=2)
    out["ref_b_segs"] = np.array(fl, dtype=np.int64).reshape(-1, 2)
    out["ref_b_max"] = np.array(mx, dtype=float).reshape(-1, 2)
    # (c) step hypothesis on, edges kept, positive impulse
    ts, clc = sim_curve(bf, 400, 303, flares=[(200, 10., 800., 2500.)])
    lc = ref_loader.RefLightcurve(ts, clc)
    det = bf.OddsRatioDetector(lc, noiseestmethod="tailveto", noisestep=True, ignoreedges=False)
    det.set_noise_impulse(noiseimpulse=True, noiseimpulseparams={"t0": (np.inf,)}, positive=True)
    lnO, tso = det.oddsratio()
    out.update(c_cts=ts, c_clc=clc, ref_c_lnO=np.array(lnO))
    save("detector", **out)


def gold_config1(bf):
    """BASELINE config 1 shape: N=4400, default detector; the SURVEY 8(c) end-to-end KAT."""
    np.random.seed(1)
    N = 4400
    ts = np.arange(N) * DT_LONG
    clc = np.random.randn(N) + ref_flare(bf, ts, ts[2200], 8., 1000., 3000.)
    lc = ref_loader.RefLightcurve(ts, clc)
    out = dict(cts=ts, clc=clc)
    for meth in ("powerspectrum", "tailveto"):
        det = bf.OddsRatioDetector(lc, noiseestmethod=meth)
        lnO, tso = det.oddsratio()
        out["ref_lnO_" + meth] = np.array(lnO)
        fl, nfl, mx = det.thresholder(lnO, 16.5, expand=1)
        out["ref_segs_" + meth] = np.array(fl, dtype=np.int64).reshape(-1, 2)
        out["ref_max_" + meth] = np.array(mx, dtype=float).reshape(-1, 2)
        print(meth, "max lnO", np.max(lnO), int(np.argmax(lnO)), fl)
    save("config1", **out)


def gold_short_cadence(bf):
    """BASELINE config 2 shape at reduced length: short cadence (58.8 s), N=6000."""
    ts, clc = sim_curve(bf, 6000, 404, dt=DT_SHORT, flares=[(2500, 6., 300., 1500.), (4000, 15., 100., 600.)])
    lc = ref_loader.RefLightcurve(ts, clc)
    det = bf.OddsRatioDetector(lc, noiseestmethod="powerspectrum")
    lnO, tso = det.oddsratio()
    save("shortcad", cts=ts, clc=clc, ref_lnO=np.array(lnO))


def gold_pe(bf):
    """BASELINE config 4 (reduced grid): scripts/parameter_estimation_example.py:27-64."""
    rng = np.random.RandomState(505)
    ts = np.arange(0., 2893536., DT_LONG, dtype="float64")
    clc = rng.randn(len(ts))
    t0, amp, taug, taue = 400000., 20.0, 1060.0, 2168.0
    clc = clc + ref_flare(bf, ts, t0, amp, taug, taue)
    lc = ref_loader.RefLightcurve(ts, clc)
    idxt0 = int(t0 / lc.dt())
    out = dict(cts=ts, clc=clc, t0=np.array(t0), idxt0=np.array(idxt0))
    for bgorder in (3, 4, 2):
        pe = bf.ParameterEstimationGrid("flare", lc)
        pe.lightcurve_chunk(idxt0, 55)
        amprange = (0., 1.6 * (np.amax(pe.lightcurve.clc) - np.amin(pe.lightcurve.clc)), 9)
        pe.set_grid(ranges={"taugauss": (0., 3600., 12), "tauexp": (0., 12000., 10), "amp": amprange, "t0": (t0,)})
        pe.calculate_posterior(bgorder=bgorder)
        out["ref_sigma"] = np.array(pe.noiseSigma)
        out["amprange"] = np.array(amprange)
        out["ref_post_o%d" % bgorder] = pe.posterior
        if bgorder == 3:
            pe.marginalise_all()
            for p in ("taugauss", "tauexp", "amp"):
                out["ref_margpost_" + p] = pe.margposteriors[p]
            mp, mpar = pe.maximum_posterior()
            out["ref_maxpost"] = np.array(mp)
            out["ref_maxpost_params"] = np.array([mpar[p] for p in ("t0", "tauexp", "taugauss", "amp")], dtype=float)
            out["ref_snr"] = np.array(pe.maximum_posterior_snr())
            out["ref_marg2d_taugauss_tauexp"] = pe.marginalised_posterior_2D(["taugauss", "tauexp"])
            out["ref_marg2d_amp_taugauss"] = pe.marginalised_posterior_2D(["amp", "taugauss"])
    pe = bf.ParameterEstimationGrid("flare", lc)
    pe.lightcurve_chunk(idxt0, 55)
    pe.set_grid(ranges={"taugauss": (0., 3600., 6), "tauexp": (0., 12000., 5), "amp": amprange, "t0": (t0,)})
    pe.calculate_posterior(margbackground=False)
    out["ref_post_nobg"] = pe.posterior
    # a Savitzky-Golay detrended chunk: the reference filters every model the same way (bayes.py:1018-1021).
    # (set_detrend stores the window as detrend_length but calculate_posterior reads detrend_nbins, which
    # stays 0 -- data.py:373 vs bayes.py:1020 -- so the window is set by hand here, as a user would have to.)
    pe = bf.ParameterEstimationGrid("flare", lc)
    pe.lightcurve_chunk(idxt0, 55)
    pe.lightcurve.detrend(method="savitzkygolay", nbins=21, order=3)
    pe.lightcurve.detrend_nbins = 21
    pe.set_grid(ranges={"taugauss": (0., 3600., 6), "tauexp": (0., 12000., 5), "amp": amprange, "t0": (t0,)})
    pe.calculate_posterior(bgorder=2)
    out["ref_post_sg21_o2"] = pe.posterior
    out["ref_sg_chunk_clc"] = pe.lightcurve.clc.copy()
    save("pegrid", **out)


def _force_sk_detector(sk):
    """Make the reference's noise estimators return a given sigma (so that the golden does not depend on the
    un-pinned mlab.psd shim): bayes.py:235-246 calls estimate_noise_ps / estimate_noise_tv by module name."""
    import bayesflare.stats.bayes as rb
    rb.estimate_noise_ps = lambda lc, estfrac=0.5: (sk, sk)
    rb.estimate_noise_tv = lambda d, sigma=1.0: (sk, 0.0)


def gold_trend(bf):
    """Data whose polynomial-like component is hundreds of sigma high -- the regime where the reference's own
    elimination of raw monomials loses digits (hard part 1).  The fixtures are
    the outputs of the REFERENCE ITSELF (np.correlate and all), so a test can tell the reference's rounding noise
    from an error of ours: (a) slow sinusoidal trend 400 sigma high, (b) sigma = 50 on a DC offset of 3000."""
    out = {}
    rng = np.random.RandomState(911)
    N = 1100
    ts = np.arange(N) * DT_LONG
    base = rng.randn(N) + ref_flare(bf, ts, ts[500], 9., 1500., 3500.)
    a = base + 400.0 * np.sin(2 * np.pi * ts / (6.3 * 86400.))
    N2 = 600
    b = 50.0 * rng.randn(N2) + 3000.0 + ref_flare(bf, ts[:N2], ts[300], 400., 1500., 3500.)
    import importlib
    rb = importlib.import_module("bayesflare.stats.bayes")
    keep = (rb.estimate_noise_ps, rb.estimate_noise_tv)
    try:
        for name, clc, sk in (("a", a, 1.0), ("b", b, 50.0)):
            _force_sk_detector(sk)
            lc = ref_loader.RefLightcurve(ts[:len(clc)], clc)
            det = bf.OddsRatioDetector(lc)
            lnO, _ = det.oddsratio()
            out[name + "_cts"], out[name + "_clc"], out[name + "_sk"] = ts[:len(clc)], clc, np.array(sk)
            out["ref_" + name + "_lnO"] = np.array(lnO)
    finally:
        rb.estimate_noise_ps, rb.estimate_noise_tv = keep
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        np.show_config()
    out["numpy_config"] = np.array(buf.getvalue()[:4000])
    save("trend", **out)


def gold_transit5(bf):
    """BASELINE config 5's transit-model odds at the sweep's shape: N = 1639 (tlength 2893536 s,
    scripts/flare_detection_efficiency.py:210-216), sinusoidal variation, one injected transit
    (:291-312, 444-453), Bayes(lc, Transit).bayes_factors_marg_poly_bgd with the scripts' settings (bglen 55,
    bgorder 4, tail-veto noise).  Transit.prior is executable only with the two undefined names fixed
    (oracle/ref_loader.py PATCHES)."""
    rng = np.random.RandomState(515)
    N = int(np.ceil(2893536. / DT_LONG))
    ts = np.arange(N) * DT_LONG
    clc = rng.randn(N) + 2.5 * np.sin(2 * np.pi * ts / (9.1 * 86400.) + 1.1)
    tr = bf.Transit(ts, amp=1.)
    clc = clc + tr.model({"t0": ts[700], "amp": 6.0, "sigmag": 2400., "tauf": 7000.})
    clc = clc - np.median(clc)
    lc = ref_loader.RefLightcurve(ts, clc)
    ranges = {"t0": (np.inf,), "sigmag": (1800., 10800., 4), "tauf": (1800., 14400., 5), "amp": (1.,)}
    B = bf.Bayes(lc, bf.Transit(ts, amp=1, paramranges=ranges))
    B.bayes_factors_marg_poly_bgd(bglen=55, bgorder=4, noiseestmethod="tailveto", tvsigma=1.0, halfrange=True)
    from copy import deepcopy
    tmp = deepcopy(lc)
    tmp.detrend(method="savitzkygolay", nbins=55, order=4)
    sk = bf.estimate_noise_tv(tmp.clc, sigma=1.0)[0]
    M = B.marginalise_full()
    save("transit5", cts=ts, clc=clc, ref_sk=np.array(sk), ref_lnB=B.lnBmargAmp, ref_marg=M.lnBmargAmp,
         sigmag=np.array(ranges["sigmag"]), tauf=np.array(ranges["tauf"]))


def gold_whole(bf):
    """SURVEY row f3, the executable half: bglen=None, nsinusoids=0 -- ONE polynomial background over the whole
    series (a single global Gram matrix), the template slid along it (bayes.py:247-437 with the `else` branches
    at :292-294, 330-331, 394-395, 408-409).  The reference only gets there with an already detrended light curve
    (otherwise it asks savitzky_golay for a window of None samples and raises), so `detrended` is set by hand."""
    out = {}
    import importlib
    rb = importlib.import_module("bayesflare.stats.bayes")
    keep = (rb.estimate_noise_ps, rb.estimate_noise_tv)
    try:
        for tag, N, bgorder in (("odd", 301, 4), ("even", 300, 2)):
            rng = np.random.RandomState(1300 + N)
            ts = np.arange(N) * DT_LONG
            clc = rng.randn(N) + ref_flare(bf, ts, ts[170], 7., 1200., 3000.) + 0.8 * (ts / ts[-1]) ** 2
            cle = 0.2 + 0.1 * rng.rand(N)
            sk = 0.93
            _force_sk_detector(sk)
            lc = ref_loader.RefLightcurve(ts, clc, cle)
            lc.detrended = True
            out.update({tag + "_cts": ts, tag + "_clc": clc, tag + "_cle": cle, tag + "_sk": np.array(sk),
                        tag + "_bgorder": np.array(bgorder)})
            for name, mk, hr in (("flare", lambda: bf.Flare(ts, amp=1, paramranges={"t0": (np.inf,), "taugauss": (0, 5400, 3),
                                                                                  "tauexp": (1800, 10800, 3), "amp": (1.,)}), True),
                                 ("expdecay", lambda: bf.Expdecay(ts, amp=1, paramranges={"t0": (np.inf,), "tauexp": (0., 900., 3),
                                                                                        "amp": (1.,)}), True),
                                 ("step", lambda: bf.Step(ts, amp=1, paramranges={"t0": (np.inf,), "amp": (1.,)}), False)):
                B = bf.Bayes(lc, mk())
                B.bayes_factors_marg_poly_bgd(bglen=None, bgorder=bgorder, halfrange=hr)
                out["ref_%s_lnB_%s" % (tag, name)] = B.lnBmargAmp
                out["ref_%s_marg_%s" % (tag, name)] = B.marginalise_full().lnBmargAmp
            out["ref_%s_bgonly" % tag] = B.bayes_factors_marg_poly_bgd_only(bglen=None, bgorder=bgorder)
    finally:
        rb.estimate_noise_ps, rb.estimate_noise_tv = keep
    save("whole", **out)


def gold_ingest(bf):
    """Ingest conditioning + Kepler-search detector on a file-like light curve (float32 flux with NaN
    gaps, per-sample float32 errors): data.py:240-277, 279-351; scripts/kepler_analysis_script.py:376-382.
    The FITS container itself is not exercised by the reference here (no pyfits/astropy in the image):
    the arrays a file would deliver are appended exactly as ``add_data`` appends them."""
    rng = np.random.RandomState(77)
    N = 700
    time_days = 131.5 + np.arange(N) * (DT_LONG / 86400.)
    ts = time_days * 24 * 3600
    flux = 12000. + 8. * rng.randn(N) + 30. * np.sin(2 * np.pi * np.arange(N) / 310.)
    flux += ref_flare(bf, ts, ts[400], 160., 900., 4200.) + ref_flare(bf, ts, ts[150], 110., 1500., 6000.)
    err = 8. + 0.6 * np.sin(np.arange(N) / 40.) + 0.3 * rng.rand(N)
    flux32, err32 = flux.astype(np.float32), err.astype(np.float32)
    for a, b in ((60, 61), (233, 235), (520, 523)):
        flux32[a:b] = np.nan
        err32[a:b] = np.nan
    lc = bf.Lightcurve()
    lc.clc = np.append(np.array([]), flux32.copy())
    lc.cts = np.append(np.array([]), time_days * 24 * 3600)
    lc.cle = np.append(np.array([]), err32.copy())
    out = {"time_days": time_days, "flux32": flux32, "err32": err32}
    out["ref_datagap"] = np.array(lc.gap_checker(lc.clc, maxgap=2))
    lc.interpolate()
    lc.dcoffset()
    out["ref_clc"], out["ref_cts"], out["ref_cle"], out["ref_dc"] = lc.clc.copy(), lc.cts.copy(), lc.cle.copy(), np.array(lc.dc)
    odds = bf.OddsRatioDetector(lc, noiseestmethod='tailveto', tvsigma=1.0)
    odds.set_noise_impulse(noiseimpulseparams={'t0': (0, (odds.bglen - 1.) * lc.dt(), odds.bglen)})
    lnO, tso = odds.oddsratio()
    out["ref_lnO"] = np.array(lnO)
    segs, n, mx = odds.thresholder(lnO, 16.5, expand=1)
    out["ref_segs"] = np.array(segs, dtype=np.int64).reshape(-1, 2)
    out["ref_max"] = np.array(mx, dtype=float).reshape(-1, 2)
    import copy
    tmp = copy.copy(lc)
    tmp.detrend(method='savitzkygolay', nbins=odds.bglen, order=odds.bgorder)
    out["ref_sk"] = np.array(bf.estimate_noise_tv(tmp.clc, sigma=odds.tvsigma)[0])
    # the running-median detrender (noise.py:291-317) on the conditioned curve
    out["ref_running_median_21"] = bf.running_median(out["ref_clc"], 21)
    save("ingest", **out)
    print("   ingest: %d segments, nan lnO %d, max %.3f" % (n, int(np.isnan(out["ref_lnO"]).sum()), np.nanmax(out["ref_lnO"])))


def main():
    bf = ref_loader.load_reference()
    which = sys.argv[1:] or ["scalars", "elimination", "kat200", "detector", "config1", "shortcad", "pegrid", "ingest",
                             "trend", "transit5", "whole"]
    if "whole" in which:
        gold_whole(bf)
    if "trend" in which:
        gold_trend(bf)
    if "transit5" in which:
        gold_transit5(bf)
    if "ingest" in which:
        gold_ingest(bf)
    if "scalars" in which:
        gold_scalars(bf)
    if "elimination" in which:
        gold_elimination()
    if "kat200" in which:
        gold_kat200(bf)
    if "detector" in which:
        gold_detector(bf)
    if "pegrid" in which:
        gold_pe(bf)
    if "shortcad" in which:
        gold_short_cadence(bf)
    if "config1" in which:
        gold_config1(bf)


if __name__ == "__main__":
    main()
