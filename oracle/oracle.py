"""CPU oracle for the BayesFlare sliding-window odds-ratio path (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product package ``bayesflare_b200`` never imports it and has no CPU
fallback.

Heavy arithmetic lives in ``bfl_oracle.c`` (plain C, ``-ffp-contract=off``); this module
restates the reference's *host-side* algorithm in NumPy -- templates, priors, window
time-stamps, noise estimators, grid marginalisation, hypothesis combination, thresholding
-- each function citing the reference lines it follows (paths relative to
``/root/reference/bayesflare``).

Parity status: PINNED against the reference itself (see header of ``bfl_oracle.c``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from math import factorial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_SO = os.path.join(HERE, "_build", "libbfl_oracle.so")

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "bfl_oracle.c")
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=gnu99",
               src, "-o", _SO, "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stderr)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_log_erfc.restype = C.c_double
        L.orc_log_erfc.argtypes = [C.c_double]
        L.orc_logplus.restype = C.c_double
        L.orc_logplus.argtypes = [C.c_double, C.c_double]
        L.orc_logtrapz.restype = C.c_double
        L.orc_logtrapz.argtypes = [_dp, C.c_long, _dp, C.c_int]
        L.orc_pairwise_sum.restype = C.c_double
        L.orc_pairwise_sum.argtypes = [_dp, C.c_long]
        L.orc_marg_full.restype = C.c_double
        L.orc_marg_full.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_int]
        L.orc_marg_except_final.restype = C.c_double
        L.orc_marg_except_final.argtypes = [C.c_int, _dp, _dp, C.c_double]
        L.orc_bayes_factors.restype = C.c_int
        L.orc_bayes_factors.argtypes = [_dp, _dp, C.c_long, C.c_int, C.c_int, _dp, _dp,
                                        C.POINTER(C.c_ubyte), C.c_int, C.c_int, _dp, _dp]
        L.orc_pe_grid.restype = C.c_int
        L.orc_pe_grid.argtypes = [_dp, C.c_long, C.c_int, _dp, _dp, C.c_long, C.c_double, _dp]
        L.orc_combine.restype = None
        L.orc_combine.argtypes = [_dp, _dp, C.c_int, C.c_long, C.c_long, C.c_long, _dp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------------------
# scalar helpers
def log_erfc(x):
    return lib().orc_log_erfc(float(x))


def logplus(x, y):
    """stats/general.pyx:519-547"""
    return lib().orc_logplus(float(x), float(y))


def logtrapz(lx, t):
    """stats/general.pyx:576-607"""
    lx, t = _f64(lx), _f64(t)
    return lib().orc_logtrapz(_p(lx), 1, _p(t), len(lx))


def pairwise_sum(a):
    a = _f64(a)
    return lib().orc_pairwise_sum(_p(a), len(a))


def marg_full(M, D, sigma=1.0, lhr=0):
    """stats/log_marg_amp_full.c:3-68; M is (Nm,Nm) with the upper triangle filled."""
    M, D = _f64(M), _f64(D)
    return lib().orc_marg_full(len(D), _p(M), _p(D), float(sigma), int(lhr))


def marg_except_final(M, D, sigma=1.0):
    """stats/log_marg_amp_full.c:71-134"""
    M, D = _f64(M), _f64(D)
    return lib().orc_marg_except_final(len(D), _p(M), _p(D), float(sigma))


# ----------------------------------------------------------------------------------------
# templates: models/model.py
def centre_time(ts):
    """``t0 == inf`` convention, models/model.py:222-223"""
    return ts[int(len(ts) / 2.)]


def flare_model(ts, t0, amp, taugauss, tauexp, reverse=False):
    """Flare.model, models/model.py:197-252"""
    ts = np.asarray(ts, dtype=float)
    if t0 == np.inf:
        t0 = centre_time(ts)
    f = np.zeros(len(ts))
    f[ts == t0] = amp
    if taugauss > 0:
        sel = ts > t0 if reverse else ts < t0
        f[sel] = amp * np.exp(-(ts[sel] - t0) ** 2 / (2 * float(taugauss) ** 2))
    if tauexp > 0:
        if reverse:
            sel = ts < t0
            f[sel] = amp * np.exp((ts[sel] - t0) / float(tauexp))
        else:
            sel = ts > t0
            f[sel] = amp * np.exp(-(ts[sel] - t0) / float(tauexp))
    return f


def expdecay_model(ts, t0, amp, tauexp, reverse=False):
    """Expdecay.model, models/model.py:560-607"""
    ts = np.asarray(ts, dtype=float)
    if t0 == np.inf:
        t0 = centre_time(ts)
    f = np.zeros(len(ts))
    f[ts == t0] = amp
    if tauexp > 0:
        if reverse:
            sel = ts < t0
            f[sel] = amp * np.exp((ts[sel] - t0) / float(tauexp))
        else:
            sel = ts > t0
            f[sel] = amp * np.exp(-(ts[sel] - t0) / float(tauexp))
    return f


def impulse_model(ts, t0, amp):
    """Impulse.model, models/model.py:694-736 (finite t0 is an offset from ts[0])"""
    ts = np.asarray(ts, dtype=float)
    t0 = centre_time(ts) if t0 == np.inf else ts[0] + t0
    f = np.zeros_like(ts)
    f[np.abs(ts - t0).argmin()] = amp
    return f


def step_model(ts, t0, amp):
    """Step.model, models/model.py:943-982"""
    ts = np.asarray(ts, dtype=float)
    if t0 == np.inf:
        t0 = centre_time(ts)
    f = np.zeros_like(ts)
    f[ts > t0] = amp
    return f


def gaussian_model(ts, t0, amp, sigma):
    """Gaussian.model, models/model.py:814-858"""
    ts = np.asarray(ts, dtype=float)
    if t0 == np.inf:
        t0 = centre_time(ts)
    if sigma == 0:
        f = np.zeros(len(ts))
        tm0 = ts - t0
        f[np.amin(tm0) == tm0] = amp
        return f
    return amp * np.exp(-(ts - t0) ** 2 / (2 * float(sigma) ** 2))


def transit_model(ts, t0, amp, sigmag, tauf):
    """Transit.model, models/model.py:387-436"""
    ts = np.asarray(ts, dtype=float)
    if t0 == np.inf:
        t0 = centre_time(ts)
    f = -1. * amp * np.ones(len(ts))
    lo, hi = ts < t0 - tauf, ts > t0 + tauf
    if sigmag > 0:
        f[lo] = -1 * amp * np.exp(-(ts[lo] - (t0 - tauf)) ** 2 / (2 * float(sigmag) ** 2))
        f[hi] = -1 * amp * np.exp(-(ts[hi] - (t0 + tauf)) ** 2 / (2 * float(sigmag) ** 2))
    else:
        f[lo] = 0
        f[hi] = 0
    return f


MODEL_PARAMNAMES = {
    "flare": ["t0", "tauexp", "taugauss", "amp"],       # model.py:192
    "transit": ["t0", "sigmag", "tauf", "amp"],         # model.py:368
    "expdecay": ["t0", "amp", "tauexp"],                # model.py:555
    "impulse": ["t0", "amp"],                           # model.py:689
    "gaussian": ["t0", "sigma", "amp"],                 # model.py:809
    "step": ["t0", "amp"],                              # model.py:938
}


def make_ranges(mtype, paramranges):
    """Model.set_params, models/model.py:69-92"""
    ranges, shape = {}, []
    for p in MODEL_PARAMNAMES[mtype]:
        r = paramranges[p]
        if len(r) == 1:
            ranges[p] = np.array([r[0]])
        elif len(r) == 3:
            ranges[p] = np.linspace(r[0], r[1], r[2])
        else:
            raise ValueError("Error... range must either contain 1 or 3 values")
        shape.append(len(ranges[p]))
    return ranges, shape


def eval_model(mtype, pdict, ts, reverse=False):
    if mtype == "flare":
        return flare_model(ts, pdict["t0"], pdict["amp"], pdict["taugauss"], pdict["tauexp"], reverse)
    if mtype == "expdecay":
        return expdecay_model(ts, pdict["t0"], pdict["amp"], pdict["tauexp"], reverse)
    if mtype == "impulse":
        return impulse_model(ts, pdict["t0"], pdict["amp"])
    if mtype == "step":
        return step_model(ts, pdict["t0"], pdict["amp"])
    if mtype == "gaussian":
        return gaussian_model(ts, pdict["t0"], pdict["amp"], pdict["sigma"])
    if mtype == "transit":
        return transit_model(ts, pdict["t0"], pdict["amp"], pdict["sigmag"], pdict["tauf"])
    raise ValueError(mtype)


def _flat(rng):
    return -np.log(rng[-1] - rng[0]) if len(rng) > 1 else 0.


def model_prior(mtype, pdict, ranges, maxdur=np.inf, sigmacutoff=3.):
    """``prior`` methods, models/model.py:254-329 (Flare), :438-517 (Transit, with the two
    undefined names fixed as in ref_loader.PATCHES), :609-652, :738-773, :984-1019."""
    if mtype == "flare":
        tg, te = pdict["taugauss"], pdict["tauexp"]
        tgr, ter = ranges["taugauss"], ranges["tauexp"]
        base = _flat(ranges["amp"]) + _flat(ranges["t0"])
        if tg > te or tg > ter[-1]:
            return base + (-np.inf)
        tgmin, tgmax, temin, temax = tgr[0], tgr[-1], ter[0], ter[-1]
        dtg, dte = tgmax - tgmin, temax - temin
        parea = None
        if tgmin <= temin and tgmax <= temax:
            parea = dte * dtg - 0.5 * (tgmax - temin) ** 2
        elif tgmin > temin and tgmax > temax:
            parea = 0.5 * (temax - tgmin) ** 2
        elif tgmin > temin and tgmax < temax:
            parea = 0.5 * dtg * ((temax - tgmin) + (temax - tgmax))
        elif tgmin < temin and tgmax > temax:
            parea = 0.5 * dte * ((temin - tgmin) + (temax - tgmin))
        return base + (-np.log(parea))
    if mtype == "transit":
        sgr, tfr = ranges["sigmag"], ranges["tauf"]
        if pdict["tauf"] + 2. * sigmacutoff * pdict["sigmag"] > maxdur:
            return -np.inf
        t0p, ampp = _flat(ranges["t0"]), _flat(ranges["amp"])
        maxtf, mintf, maxsg, minsg = tfr[-1], tfr[0], sgr[-1], sgr[0]
        sg1 = (maxdur - mintf) / (2. * sigmacutoff)
        sg2 = (maxdur - maxtf) / (2. * sigmacutoff)
        tf1 = maxdur - (2. * sigmacutoff) * minsg
        tf2 = maxdur - (2. * sigmacutoff) * maxsg
        dsg, dtf = maxsg - minsg, maxtf - mintf
        lntf = 0.
        lnsg = None
        if sg1 < minsg and (sg2 > minsg and sg2 <= maxsg):
            lnsg = -np.log((sg2 - minsg) * (tf1 - mintf) / 2.)
        elif (sg1 > minsg and sg1 <= maxsg) and (sg2 > minsg and sg2 <= maxsg):
            lnsg = -np.log((dtf / 2.) * (sg2 + sg1 - 2. * minsg))
        elif (tf1 > mintf and tf1 < maxtf) and (tf2 > mintf and tf2 < maxtf):
            lnsg = -np.log((dsg / 2.) * (tf2 + tf1 - 2. * mintf))
        elif (sg1 > minsg and sg1 <= maxsg) and sg2 > maxsg:
            lnsg = -np.log(dtf * (sg1 - minsg) + (maxsg - sg1) * (dtf + (tf2 - mintf)) / 2.)
        else:
            if len(sgr) > 1:
                lnsg = -np.log(sgr[-1] - sgr[0])
            if len(tfr) > 1:
                lntf = -np.log(tfr[-1] - tfr[0])
        return ampp + t0p + lntf + lnsg
    if mtype == "expdecay":
        return _flat(ranges["t0"]) + _flat(ranges["amp"]) + _flat(ranges["tauexp"])
    if mtype in ("impulse", "step"):
        return _flat(ranges["t0"]) + _flat(ranges["amp"])
    raise ValueError(mtype)


# ----------------------------------------------------------------------------------------
# noise estimation: noise/noise.py, data/data.py
def savitzky_golay(y, window_size, order, deriv=0, rate=1):
    """noise/noise.py:176-252"""
    window_size, order = abs(int(window_size)), abs(int(order))
    if window_size % 2 != 1 or window_size < 1:
        raise TypeError("window_size size must be a positive odd number")
    if window_size < order + 2:
        raise TypeError("window_size is too small for the polynomials order")
    half = (window_size - 1) // 2
    b = np.array([[k ** i for i in range(order + 1)] for k in range(-half, half + 1)], dtype=float)
    m = np.linalg.pinv(b)[deriv] * rate ** deriv * factorial(deriv)
    y = np.asarray(y, dtype=float)
    firstvals = y[0] - np.abs(y[1:half + 1][::-1] - y[0])
    lastvals = y[-1] + np.abs(y[-half - 1:-1][::-1] - y[-1])
    y = np.concatenate((firstvals, y, lastvals))
    return np.convolve(m[::-1], y, mode="valid")


def psd_onesided(x, fs):
    """Lightcurve.psd, data/data.py:184-203 -> matplotlib.mlab.psd with a boxcar window,
    NFFT=len(x), no overlap (modern-mlab scaling; see oracle/ref_loader.mlab_psd)."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    spec = np.abs(np.fft.rfft(x, n=n)) ** 2 / n
    if n % 2:
        spec[1:] *= 2.0
    else:
        spec[1:-1] *= 2.0
    return spec / fs, np.fft.rfftfreq(n, 1.0 / fs)


def estimate_noise_ps(clc, cts, estfrac=0.5):
    """noise/noise.py:10-44"""
    fs = 1.0 / (cts[1] - cts[0])
    sk, _ = psd_onesided(clc, fs)
    sk = np.mean(sk[int(np.floor((1. - estfrac) * len(sk))):])
    sk = sk * fs / 2.
    return np.sqrt(sk), sk


def estimate_noise_tv(d, sigma=1.0):
    """noise/noise.py:47-103"""
    from scipy.interpolate import interp1d
    from scipy.special import erf
    d = np.asarray(d, dtype=float)
    n, bins = np.histogram(d, bins=len(d), density=True)
    centres = (bins[:-1] + bins[1:]) / 2.
    cs = np.cumsum(n * (bins[1] - bins[0]))
    csu, idx = np.unique(cs, return_index=True)
    binsu = centres[idx]
    cp = erf(sigma / np.sqrt(2.))
    f = interp1d(csu, binsu)
    lowS, highS, m = f(0.5 - cp / 2.), f(0.5 + cp / 2.), f(0.5)
    return (highS - lowS) / (2. * sigma), m


def noise_sigma(clc, cts, bglen, bgorder, noiseestmethod="powerspectrum", psestfrac=0.5, tvsigma=1.0,
                detrended=False):
    """stats/bayes.py:235-246: estimate on a Savitzky-Golay detrended copy."""
    clc = np.asarray(clc, dtype=float)
    if not detrended:
        clc = clc - savitzky_golay(clc, bglen, bgorder)
    if noiseestmethod == "powerspectrum":
        return float(estimate_noise_ps(clc, cts, psestfrac)[0])
    if noiseestmethod == "tailveto":
        return float(estimate_noise_tv(clc, tvsigma)[0])
    raise ValueError("Noise estimation method must be 'powerspectrum' or 'tailveto'")


# ----------------------------------------------------------------------------------------
# Bayes.bayes_factors_marg_poly_bgd / _only
def window_times(ts, bglen, model_t0=None):
    """stats/bayes.py:261-276: the W time stamps the templates are evaluated on."""
    ts = np.asarray(ts, dtype=float)
    N = len(ts)
    nsteps = int(bglen / 2)
    if model_t0 is None:
        model_t0 = ts[int(np.floor(N / 2))]          # Model.__init__, model.py:48-49
    dt = ts[1] - ts[0]
    idxt0 = int((model_t0 - ts[0]) / dt) + 1
    idx1, idx2 = idxt0 - nsteps, idxt0 + nsteps + 1
    if idx1 < 0:
        return ts[:bglen]
    if idx2 > N - 1:
        return ts[-bglen:]
    return ts[idx1:idx2]


def background_models(bglen, bgorder):
    """stats/bayes.py:292-307"""
    tsp = np.linspace(0., 1., bglen)
    return np.array([tsp ** i for i in range(bgorder + 1)])


def bayes_factors(clc, cle, cts, mtype, paramranges, sk, bglen=55, bgorder=4, halfrange=True,
                  reverse=False, with_background=False, model_t0=None, maxdur=np.inf):
    """``Bayes.bayes_factors_marg_poly_bgd`` (stats/bayes.py:153-437) for one hypothesis with
    the noise sigma ``sk`` given.  Returns ``lnBmargAmp`` of shape ``model.shape + (N,)``,
    the ranges dict, and (optionally) the background-only factor of
    ``bayes_factors_marg_poly_bgd_only`` (stats/bayes.py:439-563)."""
    clc, cle, cts = _f64(clc), _f64(cle), _f64(cts)
    N = len(cts)
    ranges, shape = make_ranges(mtype, paramranges)
    names = MODEL_PARAMNAMES[mtype]
    G = int(np.prod(shape))
    mts = window_times(cts, bglen, model_t0)
    bg = _f64(background_models(bglen, bgorder))
    v = _f64(sk ** 2 + cle ** 2)
    ms = np.zeros((G, bglen))
    valid = np.zeros(G, dtype=np.uint8)
    priors = np.zeros(G)
    for g in range(G):
        q = np.unravel_index(g, shape)
        pd = {names[k]: ranges[names[k]][q[k]] for k in range(len(names))}
        priors[g] = model_prior(mtype, pd, ranges, maxdur=maxdur)
        if priors[g] != -np.inf:
            ms[g] = eval_model(mtype, pd, mts, reverse)
            valid[g] = 1
    lnB = np.empty((G, N))
    lnBbg = np.empty(N) if with_background else None
    rc = lib().orc_bayes_factors(_p(clc), _p(v), N, bglen, bgorder + 1, _p(bg), _p(ms),
                                 valid.ctypes.data_as(C.POINTER(C.c_ubyte)), G, 1 if halfrange else 0,
                                 _p(lnB), _p(lnBbg) if with_background else None)
    if rc != 0:
        raise RuntimeError("orc_bayes_factors failed: %d" % rc)
    # general.pyx:302-304: the 2-parameter variant blanks the *whole* row if any mdcross
    # is infinite; with finite inputs this only happens for invalid rows (already -inf).
    ampprior = np.log(0.5) if halfrange else 0.          # bayes.py:284-288
    lnB = lnB + priors[:, None] + ampprior                # bayes.py:431-435
    lnB = lnB.reshape(tuple(shape) + (N,))
    if with_background:
        return lnB, ranges, lnBbg
    return lnB, ranges


def background_only(clc, cle, cts, sk, bglen=55, bgorder=4):
    """``Bayes.bayes_factors_marg_poly_bgd_only`` (stats/bayes.py:439-563)"""
    clc, cle, cts = _f64(clc), _f64(cle), _f64(cts)
    N = len(cts)
    bg = _f64(background_models(bglen, bgorder))
    v = _f64(sk ** 2 + cle ** 2)
    out = np.empty(N)
    rc = lib().orc_bayes_factors(_p(clc), _p(v), N, bglen, bgorder + 1, _p(bg), None, None, 0, 0,
                                 None, _p(out))
    if rc != 0:
        raise RuntimeError("orc_bayes_factors failed: %d" % rc)
    return out


def _marg_full(M, D, lhr):
    M, D = _f64(M), _f64(D)
    return lib().orc_marg_full(len(D), _p(M), _p(D), 1.0, 1 if lhr else 0)


def bayes_factors_whole_series(clc, cle, cts, mtype, paramranges, sk, bgorder=4, halfrange=True, reverse=False,
                               with_background=False):
    """``Bayes.bayes_factors_marg_poly_bgd(bglen=None, nsinusoids=0)`` (stats/bayes.py:247-437 with the `else`
    branches at :292-294, 330-331, 394-395, 408-409): ONE polynomial background over the whole series -- global
    Gram matrix and data projections, constant along the series -- and the full-length template slid along it
    with the reference's own truncation slices (:367-377; for an even number of samples the end branch places
    the template one sample late, kept).  np.correlate / np.sum as the reference calls them; the per-sample
    elimination is the C restatement (log_marg_amp_full.c:3-68)."""
    clc, cle, cts = _f64(clc), _f64(cle), _f64(cts)
    N = len(cts)
    nsteps = int(N / 2)
    P = bgorder + 1
    Nm = P + 1
    ranges, shape = make_ranges(mtype, paramranges)
    names = MODEL_PARAMNAMES[mtype]
    G = int(np.prod(shape))
    L = lib()
    L.orc_marg_full.restype = C.c_double
    L.orc_marg_full.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_int]
    tsp = np.linspace(0., 1., N)
    bg = np.array([tsp ** i for i in range(P)])
    v = sk ** 2 + cle ** 2
    bgcross = np.zeros((P, P))
    for i in range(P):
        for j in range(i, P):
            bgcross[i, j] = np.sum(bg[i] * bg[j] / v)
    d = clc / v
    dbgr = np.array([np.sum(d * bg[i]) for i in range(P)])
    sumlogsigma = np.sum(np.log(np.sqrt(v)[np.isfinite(np.sqrt(v))]))
    lnB = np.full((G, N), -np.inf)
    ampprior = np.log(0.5) if halfrange else 0.
    for g in range(G):
        q = np.unravel_index(g, shape)
        pd = {names[k]: ranges[names[k]][q[k]] for k in range(len(names))}
        prior = model_prior(mtype, pd, ranges)
        if prior == -np.inf:
            continue
        m = eval_model(mtype, pd, cts, reverse)
        mdcross = np.empty(N)
        for k in range(N):
            if k < nsteps:
                mm = m[nsteps - k:] ** 2
                mm = mm / v[:len(mm)]
            elif k >= N - nsteps:
                mm = m[:(N - nsteps - k - 1)] ** 2
                mm = mm / v[-len(mm):] if len(mm) else mm
            else:
                mm = m ** 2 / v[k - nsteps:k + nsteps + 1]
            mdcross[k] = np.sum(mm)
        mdbg = np.array([np.correlate(bg[j] / v, m, 'same') for j in range(P)])
        dm = np.correlate(d, m, 'same')
        M = np.zeros((Nm, Nm))
        M[:P, :P] = bgcross
        D = np.zeros(Nm)
        D[:P] = dbgr
        for k in range(N):
            M[:P, P] = mdbg[:, k]
            M[P, P] = mdcross[k]
            D[P] = dm[k]
            lnB[g, k] = _marg_full(M, D, halfrange) + sumlogsigma
        lnB[g] += prior + ampprior
    lnB = lnB.reshape(tuple(shape) + (N,))
    if with_background:
        return lnB, ranges, np.full(N, _marg_full(bgcross, dbgr, False) + sumlogsigma)
    return lnB, ranges


def marginalise_full(lnB, ranges, names):
    """``Bayes.marginalise_full`` (stats/bayes.py:565-621): for every parameter in order,
    drop a singleton axis or apply ``logtrapz`` along it."""
    arr = np.asarray(lnB)
    names = list(names)
    for p in list(names):
        axis = names.index(p)
        places = ranges[p]
        if len(places) > 1:
            arr = np.apply_along_axis(lambda col: logtrapz(col, places), axis, arr)
        else:
            arr = arr.reshape(tuple(s for i, s in enumerate(arr.shape) if i != axis))
        names.remove(p)
    return arr


DEFAULT_FLAREPARAMS = {"taugauss": (0, 1.5 * 60 * 60, 10), "tauexp": (0.5 * 60 * 60, 3. * 60 * 60, 10)}
DEFAULT_EXPDECAYPARAMS = {"tauexp": (0.0, 0.25 * 60 * 60, 3)}


def oddsratio(clc, cle, cts, bglen=55, bgorder=4, noiseestmethod="powerspectrum", psestfrac=0.5, tvsigma=1.0,
              flareparams=None, noisepoly=True, noiseimpulse=True, noiseimpulseparams=None,
              noiseimpulsepositive=False, noiseexpdecay=True, noiseexpdecayparams=None,
              noiseexpdecaywithreverse=True, noisestep=False, noisestepparams=None, ignoreedges=True,
              sk=None, return_parts=False):
    """``OddsRatioDetector.oddsratio`` (finder/find.py:741-864)."""
    clc, cle, cts = _f64(clc), _f64(cle), _f64(cts)
    N = len(cts)
    if sk is None:
        sk = noise_sigma(clc, cts, bglen, bgorder, noiseestmethod, psestfrac, tvsigma)
    fp = dict(DEFAULT_FLAREPARAMS if flareparams is None else flareparams)
    fp.setdefault("t0", (np.inf,))
    fp["amp"] = (1.,)
    lnBf, rf, Bg = bayes_factors(clc, cle, cts, "flare", fp, sk, bglen, bgorder, True, with_background=True)
    Of = marginalise_full(lnBf, rf, MODEL_PARAMNAMES["flare"])
    parts = {"flare": Of}
    noise = []
    if noisepoly:
        noise.append(Bg)
        parts["poly"] = Bg
    if noiseimpulse:
        ip = dict({"t0": (np.inf,)} if noiseimpulseparams is None else noiseimpulseparams)
        ip["amp"] = (1.,)
        lnBi, ri = bayes_factors(clc, cle, cts, "impulse", ip, sk, bglen, bgorder, noiseimpulsepositive)
        noise.append(marginalise_full(lnBi, ri, MODEL_PARAMNAMES["impulse"]))
        parts["impulse"] = noise[-1]
    if noiseexpdecay:
        ep = dict(DEFAULT_EXPDECAYPARAMS if noiseexpdecayparams is None else noiseexpdecayparams)
        ep.setdefault("t0", (np.inf,))
        ep["amp"] = (1.,)
        lnBe, re_ = bayes_factors(clc, cle, cts, "expdecay", ep, sk, bglen, bgorder, True)
        noise.append(marginalise_full(lnBe, re_, MODEL_PARAMNAMES["expdecay"]))
        parts["expdecay"] = noise[-1]
        if noiseexpdecaywithreverse:
            lnBr, rr = bayes_factors(clc, cle, cts, "expdecay", ep, sk, bglen, bgorder, True, reverse=True)
            noise.append(marginalise_full(lnBr, rr, MODEL_PARAMNAMES["expdecay"]))
            parts["expdecay_rev"] = noise[-1]
    if noisestep:
        sp = dict({"t0": (np.inf,)} if noisestepparams is None else noisestepparams)
        sp["amp"] = (1.,)
        lnBs, rs = bayes_factors(clc, cle, cts, "step", sp, sk, bglen, bgorder, False)
        noise.append(marginalise_full(lnBs, rs, MODEL_PARAMNAMES["step"]))
        parts["step"] = noise[-1]
    if ignoreedges and bglen is not None:
        k0, k1 = int(bglen / 2), N - int(bglen / 2)
    else:
        k0, k1 = 0, N
    lnO = np.empty(k1 - k0)
    nz = _f64(np.array(noise)) if noise else np.zeros((0, N))
    lib().orc_combine(_p(_f64(Of)), _p(nz), len(noise), N, k0, k1, _p(lnO))
    ts = np.copy(cts[k0:k1])
    if return_parts:
        return lnO, ts, parts, sk
    return lnO, ts


# ----------------------------------------------------------------------------------------
# thresholding: finder/find.py:15-52, :915-1019
def contiguous_regions(condition):
    """finder/find.py:15-52"""
    condition = np.asarray(condition, dtype=bool)
    d = np.diff(condition)
    idx, = d.nonzero()
    idx = idx + 1
    if condition[0]:
        idx = np.r_[0, idx]
    if condition[-1]:
        idx = np.r_[idx, condition.size]
    return idx.reshape(-1, 2)


def thresholder(lnO, thresh, expand=0, returnmax=True):
    """``OddsRatioDetector.thresholder`` (finder/find.py:915-1019)"""
    lnOc = np.array(lnO, dtype=float)
    n = len(lnOc)
    segs = [(int(a), int(b)) for a, b in contiguous_regions(lnOc > thresh)]
    if expand > 0 and segs:
        ex = []
        for a, b in segs:
            a, b = a - expand, b + expand
            if a < 0:
                a = 0
            if b >= n:
                b = n
            ex.append((a, b))
        if len(ex) == 1:
            segs = [list(ex[0])]                # the reference returns a list here (find.py:961)
        else:
            merged, j = [], 0
            while True:
                this = ex[j]
                j += 1
                for k in range(j, len(ex)):
                    nxt = ex[k]
                    if this[-1] >= nxt[0]:
                        this = (this[0], nxt[-1])
                        j += 1
                    else:
                        break
                merged.append(this)
                if j >= len(ex):
                    break
            segs = merged
    if not returnmax:
        return segs, len(segs)
    maxlist = []
    for a, b in segs:
        i = int(np.argmax(lnOc[a:b])) + a
        maxlist.append((lnOc[i], i))
    return segs, len(segs), maxlist


# ----------------------------------------------------------------------------------------
# ParameterEstimationGrid.calculate_posterior: stats/bayes.py:912-1055
def pe_posterior(clc, cts, mtype, grid_values, sigma, bgorder=4, margbackground=True):
    """grid_values: dict name -> 1-D array (already built as ``set_grid`` does,
    stats/bayes.py:771-816).  Returns the log posterior on the grid."""
    clc, cts = _f64(clc), _f64(cts)
    names = MODEL_PARAMNAMES[mtype]
    sp = [len(grid_values[p]) for p in names]
    G = int(np.prod(sp))
    L = len(cts)
    ranges = {p: np.asarray(grid_values[p]) for p in names}
    ms = np.empty((G, L))
    prior = np.empty(G)
    for g in range(G):
        q = np.unravel_index(g, sp)
        pd = {names[i]: grid_values[names[i]][q[i]] for i in range(len(names))}
        ms[g] = eval_model(mtype, pd, cts)
        prior[g] = model_prior(mtype, pd, ranges)
    if margbackground:
        t = np.linspace(0, 1, L)
        polyms = _f64(np.array([t ** i for i in range(bgorder + 1)]))
        post = np.empty(G)
        rc = lib().orc_pe_grid(_p(clc), L, bgorder + 1, _p(polyms), _p(_f64(ms)), G, float(sigma), _p(post))
        if rc != 0:
            raise RuntimeError("orc_pe_grid failed")
    else:
        # general.pyx:659-691
        post = np.array([-np.sum(m ** 2 - 2. * clc * m) / (2. * sigma ** 2) for m in ms])
    return (post + prior).reshape(sp)
