"""ctypes binding of ``libbfl.so`` (C ABI declared in ``include/bfl.h``).

There is no CPU fallback: if the library is missing, or no CUDA device is present when a
compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFL_LIB") or os.path.join(HERE, "libbfl.so")

MODEL_KINDS = {"flare": 0, "expdecay": 1, "impulse": 2, "step": 3, "transit": 4, "gaussian": 5, "table": 6}

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


class BflError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbfl error %d: %s" % (code, msg))
        self.code = code


class Hypothesis(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reverse", C.c_int32), ("halfrange", C.c_int32), ("ngrid", C.c_int32),
                ("params", _dp), ("logprior", _dp), ("logweight", _dp), ("table", _dp)]


class Region(C.Structure):
    _fields_ = [("curve", C.c_int32), ("start", C.c_int32), ("stop", C.c_int32), ("argmax", C.c_int32),
                ("maxval", C.c_double)]


REGION_DTYPE = np.dtype([("curve", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("argmax", "<i4"), ("maxval", "<f8")])


class NoiseSpec(C.Structure):
    _fields_ = [("method", C.c_int32), ("bglen", C.c_int32), ("param", C.c_double), ("sgcoef", _dp)]


NOISE_METHODS = {"powerspectrum": 0, "tailveto": 1}


def make_noise_spec(method, bglen, bgorder, psestfrac=0.5, tvsigma=1.0, detrended=False):
    """``bfl_noise_spec`` for the reference's noise estimate (bayes.py:235-246).  Returns the
    structure and the arrays that must stay alive while it is in use."""
    from .noise import _sg_coefficients
    if method not in NOISE_METHODS:
        raise ValueError("Noise estimation method must be 'powerspectrum' or 'tailveto'")
    spec = NoiseSpec()
    spec.method = NOISE_METHODS[method]
    coef = None
    if detrended:
        spec.bglen = 0
    else:
        coef = f64(_sg_coefficients(int(bglen), int(bgorder)))
        spec.bglen = int(bglen)
        spec.sgcoef = ptr(coef)
    spec.param = float(psestfrac if method == "powerspectrum" else tvsigma)
    return spec, coef


# every symbol include/bfl.h declares: (restype, argtypes)
SIGNATURES = {
    "bfl_version": (C.c_int, []),
    "bfl_last_error": (C.c_char_p, []),
    "bfl_device_count": (C.c_int, []),
    "bfl_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bfl_ctx_destroy": (C.c_int, [C.c_void_p]),
    "bfl_ctx_sync": (C.c_int, [C.c_void_p]),
    "bfl_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "bfl_ctx_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "bfl_ctx_window_time": (C.c_int, [C.c_void_p, _dp, _i64p]),
    "bfl_count_regions_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p]),
    "bfl_count_regions_total_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p,
                                              C.c_void_p]),
    "bfl_sweep_simulate_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "bfl_sweep_inject_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "bfl_sweep_score_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_double,
                                      C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bfl_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.POINTER(Hypothesis), C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "bfl_plan_destroy": (C.c_int, [C.c_void_p]),
    "bfl_plan_templates": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "bfl_plan_shape_count": (C.c_int, [C.c_void_p]),
    "bfl_window_lnB": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _dp, _dp, C.c_double, C.c_int64, C.c_double,
                                 _dp, _dp]),
    "bfl_noise_sigma": (C.c_int, [C.c_void_p, C.POINTER(NoiseSpec), _dp, C.c_int32, C.c_int64, _dp]),
    "bfl_noise_sigma_dev": (C.c_int, [C.c_void_p, C.POINTER(NoiseSpec), C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "bfl_detector_odds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _dp, _dp, _dp, C.POINTER(NoiseSpec), C.c_int32,
                                    C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp]),
    "bfl_detector_odds_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_void_p]),
    "bfl_batch_regions_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p, C.c_int32,
                                        C.c_void_p]),
    "bfl_pipe_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(NoiseSpec), C.c_int32, C.c_int64, C.c_int64,
                                  C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "bfl_pipe_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p, _i64p]),
    "bfl_pipe_wait": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]),
    "bfl_pipe_launch_count": (C.c_int64, [C.c_void_p]),
    "bfl_pipe_graph_count": (C.c_int32, [C.c_void_p]),
    "bfl_pipe_destroy": (C.c_int, [C.c_void_p]),
    "bfl_threshold": (C.c_int, [C.c_void_p, _dp, C.c_int64, C.c_double, C.c_int32, _i64p, _dp, _i64p, C.c_int32,
                                C.POINTER(C.c_int32)]),
    "bfl_pe_grid": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _dp, _dp, C.c_int64, _dp, _dp, C.c_int64, C.c_int32,
                              _dp, C.c_double, C.c_int32, _dp, C.c_int32, _dp]),
    "bfl_whole_series_lnB": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, C.c_int64, _dp, _dp, C.c_int64,
                                       C.c_int32, _dp, _dp, _dp, C.c_double, _dp, _dp]),
    "bfl_logtrapz": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64, C.c_int64, _dp]),
    "bfl_fp64_peak_probe": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libbfl.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError("%s is not built (run `python -m bayesflare_b200.build`); "
                                  "bayesflare_b200 has no CPU fallback" % LIB_PATH)
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise BflError(rc, (load().bfl_last_error() or b"").decode("utf-8", "replace"))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Context:
    """One device + one stream + grow-only scratch (``bfl_ctx``)."""

    def __init__(self, device=-1, stream=None):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.bfl_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = int(device)          # -1: the device that was current at creation
        self._plans = weakref.WeakSet()    # plans hold a pointer into this context: closed before it

    def sync(self):
        check(self.lib.bfl_ctx_sync(self.handle))

    @property
    def launches(self):
        return int(self.lib.bfl_ctx_launch_count(self.handle))

    def set_timing(self, enable=True):
        check(self.lib.bfl_ctx_set_timing(self.handle, 1 if enable else 0))

    def window_time(self):
        """(summed ms, launches) of the fused window kernel since the last call."""
        ms, n = C.c_double(0), C.c_int64(0)
        check(self.lib.bfl_ctx_window_time(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def sweep_simulate_dev(self, seed, first, B, N, halfwin, dt, nstd, sinusoid_amp, snr_lo, snr_hi, inject, d_clc, d_truth):
        """Simulate a batch of light curves on the device (device addresses; see ``bfl_sweep_simulate_dev``)."""
        check(self.lib.bfl_sweep_simulate_dev(self.handle, int(seed), int(first), int(B), int(N), int(halfwin), float(dt),
                                              float(nstd), float(sinusoid_amp), float(snr_lo), float(snr_hi),
                                              int(inject), d_clc, d_truth))

    def sweep_inject_dev(self, kind, B, N, dt, d_sk, d_truth, d_clc):
        """Add the recorded injections (kind 1 flare, 2 transit) rescaled with the per-curve sigma ``d_sk``."""
        check(self.lib.bfl_sweep_inject_dev(self.handle, int(kind), int(B), int(N), float(dt), d_sk, d_truth, d_clc))

    def sweep_score_dev(self, d_lnO, B, nk, halfwin, d_truth, thresh, nbins, snr_max, d_injected, d_detected, d_counters):
        """Match detections to injections and accumulate the efficiency histograms (``bfl_sweep_score_dev``)."""
        check(self.lib.bfl_sweep_score_dev(self.handle, d_lnO, int(B), int(nk), int(halfwin), d_truth, float(thresh),
                                           int(nbins), float(snr_max), d_injected, d_detected, d_counters))

    def noise_sigma_dev(self, spec, d_clc, B, N, d_sk):
        """Device-resident noise estimate (``bfl_noise_sigma_dev``)."""
        check(self.lib.bfl_noise_sigma_dev(self.handle, C.byref(spec[0]), d_clc, int(B), int(N), d_sk))

    def count_regions_dev(self, d_lnO, B, n, thresh, d_counts, d_totals=None):
        """Per-curve region counts (device addresses); with ``d_totals`` also accumulates the batch totals
        [regions, curves with a region] into that 2-element int64 device array."""
        if d_totals is None:
            check(self.lib.bfl_count_regions_dev(self.handle, d_lnO, int(B), int(n), float(thresh), d_counts))
        else:
            check(self.lib.bfl_count_regions_total_dev(self.handle, d_lnO, int(B), int(n), float(thresh), d_counts,
                                                       d_totals))

    def batch_regions_dev(self, d_lnO, B, n, thresh, d_regions, max_regions, d_counts):
        """Above-threshold regions of a device-resident batch (device addresses; ``bfl_batch_regions_dev``)."""
        check(self.lib.bfl_batch_regions_dev(self.handle, d_lnO, int(B), int(n), float(thresh), d_regions, int(max_regions),
                                             d_counts))

    def close(self):
        if getattr(self, "handle", None):
            for p in list(getattr(self, "_plans", ())):     # bfl_plan_destroy synchronises the context's stream
                p.close()
            self.lib.bfl_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """Process-wide context for ``device`` (default: ``BFL_DEVICE`` or LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("BFL_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        ndev = load().bfl_device_count()
        if ndev > 0:
            device %= ndev
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx


class Plan:
    """``bfl_plan``: hypotheses (templates + per-grid constants) on one window time base."""

    def __init__(self, ctx, bglen, bgmodels, mts, hyps):
        """hyps: list of dicts(kind, reverse, halfrange, params[G,4], logprior[G], logweight[G], table)."""
        self.ctx = ctx
        self.lib = ctx.lib
        self.bglen = int(bglen)
        bg = f64(bgmodels)
        self.npoly = bg.shape[0]
        keep = [bg, f64(mts)]
        arr = (Hypothesis * max(len(hyps), 1))()
        self.ngrid = []
        for i, h in enumerate(hyps):
            lp, lw = f64(h["logprior"]).ravel(), f64(h["logweight"]).ravel()
            G = lp.size
            par = f64(h["params"]).reshape(G, 4) if h.get("params") is not None else None
            tab = f64(h["table"]).reshape(G, self.bglen) if h.get("table") is not None else None
            keep += [lp, lw, par, tab]
            arr[i].kind = MODEL_KINDS[h["kind"]] if isinstance(h["kind"], str) else int(h["kind"])
            arr[i].reverse = 1 if h.get("reverse") else 0
            arr[i].halfrange = 1 if h.get("halfrange") else 0
            arr[i].ngrid = G
            arr[i].params = ptr(par)
            arr[i].logprior = ptr(lp)
            arr[i].logweight = ptr(lw)
            arr[i].table = ptr(tab)
            self.ngrid.append(G)
        self.nhyp = len(hyps)
        hnd = C.c_void_p()
        check(self.lib.bfl_plan_create(ctx.handle, self.bglen, self.npoly, ptr(bg), ptr(keep[1]), arr, self.nhyp,
                                       C.byref(hnd)))
        self.handle = hnd
        ctx._plans.add(self)

    @property
    def shape_count(self):
        """Correlation shapes per sample in the separable form, or -1 (dense form only)."""
        return int(self.lib.bfl_plan_shape_count(self.handle))

    def templates(self, ihyp):
        out = np.zeros((self.ngrid[ihyp], self.bglen))
        check(self.lib.bfl_plan_templates(self.handle, int(ihyp), ptr(out)))
        return out

    def window_lnB(self, ihyp, clc, cle, sk, sumlogsigma, want_lnB=True, want_bg=False):
        clc = f64(clc)
        cle = None if cle is None else f64(cle)
        N = clc.size
        out = np.empty((self.ngrid[ihyp], N)) if want_lnB else None
        bg = np.empty(N) if want_bg else None
        check(self.lib.bfl_window_lnB(self.ctx.handle, self.handle, int(ihyp) if want_lnB else -1, ptr(clc), ptr(cle),
                                      float(sk), N, float(sumlogsigma), ptr(out), ptr(bg)))
        return out, bg

    def detector_odds(self, clc, cle, sk, noisepoly=True, want_marg=False, k0=0, k1=None, noise=None,
                      want_sk=False):
        """Log odds for samples [k0, k1) of every curve (default: all N samples).  With ``sk=None`` the
        noise sigma of every curve is estimated on the device as ``noise`` (a :func:`make_noise_spec`
        result) describes; ``want_sk`` returns the sigmas used as the last element."""
        clc = f64(clc)
        single = clc.ndim == 1
        clc = np.atleast_2d(clc)
        B, N = clc.shape
        cle = None if cle is None else np.atleast_2d(f64(cle))
        spec = None
        if sk is None:
            if noise is None:
                raise ValueError("either sk or a noise specification is required")
            spec = C.byref(noise[0])
        else:
            sk = f64(np.atleast_1d(sk))
            if sk.size != B:
                raise ValueError("batch shapes of clc and sk do not agree")
        if cle is not None and cle.shape != clc.shape:
            raise ValueError("batch shapes of clc and cle do not agree")
        k1 = N if k1 is None else int(k1)
        k0 = int(k0)
        if not 0 <= k0 < k1 <= N:
            raise ValueError("bad sample range [%d, %d) for a series of %d samples" % (k0, k1, N))
        lnO = np.empty((B, k1 - k0))
        marg = np.empty((self.nhyp + 1, B, k1 - k0)) if want_marg else None
        sk_out = np.empty(B) if want_sk else None
        check(self.lib.bfl_detector_odds(self.ctx.handle, self.handle, 1 if noisepoly else 0, ptr(clc), ptr(cle),
                                         ptr(sk), spec, B, N, int(k0), k1, ptr(lnO), ptr(marg), ptr(sk_out)))
        if single:
            lnO = lnO[0]
            marg = None if marg is None else marg[:, 0]
        out = (lnO, marg) if want_marg else (lnO,)
        if want_sk:
            out = out + (sk_out,)
        return out if len(out) > 1 else out[0]

    def detector_odds_dev(self, d_clc, d_cle, const_var, d_sk, B, N, seg0, seglen, k0, k1, d_lnO, noisepoly=True, ctx=None):
        """Device-resident, asynchronous variant: arguments are raw device addresses (ints).  ``ctx``: run on
        another context of the same device (its stream, its scratch) -- the plan itself is read-only."""
        check(self.lib.bfl_detector_odds_dev((ctx or self.ctx).handle, self.handle, 1 if noisepoly else 0, d_clc, d_cle or None,
                                             1 if const_var else 0, d_sk, int(B), int(N), int(seg0), int(seglen),
                                             int(k0), int(k1), d_lnO))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.bfl_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipe:
    """``bfl_pipe``: asynchronous batch search.  ``submit`` takes raw host addresses (ints, e.g. of page-locked
    torch / NumPy buffers that the caller keeps alive) and returns a job number; ``wait`` blocks until that
    job's outputs are in host memory and returns the number of region records."""

    def __init__(self, plan, B, N, k0, k1, noise=None, noisepoly=True, depth=3, max_regions=None):
        self.plan, self.lib = plan, plan.lib
        self.B, self.N, self.k0, self.k1 = int(B), int(N), int(k0), int(k1)
        self.max_regions = int(8 * B if max_regions is None else max_regions)
        self._noise = noise          # keeps the coefficient array alive
        h = C.c_void_p()
        check(self.lib.bfl_pipe_create(plan.ctx.handle, plan.handle, 1 if noisepoly else 0,
                                       C.byref(noise[0]) if noise is not None else None, self.B, self.N, self.k0, self.k1,
                                       int(depth), self.max_regions, C.byref(h)))
        self.handle = h

    def submit(self, clc_addr, sk_addr=None, out_lnO_addr=None, thresh=float("nan"), out_regions_addr=None,
               out_counts_addr=None, out_sk_addr=None):
        job = C.c_int64(-1)
        check(self.lib.bfl_pipe_submit(self.handle, clc_addr, sk_addr, out_lnO_addr, float(thresh), out_regions_addr,
                                       out_counts_addr, out_sk_addr, C.byref(job)))
        return job.value

    def wait(self, job):
        n = C.c_int32(0)
        check(self.lib.bfl_pipe_wait(self.handle, int(job), C.byref(n)))
        return n.value

    @property
    def launches(self):
        return int(self.lib.bfl_pipe_launch_count(self.handle))

    @property
    def graphs(self):
        """CUDA graphs instantiated for recurring (slot, host buffers) combinations."""
        return int(self.lib.bfl_pipe_graph_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.bfl_pipe_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def noise_sigma(ctx, clc, spec):
    """Noise sigma of every curve of ``clc[B, N]`` on the device (``bfl_noise_sigma``)."""
    clc = np.atleast_2d(f64(clc))
    B, N = clc.shape
    out = np.empty(B)
    check(ctx.lib.bfl_noise_sigma(ctx.handle, C.byref(spec[0]), ptr(clc), B, N, ptr(out)))
    return out


def threshold(ctx, lnO, thresh, expand=0):
    lnO = f64(lnO)
    n = lnO.size
    cap = max(16, n // 2 + 2)
    segs = np.zeros((cap, 2), dtype=np.int64)
    mv = np.zeros(cap)
    mi = np.zeros(cap, dtype=np.int64)
    ns = C.c_int32(0)
    check(ctx.lib.bfl_threshold(ctx.handle, ptr(lnO), n, float(thresh), int(expand), segs.ctypes.data_as(_i64p),
                                ptr(mv), mi.ctypes.data_as(_i64p), cap, C.byref(ns)))
    k = ns.value
    return segs[:k], mv[:k], mi[:k]


def pe_grid(ctx, kind, reverse, cts, clc, params, logprior, npoly, polyms, sigma, margbackground, sg=None):
    """``sg = (nbins, order)``: the chunk was Savitzky-Golay detrended, filter every model the same way."""
    cts, clc, params, logprior = f64(cts), f64(clc), f64(params), f64(logprior)
    G = logprior.size
    polyms = None if polyms is None else f64(polyms)
    out = np.empty(G)
    coef = None
    if sg is not None:
        from .noise import _sg_coefficients
        coef = f64(_sg_coefficients(int(sg[0]), int(sg[1])))
    check(ctx.lib.bfl_pe_grid(ctx.handle, MODEL_KINDS[kind], 1 if reverse else 0, ptr(cts), ptr(clc), cts.size,
                              ptr(params), ptr(logprior), G, int(npoly), ptr(polyms), float(sigma),
                              1 if margbackground else 0, ptr(coef), 0 if coef is None else coef.size, ptr(out)))
    return out


def whole_series_lnB(ctx, kind, reverse, halfrange, cts, clc, noisevar, params, logprior, bgmodels, bgcross, dbgr,
                     sumlogsigma, want_bg=False):
    """``bfl_whole_series_lnB``: log Bayes factors ``[G, N]`` with one polynomial background over the whole series
    (``params=None``: only the polynomial-only factor).  Returns ``(lnB or None, bg or None)``."""
    cts, clc, noisevar = f64(cts), f64(clc), f64(noisevar)
    bgmodels, bgcross, dbgr = f64(bgmodels), f64(bgcross), f64(dbgr)
    N = cts.size
    G = 0 if params is None else f64(logprior).size
    params = None if params is None else f64(params)
    logprior = None if logprior is None else f64(logprior)
    out = np.empty((G, N)) if G else None
    bg = np.zeros(1) if (want_bg or not G) else None
    check(ctx.lib.bfl_whole_series_lnB(ctx.handle, MODEL_KINDS[kind] if kind else 0, 1 if reverse else 0, 1 if halfrange else 0,
                                       ptr(cts), ptr(clc), ptr(noisevar), N, ptr(params), ptr(logprior), G, bgmodels.shape[0],
                                       ptr(bgmodels), ptr(bgcross), ptr(dbgr), float(sumlogsigma), ptr(out), ptr(bg)))
    return out, (None if bg is None else float(bg[0]))


def logtrapz_axis0(ctx, lx, t):
    lx, t = f64(lx), f64(t)
    n = lx.shape[0]
    m = int(np.prod(lx.shape[1:])) if lx.ndim > 1 else 1
    out = np.empty(m)
    check(ctx.lib.bfl_logtrapz(ctx.handle, ptr(lx), ptr(t), n, m, ptr(out)))
    return out.reshape(lx.shape[1:])


def fp64_peak(ctx, iters=4096):
    tf, ms = C.c_double(0), C.c_double(0)
    check(ctx.lib.bfl_fp64_peak_probe(ctx.handle, int(iters), C.byref(tf), C.byref(ms)))
    return tf.value, ms.value
