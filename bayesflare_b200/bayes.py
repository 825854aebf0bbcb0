"""``Bayes`` and ``ParameterEstimationGrid`` -- drop-in mirrors of ``bayesflare/stats/bayes.py``
whose numerical work is done by ``libbfl`` on the GPU.

Host side (this file): argument checking, the scalar noise estimate, the window time stamps,
grid axes, priors and trapezium weights.  Device side: templates, every cross term, the
amplitude marginalisation for every (sample, grid point), the grid marginalisation.
"""
from __future__ import annotations

from copy import copy, deepcopy
from math import pi

import numpy as np

from . import _lib
from .general import logtrapz
from .noise import estimate_noise_ps, estimate_noise_tv

__all__ = ["Bayes", "ParameterEstimationGrid", "window_times", "background_models", "estimate_sigma",
           "hypothesis_from_model"]


# ---- host-side pieces of bayes_factors_marg_poly_bgd shared with the detector ------------
def window_times(ts, bglen, model_t0):
    """The ``bglen`` time stamps the templates are evaluated on (bayes.py:261-276)."""
    N = len(ts)
    nsteps = int(bglen / 2)
    dt = ts[1] - ts[0]
    idxt0 = int((model_t0 - ts[0]) / dt) + 1
    idx1, idx2 = idxt0 - nsteps, idxt0 + nsteps + 1
    if idx1 < 0:
        return ts[:bglen]
    if idx2 > N - 1:
        return ts[-bglen:]
    return ts[idx1:idx2]


def background_models(bglen, bgorder):
    """Monomials on [0, 1] (bayes.py:292-307), computed by NumPy exactly as the reference does."""
    tsp = np.linspace(0., 1., bglen)
    return np.array([tsp ** i for i in range(bgorder + 1)])


def estimate_sigma(lightcurve, bglen, bgorder, noiseestmethod, psestfrac, tvsigma):
    """Noise sigma on a Savitzky-Golay detrended copy (bayes.py:235-246).  Returns ``None`` for
    an unknown method, after printing the reference's message."""
    tmpcurve = deepcopy(lightcurve)
    if tmpcurve.detrended is False:
        tmpcurve.detrend(method='savitzkygolay', nbins=bglen, order=bgorder)
    if noiseestmethod == 'powerspectrum':
        return estimate_noise_ps(tmpcurve, estfrac=psestfrac)[0]
    if noiseestmethod == 'tailveto':
        return estimate_noise_tv(tmpcurve.clc, sigma=tvsigma)[0]
    print("Noise estimation method must be 'powerspectrum' or 'tailveto'")
    return None


def hypothesis_from_model(model, mts, halfrange, ranges=None):
    """Describe ``model``'s parameter grid for ``libbfl`` (``bfl_hypothesis`` in include/bfl.h)."""
    if ranges is not None:
        model = copy(model)
        model.ranges = ranges
    logprior = model.grid_logprior() + (np.log(0.5) if halfrange else 0.)    # bayes.py:284-288
    h = dict(kind=model.device_kind or "table", reverse=bool(model.reverse), halfrange=bool(halfrange),
             params=model.grid_params(), logprior=logprior, logweight=model.grid_logweight(), table=None)
    if model.device_kind is None:
        # a user-defined Model subclass: evaluate its templates on the host, upload as a table
        G = int(np.prod(model.shape))
        tab = np.zeros((G, len(mts)))
        for q in range(G):
            if np.isfinite(logprior[q]):
                tab[q] = model(q, ts=mts, filt=False).clc
        h["table"] = tab
    return h


def _sumlogsigma(noisevar):
    sk = np.sqrt(noisevar)                      # bayes.py:419
    return float(np.sum(np.log(sk[np.isfinite(sk)])))   # general.pyx:216-221


class Bayes:
    """Bayesian odds ratios of a model against Gaussian noise for every sample of a light curve
    (bayes.py:76-639).

    Parameters
    ----------
    lightcurve : Lightcurve-like (``cts, clc, cle, detrended, detrend()``)
    model : a :class:`Model` with a parameter grid
    """

    premarg = {}

    def __init__(self, lightcurve, model):
        self.lightcurve = lightcurve
        self.model = deepcopy(model)
        self.ranges = deepcopy(model.ranges)
        self.confidence = 0.999
        self.noise_ev = self.noise_evidence()

    # bayes.py:623-639
    def noise_evidence(self):
        var = estimate_noise_tv(self.lightcurve.clc, 1.0)[0] ** 2
        clc = self.lightcurve.clc
        return -0.5 * len(clc) * np.log(2. * pi * var) - np.sum(clc ** 2) / (2. * var)

    def _check(self, bglen, nsinusoids):
        if nsinusoids != 0:
            # the reference needs Lightcurve.sinusoid_freqs here, which no Lightcurve defines (bayes.py:278): it raises too
            raise NotImplementedError("sinusoidal background terms (nsinusoids != 0, bayes.py:278-307) are not supported")
        if bglen is None:
            return True
        if bglen % 2 == 0:
            print("Error... Background length (bglen) must be an odd number")      # bayes.py:229-233
            return False
        return True

    def bayes_factors_marg_poly_bgd(self, bglen=55, bgorder=4, nsinusoids=0, noiseestmethod='powerspectrum',
                                    psestfrac=0.5, tvsigma=1.0, halfrange=True, ncpus=None):
        """Log Bayes factor of "model + polynomial background" versus Gaussian noise for every
        grid point and every sample, amplitudes marginalised analytically (bayes.py:153-437).
        Sets ``self.lnBmargAmp`` (shape ``model.shape + (N,)``) and ``self.premarg``.
        ``ncpus`` is accepted for compatibility and ignored."""
        if not self._check(bglen, nsinusoids):
            return None
        if bglen is None:
            return self._whole_series(bgorder, noiseestmethod, psestfrac, tvsigma, halfrange, signal=True)
        sk = estimate_sigma(self.lightcurve, bglen, bgorder, noiseestmethod, psestfrac, tvsigma)
        if sk is None:
            return None
        lc, model = self.lightcurve, self.model
        N = len(lc.cts)
        if bglen > N:
            print("Error... bglen is greater than the data length!")
            return None
        mts = window_times(model.ts, bglen, model.t0)
        noisevar = sk ** 2 + np.asarray(lc.cle, dtype=float) ** 2             # bayes.py:312
        hyp = hypothesis_from_model(model, mts, halfrange, ranges=self.ranges)
        ctx = _lib.default_context()
        plan = _lib.Plan(ctx, bglen, background_models(bglen, bgorder), mts, [hyp])
        try:
            lnB, _ = plan.window_lnB(0, lc.clc, lc.cle, sk, _sumlogsigma(noisevar))
        finally:
            plan.close()
        self.lnBmargAmp = lnB.reshape(tuple(model.shape) + (N,))
        self.premarg = np.copy(self.lnBmargAmp)

    def bayes_factors_marg_poly_bgd_only(self, bglen=55, bgorder=4, nsinusoids=0, noiseestmethod='powerspectrum',
                                         psestfrac=0.5, tvsigma=1.0):
        """Log Bayes factor of a polynomial background alone versus Gaussian noise (bayes.py:439-563)."""
        if not self._check(bglen, nsinusoids):
            return None
        if bglen is None:
            return self._whole_series(bgorder, noiseestmethod, psestfrac, tvsigma, False, signal=False)
        lc = self.lightcurve
        N = len(lc.cts)
        if bglen > N:
            print("Error... bglen is greater than the data length!")
            return None
        sk = estimate_sigma(lc, bglen, bgorder, noiseestmethod, psestfrac, tvsigma)
        if sk is None:
            return None
        noisevar = sk ** 2 + np.asarray(lc.cle, dtype=float) ** 2
        ctx = _lib.default_context()
        plan = _lib.Plan(ctx, bglen, background_models(bglen, bgorder), np.asarray(lc.cts[:bglen], dtype=float), [])
        try:
            _, B = plan.window_lnB(0, lc.clc, lc.cle, sk, _sumlogsigma(noisevar), want_lnB=False, want_bg=True)
        finally:
            plan.close()
        self.lnBmargBackground = B
        return B

    def _whole_series(self, bgorder, noiseestmethod, psestfrac, tvsigma, halfrange, signal):
        """``bglen=None`` (bayes.py:292-294, 330-331, 394-395, 408-409): ONE polynomial background over the whole
        series -- a single Gram matrix -- and the full-length template slid along it.  The series' constants (Gram
        matrix, data projections, sum of log sigma) are NumPy expressions of the reference taken over as they are;
        the O(G N^2) template terms and every per-sample marginalisation run on the device
        (``bfl_whole_series_lnB``).  As in the reference, the light curve must already be detrended: the noise
        estimate would otherwise ask for a Savitzky-Golay window of ``None`` samples."""
        lc, model = self.lightcurve, self.model
        if lc.detrended is False:
            raise ValueError("bglen=None needs a light curve that is already detrended (the reference's noise estimate "
                             "calls savitzky_golay with a window of None samples otherwise and raises)")
        sk = estimate_sigma(lc, None, bgorder, noiseestmethod, psestfrac, tvsigma)
        if sk is None:
            return None
        N = len(lc.cts)
        npoly = bgorder + 1
        tsp = np.linspace(0., 1., N)
        bgmodels = np.array([tsp ** i for i in range(npoly)])
        noisevar = sk ** 2 + np.asarray(lc.cle, dtype=float) ** 2
        bgcross = np.zeros((npoly, npoly))
        for i in range(npoly):
            for j in range(i, npoly):
                bgcross[i, j] = np.sum(bgmodels[i] * bgmodels[j] / noisevar)
        d = np.asarray(lc.clc, dtype=float) / noisevar
        dbgr = np.array([np.sum(d * bgmodels[i]) for i in range(npoly)])
        ctx = _lib.default_context()
        if not signal:
            _, bg = _lib.whole_series_lnB(ctx, None, False, False, lc.cts, lc.clc, noisevar, None, None, bgmodels, bgcross,
                                          dbgr, _sumlogsigma(noisevar))
            self.lnBmargBackground = np.full(N, bg)
            return self.lnBmargBackground
        if model.device_kind is None:
            raise NotImplementedError("bglen=None supports the built-in models")
        hyp = hypothesis_from_model(model, np.asarray(model.ts, dtype=float), halfrange, ranges=self.ranges)
        lnB, _ = _lib.whole_series_lnB(ctx, hyp["kind"], hyp["reverse"], halfrange, model.ts, lc.clc, noisevar, hyp["params"],
                                       hyp["logprior"], bgmodels, bgcross, dbgr, _sumlogsigma(noisevar))
        self.lnBmargAmp = lnB.reshape(tuple(model.shape) + (N,))
        self.premarg = np.copy(self.lnBmargAmp)

    def marginalise(self, pname):
        """Integrate ``lnBmargAmp`` over one parameter with the log-space trapezium rule, or drop
        the axis if the parameter is fixed (bayes.py:565-604).  Returns a new :class:`Bayes`."""
        arr = self.lnBmargAmp
        places = self.ranges[pname]
        axis = self.model.paramnames.index(pname)
        if len(places) > 1:
            moved = np.moveaxis(arr, axis, 0)
            x = _lib.logtrapz_axis0(_lib.default_context(), moved, places)
        else:
            x = np.reshape(arr, tuple(s for i, s in enumerate(arr.shape) if i != axis))
        model = copy(self.model)
        model.paramnames = [p for p in self.model.paramnames if p != pname]
        model.shape = [s for i, s in enumerate(self.model.shape) if i != axis]
        B = Bayes(self.lightcurve, model)
        ranges = copy(self.ranges)
        del ranges[pname]
        B.ranges = ranges
        B.lnBmargAmp = x
        return B

    def marginalise_full(self):
        """bayes.py:606-621"""
        A = self
        for p in list(self.ranges):
            A = A.marginalise(p)
        return A


class ParameterEstimationGrid:
    """Posterior of a model's parameters on a grid for a short light-curve chunk
    (bayes.py:688-1391); the likelihood of every grid point is evaluated on the GPU."""

    def __init__(self, modelType=None, lightcurve=None):
        self.modelType = None
        self.model = None
        self.paramNames = None
        self.paramValues = {}
        self.noiseSigma = None
        self.prior = None
        self.posterior = None
        self.margposteriors = {}
        self.maxposterior = None
        self.maxpostparams = {}
        if lightcurve is None:
            print("A lightcurve is required as input")
        else:
            self.lightcurve = deepcopy(lightcurve)
        if modelType is None:
            print("Specify the model type with set_model_type()")
        elif modelType.lower() in ('flare', 'transit'):
            self.set_model_type(modelType)
        else:
            print("Unknown model type")
        self.set_sigma(self.lightcurve)

    def set_model_type(self, modelType):
        """bayes.py:735-769"""
        from . import models
        self.modelType = modelType.lower()
        cls = {'flare': models.Flare, 'transit': models.Transit, 'gaussian': models.Gaussian,
               'expdecay': models.Expdecay, 'impulse': models.Impulse, 'step': models.Step}[self.modelType]
        self.model = cls(self.lightcurve.cts)
        self.paramNames = self.model.paramnames
        self.prior = self.model.prior

    def set_grid(self, ranges={}):
        """Grid axes from a dictionary of ranges (bayes.py:771-816): ``(lo, hi, n)`` -> ``np.linspace``, a single
        value -> a one-point axis, which the reference stores as float32 -- kept, it changes the value."""
        if not ranges:
            print("Must specify a dictionary of ranges")
            return
        if not self.model.ranges:
            self.model.set_params(ranges)
        missing = [p for p in self.paramNames if p not in ranges]
        if missing:
            print("Error. The parameter %s is not in the dictionary" % missing[0])
            return
        for p in self.paramNames:
            spec = ranges[p] if isinstance(ranges[p], tuple) else (ranges[p],)
            if len(spec) == 1:
                self.paramValues[p] = np.array(spec, dtype='float32')
            elif len(spec) == 3:
                lo, hi, n = spec
                if not (lo < hi and n > 1):
                    print("%s range has an upper bound smaller than the lower bound! Try again." % p)
                    return
                self.paramValues[p] = np.linspace(lo, hi, int(n))

    def lightcurve_chunk(self, centidx, length):
        """Cut the light curve down to ``length`` samples centred on ``centidx`` (bayes.py:818-846): the window is
        [centidx - length//2, centidx + length//2], clipped to the series (a stop on the last sample moves to the end)."""
        lc = self.lightcurve
        n = len(lc.cts)
        half = int(length / 2)
        stop = centidx + half + 1
        sel = slice(max(centidx - half, 0), n if stop > n - 1 else stop)
        lc.cts, lc.clc, lc.cle = lc.cts[sel], lc.clc[sel], lc.cle[sel]

    def lightcurve_dynamic_range(self):
        return (np.amin(self.lightcurve.clc), np.amax(self.lightcurve.clc))

    def set_sigma(self, lightcurve, detrend=True, dtlen=55, dtorder=4, noiseestmethod='powerspectrum', estfrac=0.5,
                  tvsigma=1.0):
        """Noise sigma of the (optionally Savitzky-Golay detrended) curve (bayes.py:862-910)."""
        work = copy(lightcurve)
        if detrend:
            work.detrend(method='savitzkygolay', nbins=dtlen, order=dtorder)
        estimators = {'powerspectrum': lambda: estimate_noise_ps(work, estfrac=estfrac)[0],
                      'tailveto': lambda: estimate_noise_tv(work.clc, sigma=tvsigma)[0],
                      'std': lambda: np.std(work.clc)}
        if noiseestmethod not in estimators:
            print("Noise estimation method must be 'powerspectrum' or 'tailveto'")
            self.noiseSigma = None
            return
        self.noiseSigma = estimators[noiseestmethod]()

    def calculate_posterior(self, paramValues=None, lightcurve=None, sigma=None, margbackground=True, bgorder=4,
                            ncpus=None):
        """Unnormalised log posterior on the grid (bayes.py:912-1055).  ``ncpus`` is ignored."""
        if paramValues is not None:
            pv = paramValues
            for p in self.paramNames:
                if p not in pv:
                    print("Parameter %s is not in the supplied paramValues" % p)
                    return None
            if len(pv) != len(self.paramNames):
                print("Input parameter values dictionary is not the right length!")
                return None
        else:
            pv = self.paramValues
        lc = lightcurve if lightcurve is not None else self.lightcurve
        # a Savitzky-Golay detrended chunk: every model is filtered the same way (bayes.py:1018-1021)
        sg = None
        if getattr(lc, "detrended", False) and getattr(lc, "detrend_method", None) == 'savitzkygolay':
            sg = (getattr(lc, "detrend_nbins", None) or lc.detrend_length, lc.detrend_order)
        # NOTE (as the reference, bayes.py:1039 and :1046): the likelihood always uses
        # self.noiseSigma, even when ``sigma`` is passed.
        sk = self.noiseSigma
        dl = len(lc.cts)
        sp = [len(pv[p]) for p in self.paramNames]
        mesh = np.meshgrid(*[np.asarray(pv[p], dtype=float) for p in self.paramNames], indexing="ij")
        params = np.zeros((int(np.prod(sp)), 4))
        for i, m in enumerate(mesh):
            params[:, i] = m.ravel()
        if self.prior == self.model.prior:
            logprior = self.model.grid_logprior(values=pv)          # vectorised for the built-in models
        else:
            pdicts = ({p: pv[p][q[i]] for i, p in enumerate(self.paramNames)} for q in np.ndindex(*sp))
            logprior = np.array([self.prior(pd) for pd in pdicts], dtype=float)
        polyms = None
        if margbackground:
            ts01 = np.linspace(0, 1, dl)
            polyms = np.array([ts01 ** i for i in range(bgorder + 1)])
        post = _lib.pe_grid(_lib.default_context(), self.model.device_kind, self.model.reverse, lc.cts, lc.clc,
                            params, logprior, bgorder + 1, polyms, sk, margbackground, sg=sg)
        self.posterior = post.reshape(tuple(sp))

    # ---- post-processing (bayes.py:1057-1301): marginals by log-trapezium integration on the device ----
    def _axis(self, name):
        return np.asarray(self.paramValues[name], dtype=float)

    def _integrate_out(self, drop):
        """Log posterior with the axes named in ``drop`` integrated out (``bfl_logtrapz``, one launch per axis, in
        the order of ``paramNames`` as the reference does); one-point axes are simply removed (bayes.py:1098-1101).
        Returns the array and the names of the axes it still has."""
        ctx = _lib.default_context()
        logp, names = np.asarray(self.posterior, dtype=float), list(self.paramNames)
        for name in [p for p in self.paramNames if p in drop]:
            ax = names.index(name)
            if logp.shape[ax] == 1:
                logp = np.take(logp, 0, axis=ax)
            else:
                logp = _lib.logtrapz_axis0(ctx, np.ascontiguousarray(np.moveaxis(logp, ax, 0)), self._axis(name))
            del names[ax]
        return logp, names

    def marginalised_posterior(self, parameter=None):
        """Normalised posterior of one parameter (bayes.py:1057-1124)."""
        if parameter not in self.paramNames:
            print("Given parameter (%s) is not in model" % parameter)
            return None
        if self.posterior is None:
            print("Posterior not yet defined!")
            return None
        name = parameter.lower()
        logp, _ = self._integrate_out([p for p in self.paramNames if p != name])
        logp = np.atleast_1d(logp)
        # a one-point axis is "normalised" by the constant 1, as the reference does (bayes.py:1117-1120)
        lognorm = logtrapz(logp, self._axis(name)) if logp.size > 1 else 1.
        self.margposteriors[name] = np.exp(logp - lognorm)
        return np.copy(self.margposteriors[name])

    def marginalise_all(self):
        for p in self.paramNames:
            self.marginalised_posterior(p)

    def marginalised_posterior_2D(self, parameters=None):
        """Normalised joint posterior of two parameters, indexed [parameters[0], parameters[1]] (bayes.py:1135-1210)."""
        if not isinstance(parameters, list) or len(parameters) != 2:
            print("Must supply a list of two parameters")
            return None
        pair = [p.lower() for p in parameters]
        for given, p in zip(parameters, pair):
            if p not in self.paramNames:
                print("Given parameter (%s) is not in model" % given)
                return None
        if self.posterior is None:
            print("Posterior not yet defined!")
            return None
        if len(self.paramNames) < 3:
            print("No need to marginalise, posterior is already 2d or 1d")
            return None
        logp, left = self._integrate_out([p for p in self.paramNames if p not in pair])
        if left != pair:
            logp = logp.T
        ctx = _lib.default_context()
        lognorm = logtrapz(_lib.logtrapz_axis0(ctx, np.ascontiguousarray(logp), self._axis(pair[0])), self._axis(pair[1]))
        return np.exp(logp - lognorm)

    def maximum_posterior(self):
        """Largest grid value of the log posterior and the parameters there (bayes.py:1212-1236)."""
        if self.posterior is None:
            print("Posterior not defined")
            return None
        at = np.unravel_index(int(np.argmax(self.posterior)), self.posterior.shape)
        self.maxposterior = self.posterior[at]
        self.maxpostparams.update({p: self.paramValues[p][i] for p, i in zip(self.paramNames, at)})
        return self.maxposterior, self.maxpostparams

    def _best_fit_curve(self):
        if self.maxposterior is None:
            self.maximum_posterior()
        return self.model.model(self.maxpostparams, ts=self.lightcurve.cts)

    def maximum_posterior_snr(self):
        """Optimal signal-to-noise ratio of the maximum-posterior model (bayes.py:1238-1266)."""
        return np.sqrt(np.sum(self._best_fit_curve() ** 2)) / self.noiseSigma

    def maximum_posterior_ew(self):
        """Equivalent width of the maximum-posterior model (bayes.py:1268-1301)."""
        best = self._best_fit_curve()
        trapezoid = getattr(np, 'trapezoid', None) or np.trapz
        return trapezoid(best, self.lightcurve.cts) / np.median(self.lightcurve.clc - best + self.lightcurve.dc)
