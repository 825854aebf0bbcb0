"""Signal / noise template models -- host-side mirror of ``bayesflare/models/model.py``.

Same class names, constructor signatures, ``paramnames`` orders, ``.model(pdict, ts)``,
``.prior(pdict)`` and ``__call__(q, ts)`` as the reference, so code written against
``bayesflare.Flare`` etc. keeps working.  These NumPy evaluations are the *user-facing*
interface (injection, plotting, custom scripts).  The detector does not call them: it
hands the parameter grid to ``libbfl`` and the templates are generated on the GPU
(``k_gen_templates`` in ``csrc/bfl_kernels.cu``).
"""
from __future__ import annotations

from math import floor

import numpy as np

__all__ = ["Model", "Flare", "Transit", "Expdecay", "Impulse", "Gaussian", "Step", "ModelCurve"]


def _centre(ts):
    # "t0 == inf" convention of every model: ts[int(len(ts)/2.)]  (model.py:222-223)
    return ts[int(len(ts) / 2.)]


def _flat_logprior(rng):
    return -np.log(rng[-1] - rng[0]) if len(rng) > 1 else 0.


class ModelCurve:
    """model.py:1022-1048"""

    def __init__(self, ts, lc):
        self.clc = lc
        self.cts = ts

    def dt(self):
        return self.cts[1] - self.cts[0]

    def fs(self):
        return 1.0 / self.dt()


class Model:
    """Generic model with a parameter grid (model.py:7-153)."""

    #: device template generator used for this model (see include/bfl.h); ``None`` means the
    #: templates are evaluated with ``self.model`` on the host and uploaded as a table
    device_kind = None

    def __init__(self, ts, mtype, amp=1, t0=None, reverse=False, paramnames=None, paramranges=None):
        if t0 is None:
            t0 = ts[floor(len(ts) / 2)]
        self.mtype = mtype.lower()
        self.modelname = self.mtype
        self.paramnames = list(paramnames) if paramnames is not None else []
        self.amp = amp
        self.t0 = t0
        self.ts = ts
        self.reverse = reverse
        self.shape = []
        self.ranges = {}
        if paramranges is not None:
            self.set_params(paramranges)

    def __str__(self):
        return "<BayesFlare " + self.mtype + " model>"

    __repr__ = __str__

    def set_params(self, paramrangedict):
        """Grid of parameter values: 1-tuples are fixed values, 3-tuples are
        ``linspace(lo, hi, n)`` (model.py:69-92)."""
        for p in self.paramnames:
            r = paramrangedict[p]
            if len(r) == 1:
                self.ranges[p] = np.array([r[0]])
            elif len(r) == 3:
                self.ranges[p] = np.linspace(r[0], r[1], r[2])
            else:
                raise ValueError("Error... range must either contain 1 or 3 values")
            self.shape.append(len(self.ranges[p]))

    def filter_model(self, m, filtermethod='savitzkygolay', nbins=101, order=3, filterknee=(1. / (0.3 * 86400.))):
        """model.py:94-131"""
        from .noise import savitzky_golay, running_median, highpass_filter_lightcurve
        if filtermethod == 'savitzkygolay':
            return m - savitzky_golay(m, nbins, order)
        if filtermethod == 'runningmedian':
            return m - running_median(m, nbins)
        if filtermethod == 'highpass':
            from .lightcurve import Lightcurve
            return highpass_filter_lightcurve(Lightcurve(cts=np.copy(self.ts), clc=np.copy(m)), knee=filterknee)
        raise ValueError('Unrecognised filter method (%s) given' % filtermethod)

    def _pdict(self, q):
        idx = np.unravel_index(q, self.shape)
        return {p: self.ranges[p][idx[i]] for i, p in enumerate(self.paramnames)}

    def __call__(self, q, ts=None, filt=False, filtermethod='savitzkygolay', nbins=101, order=3,
                 filterknee=(1. / (0.3 * 86400.))):
        if ts is None:
            ts = self.ts
        f = self.model(self._pdict(q), ts=ts)
        if filt:
            f = self.filter_model(f, filtermethod=filtermethod, nbins=nbins, order=order, filterknee=filterknee)
        return ModelCurve(ts, f)

    def _need(self, pdict):
        for p in self.paramnames:
            if p not in pdict:
                raise ValueError("Error... no '%s' value in dictionary!" % p)

    # ---- what the engine needs ---------------------------------------------------------
    def grid_params(self):
        """(G, 4) float64 array: parameter values of every grid point (C order), columns in
        ``paramnames`` order, zero padded -- the layout ``bfl_hypothesis.params`` expects."""
        G = int(np.prod(self.shape))
        out = np.zeros((G, 4))
        mesh = np.meshgrid(*[self.ranges[p] for p in self.paramnames], indexing="ij")
        for i, m in enumerate(mesh):
            out[:, i] = m.ravel()
        return out

    def grid_logprior(self, values=None):
        """log prior of every grid point (C order).  ``values`` optionally replaces the parameter
        axes (``ParameterEstimationGrid`` evaluates the prior on its own grid while the prior's
        normalisation comes from ``self.ranges``, as the reference does).  Built-in models override
        :meth:`_logprior_mesh` with a vectorised form; this generic version calls ``prior`` per point."""
        axes = [np.asarray((values or self.ranges)[p]) for p in self.paramnames]
        fast = self._logprior_mesh(dict(zip(self.paramnames, np.meshgrid(*axes, indexing="ij"))))
        if fast is not None:
            return np.asarray(fast, dtype=float).ravel()
        shape = [len(a) for a in axes]
        out = np.empty(int(np.prod(shape)))
        for q in range(out.size):
            idx = np.unravel_index(q, shape)
            out[q] = self.prior({p: axes[i][idx[i]] for i, p in enumerate(self.paramnames)})
        return out

    def _logprior_mesh(self, mesh):
        """Vectorised ``prior`` over a dict of broadcast parameter arrays, or ``None``."""
        return None

    def _flat_total(self):
        return sum(_flat_logprior(self.ranges[p]) for p in self.paramnames)

    def grid_logweight(self):
        """log of the product trapezium weight of every grid point: marginalising each
        multi-valued axis in turn with ``logtrapz`` (general.pyx:576-607) equals a weighted
        log-sum-exp with these weights."""
        logw = []
        for p in self.paramnames:
            x = np.asarray(self.ranges[p], dtype=float)
            if len(x) > 1:
                d = np.diff(x)
                w = np.empty(len(x))
                w[0], w[-1] = d[0] / 2., d[-1] / 2.
                w[1:-1] = (d[:-1] + d[1:]) / 2.
                with np.errstate(divide="ignore", invalid="ignore"):
                    logw.append(np.log(w))
            else:
                logw.append(np.zeros(1))
        mesh = np.meshgrid(*logw, indexing="ij")
        return sum(m for m in mesh).ravel()


class Flare(Model):
    """Gaussian rise + exponential decay (model.py:156-329)."""
    device_kind = "flare"

    def __init__(self, ts, paramranges=None, amp=1, t0=None, reverse=False):
        Model.__init__(self, ts, mtype='flare', amp=amp, t0=t0, reverse=reverse,
                       paramnames=['t0', 'tauexp', 'taugauss', 'amp'], paramranges=paramranges)

    def model(self, pdict, ts=None):
        self._need(pdict)
        if ts is None:
            ts = self.ts
        ts = np.asarray(ts)
        t0, amp, tg, te = pdict['t0'], pdict['amp'], pdict['taugauss'], pdict['tauexp']
        if t0 == np.inf:
            t0 = _centre(ts)
        f = np.zeros(len(ts))
        f[ts == t0] = amp
        rise, decay = (ts > t0, ts < t0) if self.reverse else (ts < t0, ts > t0)
        if tg > 0:
            f[rise] = amp * np.exp(-(ts[rise] - t0) ** 2 / (2 * float(tg) ** 2))
        if te > 0:
            if self.reverse:
                f[decay] = amp * np.exp((ts[decay] - t0) / float(te))
            else:
                f[decay] = amp * np.exp(-(ts[decay] - t0) / float(te))
        return f

    def prior(self, pdict):
        """Flat over the grid with tau_gauss <= tau_exp (model.py:254-329)."""
        self._need(pdict)
        tg, te = pdict['taugauss'], pdict['tauexp']
        tgr, ter = self.ranges['taugauss'], self.ranges['tauexp']
        base = _flat_logprior(self.ranges['amp']) + _flat_logprior(self.ranges['t0'])
        if tg > te or tg > ter[-1]:
            return base + (-np.inf)
        tgmin, tgmax, temin, temax = tgr[0], tgr[-1], ter[0], ter[-1]
        dtg, dte = tgmax - tgmin, temax - temin
        if tgmin <= temin and tgmax <= temax:       # rectangle minus lower triangle
            area = dte * dtg - 0.5 * (tgmax - temin) ** 2
        elif tgmin > temin and tgmax > temax:       # upper triangle
            area = 0.5 * (temax - tgmin) ** 2
        elif tgmin > temin and tgmax < temax:       # upper trapezium
            area = 0.5 * dtg * ((temax - tgmin) + (temax - tgmax))
        elif tgmin < temin and tgmax > temax:       # lower trapezium
            area = 0.5 * dte * ((temin - tgmin) + (temax - tgmin))
        else:
            raise UnboundLocalError("parea: the reference defines no prior area for this grid geometry")
        return base + (-np.log(area))


def _flare_logprior_mesh(self, mesh):
    # same value as Flare.prior for every point: constant inside the allowed region, -inf outside
    tg, te = np.asarray(mesh['taugauss'], dtype=float), np.asarray(mesh['tauexp'], dtype=float)
    bad = (tg > te) | (tg > self.ranges['tauexp'][-1])
    out = np.full(tg.shape, -np.inf)
    if not bad.all():
        i = np.argmin(bad.ravel())
        ok = self.prior({p: np.asarray(mesh[p]).ravel()[i] for p in self.paramnames})
        out[~bad] = ok
    return out


Flare._logprior_mesh = _flare_logprior_mesh


class Transit(Model):
    """Flat-bottomed transit with Gaussian wings (model.py:332-517)."""
    device_kind = "transit"
    maxdur = np.inf

    def __init__(self, ts, amp=1, t0=None, paramranges=None):
        Model.__init__(self, ts, mtype='transit', amp=amp, t0=t0,
                       paramnames=['t0', 'sigmag', 'tauf', 'amp'], paramranges=paramranges)

    def set_max_duration(self, maxdur=np.inf):
        self.maxdur = maxdur

    def model(self, pdict, ts=None):
        self._need(pdict)
        if ts is None:
            ts = self.ts
        ts = np.asarray(ts)
        t0, amp, sg, tf = pdict['t0'], pdict['amp'], pdict['sigmag'], pdict['tauf']
        if t0 == np.inf:
            t0 = _centre(ts)
        f = -1. * amp * np.ones(len(ts))
        lo, hi = ts < t0 - tf, ts > t0 + tf
        if sg > 0:
            f[lo] = -1 * amp * np.exp(-(ts[lo] - (t0 - tf)) ** 2 / (2 * float(sg) ** 2))
            f[hi] = -1 * amp * np.exp(-(ts[hi] - (t0 + tf)) ** 2 / (2 * float(sg) ** 2))
        else:
            f[lo] = 0
            f[hi] = 0
        return f

    def prior(self, pdict, sigmacutoff=3.):
        """model.py:438-517.  The reference raises NameError here (undefined
        ``sigmagcutoff`` / ``taufranges``, SURVEY note N4); this is the evident intent with
        those two names corrected -- the same correction the oracle's reference loader applies."""
        self._need(pdict)
        sgr, tfr = self.ranges['sigmag'], self.ranges['tauf']
        if pdict['tauf'] + 2. * sigmacutoff * pdict['sigmag'] > self.maxdur:
            return -np.inf
        t0p, ampp = _flat_logprior(self.ranges['t0']), _flat_logprior(self.ranges['amp'])
        maxtf, mintf, maxsg, minsg = tfr[-1], tfr[0], sgr[-1], sgr[0]
        sg1 = (self.maxdur - mintf) / (2. * sigmacutoff)
        sg2 = (self.maxdur - maxtf) / (2. * sigmacutoff)
        tf1 = self.maxdur - (2. * sigmacutoff) * minsg
        tf2 = self.maxdur - (2. * sigmacutoff) * maxsg
        dsg, dtf = maxsg - minsg, maxtf - mintf
        lntf = 0.
        lnsg = 0.
        if sg1 < minsg and (minsg < sg2 <= maxsg):                       # triangle
            lnsg = -np.log((sg2 - minsg) * (tf1 - mintf) / 2.)
        elif (minsg < sg1 <= maxsg) and (minsg < sg2 <= maxsg):          # trapezium
            lnsg = -np.log((dtf / 2.) * (sg2 + sg1 - 2. * minsg))
        elif (mintf < tf1 < maxtf) and (mintf < tf2 < maxtf):            # trapezium
            lnsg = -np.log((dsg / 2.) * (tf2 + tf1 - 2. * mintf))
        elif (minsg < sg1 <= maxsg) and sg2 > maxsg:                     # pentagon
            lnsg = -np.log(dtf * (sg1 - minsg) + (maxsg - sg1) * (dtf + (tf2 - mintf)) / 2.)
        else:
            if len(sgr) > 1:
                lnsg = -np.log(sgr[-1] - sgr[0])
            if len(tfr) > 1:
                lntf = -np.log(tfr[-1] - tfr[0])
        return ampp + t0p + lntf + lnsg


def _transit_logprior_mesh(self, mesh):
    sg, tf = np.asarray(mesh['sigmag'], dtype=float), np.asarray(mesh['tauf'], dtype=float)
    bad = (tf + 2. * 3. * sg) > self.maxdur
    out = np.full(sg.shape, -np.inf)
    if not bad.all():
        i = np.argmin(bad.ravel())
        out[~bad] = self.prior({p: np.asarray(mesh[p]).ravel()[i] for p in self.paramnames})
    return out


Transit._logprior_mesh = _transit_logprior_mesh


class Expdecay(Model):
    """Pure exponential decay, or rise when reversed (model.py:520-652)."""
    device_kind = "expdecay"

    def __init__(self, ts, amp=1, t0=None, reverse=False, paramranges=None):
        Model.__init__(self, ts, mtype='expdecay', amp=amp, t0=t0, reverse=reverse,
                       paramnames=['t0', 'amp', 'tauexp'], paramranges=paramranges)

    def model(self, pdict, ts=None):
        self._need(pdict)
        if ts is None:
            ts = self.ts
        ts = np.asarray(ts)
        t0, amp, te = pdict['t0'], pdict['amp'], pdict['tauexp']
        if t0 == np.inf:
            t0 = _centre(ts)
        f = np.zeros(len(ts))
        f[ts == t0] = amp
        if te > 0:
            if self.reverse:
                sel = ts < t0
                f[sel] = amp * np.exp((ts[sel] - t0) / float(te))
            else:
                sel = ts > t0
                f[sel] = amp * np.exp(-(ts[sel] - t0) / float(te))
        return f

    def prior(self, pdict):
        self._need(pdict)
        return sum(_flat_logprior(self.ranges[p]) for p in ('t0', 'amp', 'tauexp'))


def _flat_logprior_mesh(self, mesh):
    first = next(iter(mesh.values()))
    return np.full(np.shape(first), self._flat_total())


Expdecay._logprior_mesh = _flat_logprior_mesh


class Impulse(Model):
    """Single-sample delta function (model.py:655-773); a finite ``t0`` is an offset
    from ``ts[0]``."""
    device_kind = "impulse"

    def __init__(self, ts, amp=1, t0=None, paramranges=None):
        Model.__init__(self, ts, mtype='impulse', amp=amp, t0=t0, paramnames=['t0', 'amp'],
                       paramranges=paramranges)

    def model(self, pdict, ts=None):
        self._need(pdict)
        if ts is None:
            ts = self.ts
        ts = np.asarray(ts)
        t0, amp = pdict['t0'], pdict['amp']
        t0 = _centre(ts) if t0 == np.inf else ts[0] + t0
        f = np.zeros_like(ts)
        f[np.abs(ts - t0).argmin()] = amp
        return f

    def prior(self, pdict):
        self._need(pdict)
        return _flat_logprior(self.ranges['t0']) + _flat_logprior(self.ranges['amp'])


class Gaussian(Model):
    """Gaussian profile (model.py:776-903)."""
    device_kind = "gaussian"

    def __init__(self, ts, amp=1, t0=None, paramranges=None):
        Model.__init__(self, ts, mtype='gaussian', amp=amp, t0=t0, paramnames=['t0', 'sigma', 'amp'],
                       paramranges=paramranges)

    def model(self, pdict, ts=None):
        self._need(pdict)
        if ts is None:
            ts = self.ts
        ts = np.asarray(ts)
        t0, amp, sigma = pdict['t0'], pdict['amp'], pdict['sigma']
        if t0 == np.inf:
            t0 = _centre(ts)
        if sigma == 0:
            f = np.zeros(len(ts))
            tm0 = ts - t0
            f[np.amin(tm0) == tm0] = amp      # as the reference: marks min(ts - t0)
            return f
        return amp * np.exp(-(ts - t0) ** 2 / (2 * float(sigma) ** 2))

    def prior(self, pdict):
        """The reference's version (model.py:860-903) calls ``len()`` on a float and always
        raises; this is the flat prior its docstring describes."""
        self._need(pdict)
        return sum(_flat_logprior(self.ranges[p]) for p in ('t0', 'amp', 'sigma'))


class Step(Model):
    """Step from zero to ``amp`` after ``t0`` (model.py:906-1019)."""
    device_kind = "step"

    def __init__(self, ts, amp=1, t0=None, reverse=False, paramranges=None):
        Model.__init__(self, ts, mtype='step', amp=amp, t0=t0, paramnames=['t0', 'amp'], paramranges=paramranges)

    def model(self, pdict, ts=None):
        self._need(pdict)
        if ts is None:
            ts = self.ts
        ts = np.asarray(ts)
        t0, amp = pdict['t0'], pdict['amp']
        if t0 == np.inf:
            t0 = _centre(ts)
        f = np.zeros_like(ts)
        f[ts > t0] = amp
        return f

    def prior(self, pdict):
        self._need(pdict)
        return _flat_logprior(self.ranges['t0']) + _flat_logprior(self.ranges['amp'])


Impulse._logprior_mesh = _flat_logprior_mesh
Step._logprior_mesh = _flat_logprior_mesh
Gaussian._logprior_mesh = _flat_logprior_mesh
