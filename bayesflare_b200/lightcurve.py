"""Light-curve container -- the duck type the hot path needs from ``bayesflare/data/data.py``:
fields ``cts, clc, cle, detrended, detrend_*`` and methods ``detrend, dt, fs, psd, dcoffset``
(used at ``bayes.py:128-137, 236-246, 259, 299, 312``).

FITS ingest (``Lightcurve.add_data``, data.py:205-269) is outside this repository's scope
(SURVEY.md 8(f) row f4); build curves from arrays, as the reference's own scripts do
(``scripts/flare_detection_efficiency.py:416-424``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["Lightcurve"]


class Lightcurve:
    id = 0
    dc = 0
    quarter = ""
    cadence = "long"
    datagap = False

    def __init__(self, curve=None, detrend=False, detrendmethod='savitzkygolay', nbins=101, order=3,
                 knee=(1. / (0.3 * 86400)), maxgap=1, cts=None, clc=None, cle=None):
        self.clc = np.array([])
        self.cts = np.array([])
        self.cle = np.array([])
        self.ulc = np.array([])
        self.detrended = False
        self.detrend_method = None
        self.detrend_nbins = 0
        self.detrend_order = 0
        self.detrend_knee = 0
        self.detrend_fit = np.array([])
        self.sinusoid_freqs = None
        if curve is not None:
            raise NotImplementedError("FITS ingest is out of scope for bayesflare_b200: pass cts/clc/cle arrays "
                                      "(or set the attributes) instead of a file name")
        if cts is not None:
            self.cts = np.array(cts, dtype=float)
            self.clc = np.array(clc, dtype=float)
            self.cle = np.zeros(len(self.cts)) if cle is None else np.array(cle, dtype=float)
            if detrend:
                self.detrend(detrendmethod, nbins, order, knee)

    def __str__(self):
        return "<bayesflare Lightcurve for KIC " + str(self.id) + ">"

    __repr__ = __str__

    def identity_string(self):
        return "lc_len_" + str(len(self.clc)) + "_cad_" + str(self.cadence)

    def dt(self):
        """data.py:162-171"""
        return self.cts[1] - self.cts[0]

    def fs(self):
        """data.py:173-182"""
        return 1.0 / self.dt()

    def psd(self):
        """One-sided, un-windowed, single-segment periodogram (data.py:184-203).  The reference
        calls ``matplotlib.mlab.psd(x, window=boxcar(l), noverlap=0, NFFT=l, Fs=fs)``; this is that
        estimator with modern-mlab scaling: ``|rfft(x)|^2 / (Fs * l)``, doubled except at DC and
        (even l) Nyquist."""
        x = np.asarray(self.clc, dtype=float)
        n = len(x)
        fs = self.fs()
        spec = np.abs(np.fft.rfft(x, n=n)) ** 2 / n
        if n % 2:
            spec[1:] *= 2.0
        else:
            spec[1:-1] *= 2.0
        return spec / fs, np.fft.rfftfreq(n, 1.0 / fs)

    def dcoffset(self):
        """Subtract the median (data.py:271-277)."""
        self.dc = np.median(self.clc)
        self.clc = self.clc - self.dc

    def set_detrend(self, method='none', nbins=None, order=None, knee=None):
        self.detrend_method = method
        self.detrend_length = nbins
        self.detrend_order = order
        self.detrend_knee = knee

    def detrend(self, method='none', nbins=None, order=None, knee=None):
        """data.py:377-424 (Savitzky-Golay; the running-median and Butterworth detrenders are
        not on the odds-ratio path)."""
        from .noise import savitzky_golay
        self.set_detrend(method=method, nbins=nbins, order=order, knee=knee)
        self.detrended = True
        self.ulc = np.copy(self.clc)
        if method == 'savitzkygolay':
            if nbins is None or order is None:
                raise ValueError("Number of bins, or polynomial order, for Savitsky-Golay filter not set")
            ffit = savitzky_golay(self.clc, nbins, order)
            self.clc = self.clc - ffit
            self.detrend_fit = np.copy(ffit)
        elif method in ('runningmedian', 'highpassfilter'):
            raise NotImplementedError("detrend method '%s' is outside the odds-ratio hot path" % method)
        else:
            raise ValueError("No detrend method set")
