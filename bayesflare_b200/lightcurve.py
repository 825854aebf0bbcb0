"""Light-curve container -- the duck type the hot path needs from ``bayesflare/data/data.py``:
fields ``cts, clc, cle, detrended, detrend_*`` and methods ``detrend, dt, fs, psd, dcoffset``
(used at ``bayes.py:128-137, 236-246, 259, 299, 312``).

Ingest: ``Lightcurve(curve=path)`` / ``add_data`` read a Kepler light-curve
FITS file (``data.py:205-269``) through the NumPy-only reader in :mod:`bayesflare_b200.fitsio` and apply
the reference's conditioning -- NaN interpolation with its float32 rounding (``data.py:337-351``), median
removal (``data.py:271-277``), gap flag (``data.py:279-303``).  Curves can also be built from arrays, as
the reference's own scripts do (``scripts/flare_detection_efficiency.py:416-424``).  Real files carry
``PDCSAP_FLUX_ERR``, i.e. per-sample variances: those curves run on the weighted kernels.
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
            # the reference passes a literal maxgap=1 here whatever the argument says (data.py:133)
            self.add_data(curve=curve, detrend=detrend, detrendmethod=detrendmethod, nbins=nbins, order=order,
                          knee=knee, maxgap=1)
        elif cts is not None:
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

    def add_data(self, curve=None, detrend=False, detrendmethod='none', nbins=101, order=3, knee=None, maxgap=1):
        """Append one Kepler light-curve file (data.py:205-269): ``PDCSAP_FLUX``, ``TIME`` (days -> s)
        and ``PDCSAP_FLUX_ERR`` of the first extension, then gap flag, NaN interpolation and median
        removal.  I/O problems raise ``NameError`` like the reference.  (The reference's own
        ``detrend=True`` branch calls ``self.detrend(nbins, order)`` with the arguments shifted by one
        and cannot run; here it detrends with the method given.)"""
        from .fitsio import read_kepler_lightcurve
        if curve is None:
            raise NameError("[Error] No light curve file given")
        try:
            d = read_kepler_lightcurve(curve)
        except (IOError, OSError, KeyError) as e:
            raise NameError("[Error] An IO error occured when trying to access " + str(curve) + " (" + str(e) + ")")
        if len(self.clc) == 0:
            self.id = d["KEPLERID"]
        elif d["KEPLERID"] != self.id:
            raise NameError("Tried to add data from KIC" + str(d["KEPLERID"]) + " to KIC" + str(self.id))
        self.cadence = 'long' if d["OBSMODE"] == 'long cadence' else 'short'
        self.quarter = str(self.quarter) + str(d["QUARTER"])
        # float32 columns widen to float64 on append, exactly as np.append does in the reference
        self.clc = np.append(self.clc, np.array(d["PDCSAP_FLUX"]))
        self.cts = np.append(self.cts, np.array(d["TIME"]) * 24 * 3600)
        self.cle = np.append(self.cle, np.array(d["PDCSAP_FLUX_ERR"]))
        self.datagap = self.gap_checker(self.clc, maxgap=maxgap)
        self.interpolate()
        self.dcoffset()
        if detrend:
            self.detrend(detrendmethod, nbins, order, knee)

    def gap_checker(self, d, maxgap=1):
        """data.py:279-303, literally: ``len(y < maxgap+1)`` is the length of a boolean array, which always
        equals ``len(y)``, so the reference never reports a gap; same here (same stars get analysed)."""
        z = np.invert(np.isnan(d))
        y = np.diff(z.nonzero()[0])
        if len(y < maxgap + 1) != len(y):
            return True
        return False

    def nan_helper(self, y):
        """data.py:305-335"""
        return np.isnan(y), lambda z: z.nonzero()[0]

    def interpolate(self):
        """Linear interpolation over NaN flux values; the reference rounds the interpolated values to
        float32 before storing them (data.py:337-351).  Errors are not interpolated: a NaN in ``cle``
        poisons the windows that contain it, as in the reference (log_marg_amp_full.c:20,25)."""
        z = self.clc
        nans, za = self.nan_helper(z)
        z[nans] = np.interp(za(nans), za(~nans), z[~nans]).astype('float32')
        self.clc = z

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
        """data.py:377-424: Savitzky-Golay (the odds-ratio path), running median, Butterworth high-pass."""
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
        elif method == 'runningmedian':
            if nbins is None:
                raise ValueError("Number of bins for running median filter not set")
            from .noise import running_median
            ffit = running_median(self.clc, nbins)
            self.clc = self.clc - ffit
            self.detrend_fit = np.copy(ffit)
        elif method == 'highpassfilter':
            if knee is None:
                raise ValueError("Knee frequency for high-pass filter not set.")
            from .noise import highpass_filter_lightcurve
            # the reference reads ``.clc`` of the returned ndarray here (data.py:421-422) and cannot run;
            # the filtered series is what it means
            self.clc = np.copy(highpass_filter_lightcurve(self, knee=knee))
        else:
            raise ValueError("No detrend method set")
