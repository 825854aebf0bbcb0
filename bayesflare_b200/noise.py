"""Noise-level estimators -- host-side mirror of ``bayesflare/noise/noise.py``.

These produce the scalar noise standard deviation ``sk`` that feeds the window kernel
(``bayes.py:235-246``).  They stay on the host for now (SURVEY.md 8(f) row f1 lists the
on-device version as the next component); everything downstream of ``sk`` runs on the GPU.
"""
from __future__ import annotations

from math import factorial

import numpy as np

__all__ = ["running_median", "highpass_filter_lightcurve", "estimate_noise_ps", "estimate_noise_tv", "savitzky_golay", "nextpow2"]


def nextpow2(i):
    """misc/misc.py:3-20"""
    n = 1
    while n < i:
        n *= 2
    return n


def _sg_coefficients(window_size, order, deriv=0, rate=1):
    half = (window_size - 1) // 2
    b = np.array([[k ** i for i in range(order + 1)] for k in range(-half, half + 1)], dtype=float)
    return np.linalg.pinv(b)[deriv] * rate ** deriv * factorial(deriv)


def savitzky_golay(y, window_size, order, deriv=0, rate=1):
    """Savitzky-Golay smoothing with the reference's mirrored-difference edge padding
    (noise/noise.py:176-252)."""
    try:
        window_size = abs(int(window_size))
        order = abs(int(order))
    except ValueError:
        raise ValueError("window_size and order have to be of type int")
    if window_size % 2 != 1 or window_size < 1:
        raise TypeError("window_size size must be a positive odd number")
    if window_size < order + 2:
        raise TypeError("window_size is too small for the polynomials order")
    half = (window_size - 1) // 2
    m = _sg_coefficients(window_size, order, deriv, rate)
    y = np.asarray(y, dtype=float)
    firstvals = y[0] - np.abs(y[1:half + 1][::-1] - y[0])
    lastvals = y[-1] + np.abs(y[-half - 1:-1][::-1] - y[-1])
    y = np.concatenate((firstvals, y, lastvals))
    return np.convolve(m[::-1], y, mode='valid')


def estimate_noise_ps(lightcurve, estfrac=0.5):
    """Noise sigma from the mean of the upper ``estfrac`` of the one-sided periodogram
    (noise/noise.py:10-44).  Returns ``(sigma, variance, noise_v)``."""
    sk, f = lightcurve.psd()
    sk = np.mean(sk[int(np.floor((1. - estfrac) * len(sk))):])
    sk = sk * lightcurve.fs() / 2.
    noise_v = np.ones(nextpow2(2 * len(lightcurve.clc) - 1)) * sk
    return np.sqrt(sk), sk, noise_v


def estimate_noise_tv(d, sigma=1.0):
    """Tail-veto estimate: half-width of the central ``erf(sigma/sqrt2)`` probability mass of
    the data's histogram CDF (noise/noise.py:47-103).  Returns ``(std, centre)``."""
    from scipy.interpolate import interp1d
    from scipy.special import erf
    d = np.asarray(d, dtype=float)
    n, bins = np.histogram(d, bins=len(d), density=True)
    centres = (bins[:-1] + bins[1:]) / 2.
    cs = np.cumsum(n * (bins[1] - bins[0]))
    csu, idx = np.unique(cs, return_index=True)
    binsu = centres[idx]
    cp = erf(sigma / np.sqrt(2.))
    interpf = interp1d(csu, binsu)
    lowS = interpf(0.5 - cp / 2.)
    highS = interpf(0.5 + cp / 2.)
    m = interpf(0.5)
    return (highS - lowS) / (2. * sigma), m


def running_median(y, window):
    """Running median with a window that shrinks at the edges (noise.py:291-317): sample ``i`` uses
    the indices strictly between ``i - window/2`` and ``i + window/2``."""
    y = np.asarray(y, dtype=float)
    n = len(y)
    halfwin = int(window / 2)
    ffit = np.empty(n)
    for i in range(n):
        ffit[i] = np.median(y[max(0, i - halfwin + 1):min(n, i + halfwin)])
    return ffit


def highpass_filter_lightcurve(lightcurve, knee=(1. / (0.3 * 86400.))):
    """Third-order Butterworth high-pass applied to the time-reversed series followed by the series itself,
    keeping the second half (noise.py:254-288).  Returns the filtered array.  (The reference indexes with
    ``np.floor(len(y)/2)``, a float, which modern NumPy rejects; the integer meant is used.)"""
    from scipy import signal
    z = np.asarray(lightcurve.clc, dtype=float)
    dt = lightcurve.dt()
    if dt <= 0:
        raise NameError
    highcut = knee / (1. / (2. * dt))
    zd = np.concatenate((z[::-1], z))
    b, a = signal.butter(3, highcut, btype='highpass')
    y = signal.lfilter(b, a, zd)
    return y[len(y) // 2:]

