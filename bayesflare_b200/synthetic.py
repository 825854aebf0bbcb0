"""Synthetic Kepler-like light curves (no FITS data is available offline).

Follows the recipe the reference's own scripts use to simulate data
(``scripts/flare_detection_efficiency.py:262-289, 321-335, 416-434, 483-487``): white Gaussian noise
on the long- or short-cadence time base, an optional low-frequency sinusoid ("coloured" case),
one injected flare per curve with tau_gauss <= tau_exp and a uniform signal-to-noise ratio, then
median subtraction as ``Lightcurve.dcoffset`` does (data/data.py:271-277)."""
from __future__ import annotations

import numpy as np

from .models import Flare, Transit

DT_LONG = 1765.55929       # s
DT_SHORT = 58.84876        # s

__all__ = ["DT_LONG", "DT_SHORT", "simulate_batch"]


def script_injection_scale(base, inj, snr, bglen=55, bgorder=4):
    """The script's rescaling (flare_detection_efficiency.py:429-434, 483-487): the noise sigma is ESTIMATED from
    the base curve (tail veto of its Savitzky-Golay residual) and the injection gets the factor snr / injsnr with
    injsnr = (1 / sk) sqrt(sum inj^2).  Returns ``(factor, sk)``."""
    from .noise import estimate_noise_tv, savitzky_golay
    sk = estimate_noise_tv(base - savitzky_golay(base, bglen, bgorder), sigma=1.0)[0]
    return snr / ((1.0 / sk) * np.sqrt(np.sum(inj ** 2))), sk


def simulate_batch(B, N, dt=DT_LONG, seed=0, nstd=1.0, sinusoid_amp=0.0, inject=True, snr_range=(0., 50.), first=0,
                   recipe="nstd"):
    """Return ``(cts[N], clc[B, N], truth)``.  Curve ``i`` of the batch is realisation ``first + i`` and
    uses ``default_rng([seed, first + i])``, so a sweep sharded over ranks simulates the same curves
    whatever the number of ranks.

    ``recipe="nstd"`` (bench workloads): a flare scaled with the true noise level, median removed.
    ``recipe="script"``: the reference script to the letter -- ``inject`` is ``"flare"`` or ``"transit"`` (:291-335),
    the model is evaluated on the whole series and scaled with the sigma estimated from the base curve
    (:429-434, 483-487), no median removal; ``truth`` then also holds ``"sk"`` and, for transits, ``"sigmag"`` /
    ``"tauf"``."""
    if recipe == "script":
        return _simulate_script(B, N, dt, seed, nstd, sinusoid_amp, inject, snr_range, first)
    cts = np.arange(N, dtype=np.float64) * dt
    clc = np.empty((B, N))
    truth = {"idx": np.zeros(B, dtype=np.int64), "snr": np.zeros(B), "taugauss": np.zeros(B), "tauexp": np.zeros(B)}
    flare = Flare(cts, amp=1.)
    h = 27
    for i in range(B):
        rng = np.random.default_rng([int(seed), int(first) + i])
        x = nstd * rng.standard_normal(N)
        if sinusoid_amp > 0:
            amp = rng.uniform(0, sinusoid_amp) * nstd
            f = rng.uniform(1. / (31. * 86400.), 1. / (2. * 86400.))
            x += amp * np.sin(2 * np.pi * f * cts + rng.uniform(0, 2 * np.pi))
        if inject:
            idx = int(rng.integers(h, N - h - 1))
            te = rng.uniform(1800., 10800.)
            tg = rng.uniform(0., 5400.)
            while tg > te:
                tg = rng.uniform(0., 5400.)
            snr = rng.uniform(*snr_range)
            lo, hi = max(0, idx - 4 * h), min(N, idx + 12 * h)     # the template is negligible outside
            inj = flare.model({"t0": cts[idx], "amp": 1., "taugauss": tg, "tauexp": te}, ts=cts[lo:hi])
            x[lo:hi] += inj * (snr * nstd / np.sqrt(np.sum(inj ** 2)))
            truth["idx"][i], truth["snr"][i], truth["taugauss"][i], truth["tauexp"][i] = idx, snr, tg, te
        clc[i] = x - np.median(x)
    return cts, clc, truth


def _simulate_script(B, N, dt, seed, nstd, sinusoid_amp, inject, snr_range, first):
    kind = "flare" if inject is True else inject
    cts = np.arange(N, dtype=np.float64) * dt
    clc = np.empty((B, N))
    truth = {k: np.zeros(B) for k in ("snr", "taugauss", "tauexp", "sigmag", "tauf", "sk")}
    truth["idx"] = np.full(B, -1, dtype=np.int64)
    h = 27
    for i in range(B):
        rng = np.random.default_rng([int(seed), int(first) + i])
        x = nstd * rng.standard_normal(N)
        if sinusoid_amp > 0:
            amp = rng.uniform(0, sinusoid_amp) * nstd
            f = rng.uniform(1. / (31. * 86400.), 1. / (2. * 86400.))
            x += amp * np.sin(2 * np.pi * f * cts + rng.uniform(0, 2 * np.pi))
        if kind:
            idx = int(rng.integers(h, N - h - 1))
            snr = rng.uniform(*snr_range)
            if kind == "transit":
                tf, sg = rng.uniform(1800., 14400.), rng.uniform(0., 10800.)
                while tf + 4. * sg > 12. * 3600.:
                    tf, sg = rng.uniform(1800., 14400.), rng.uniform(0., 10800.)
                inj = Transit(cts, amp=1.).model({"t0": cts[idx], "amp": 1., "sigmag": sg, "tauf": tf})
                truth["sigmag"][i], truth["tauf"][i] = sg, tf
            else:
                te = rng.uniform(1800., 10800.)
                tg = rng.uniform(0., 5400.)
                while tg > te:
                    tg = rng.uniform(0., 5400.)
                inj = Flare(cts, amp=1.).model({"t0": cts[idx], "amp": 1., "taugauss": tg, "tauexp": te})
                truth["taugauss"][i], truth["tauexp"][i] = tg, te
            fac, sk = script_injection_scale(x, inj, snr)
            x = x + inj * fac
            truth["idx"][i], truth["snr"][i], truth["sk"][i] = idx, snr, sk
        clc[i] = x
    return cts, clc, truth
