#!/usr/bin/env python
"""Parity report: GPU path against the CPU oracle on the same inputs, next to the oracle's own sensitivity to
a one-ulp perturbation of the data (the reference algorithm's rounding-noise floor, SURVEY.md 7 hard part 1).
One JSON object per case on stdout (kept as profiles/parity_r1.jsonl).

    python tools/parity_report.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesflare_b200 as bf                                   # noqa: E402
from bayesflare_b200.synthetic import DT_LONG, simulate_batch   # noqa: E402
from oracle import oracle as orc                               # noqa: E402


def stats(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    fin = np.isfinite(ref) & np.isfinite(got)
    d = np.abs(got[fin] - ref[fin])
    sc = d / np.maximum(1.0, np.abs(ref[fin]))
    return {"max_abs": float(d.max()), "median_abs": float(np.median(d)), "max_scaled": float(sc.max()),
            "median_scaled": float(np.median(sc)), "same_nonfinite_pattern": bool(np.array_equal(np.isfinite(got), np.isfinite(ref)))}


def ulp_perturb(x, rng):
    return np.nextafter(x, np.where(rng.random(x.shape) < 0.5, -np.inf, np.inf))


def case(name, clc, cle, cts, sk, **kw):
    rng = np.random.default_rng(0)
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc, cle=cle), **kw)
    N = len(cts)
    h = 0 if kw.get("ignoreedges") is False else 27
    got = det.plan().detector_odds(clc, cle if np.any(cle != 0) else None, sk, k0=h, k1=N - h)
    ref, _ = orc.oddsratio(clc, cle, cts, sk=sk, **kw)
    ref2, _ = orc.oddsratio(ulp_perturb(clc, rng), cle, cts, sk=sk, **kw)
    rec = {"case": name, "N": N, "samples_compared": int(len(ref)), "gpu_vs_oracle": stats(got, ref),
           "oracle_one_ulp_data_perturbation": stats(ref2, ref),
           "detections_above_16.5_identical": bool(np.array_equal(np.nonzero(got > 16.5)[0], np.nonzero(ref > 16.5)[0]))}
    print(json.dumps(rec), flush=True)


def main():
    cts, clc, _ = simulate_batch(3, 4400, seed=11)
    z = np.zeros(4400)
    case("config 1: white noise + flare, interior samples (fast path)", clc[0], z, cts, 1.0)
    case("config 1 with the truncated edge windows (reference-order path)", clc[1], z, cts, 1.0, ignoreedges=False)
    _, col, _ = simulate_batch(1, 1639, seed=55, sinusoid_amp=3.0)
    case("config 5: sinusoidal trend, scripts' detector", col[0], np.zeros(1639), cts[:1639], 1.0,
         noiseimpulseparams={"t0": (0, 54 * DT_LONG, 55)})
    rng = np.random.default_rng(6)
    cle = 0.3 + 0.2 * rng.random(4400) + 0.05 * np.sin(np.arange(4400) / 50.)
    case("per-sample errors (weighted path)", clc[2], cle, cts, 1.0)
    big = clc[0] + 400.0 * np.sin(2 * np.pi * cts / (6.3 * 86400.))
    case("large slow trend, 400 sigma (fast path)", big, z, cts, 1.0)


if __name__ == "__main__":
    main()
