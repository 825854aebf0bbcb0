#!/usr/bin/env python
"""bench_configs.py -- the five BASELINE.json configurations, each measured through the public
API on one B200 and checked against the CPU oracle on a bounded sample (parity) or through a
size-independent property at full size.  One JSON object per configuration on stdout
(`bench.py` is the contract benchmark; this script is the per-configuration evidence kept under
`profiles/`).

    python bench_configs.py [1 2 3 4 5]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bayesflare_b200 as bf                      # noqa: E402
from bayesflare_b200 import _lib                  # noqa: E402
from bayesflare_b200.synthetic import DT_LONG, DT_SHORT, simulate_batch   # noqa: E402
from bayesflare_b200.parallel import time_chunks  # noqa: E402
from oracle import oracle as orc                  # noqa: E402

TOL = 1e-9


def scaled_err(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    return float(np.max(np.abs(got[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))))


def timeit(fn, reps=5, warm=1):
    for _ in range(warm):
        fn()
    _lib.default_context().sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    _lib.default_context().sync()
    return (time.perf_counter() - t0) / reps, out


def config1():
    """single simulated Kepler long-cadence quarter, default detector"""
    cts, clc, _ = simulate_batch(1, 4400, seed=11)
    lc = bf.Lightcurve(cts=cts, clc=clc[0])
    det = bf.OddsRatioDetector(lc)
    t_first = time.perf_counter()
    lnO, ts = det.oddsratio()                       # includes plan creation + host noise estimate
    t_first = time.perf_counter() - t_first
    t_api, (lnO, ts) = timeit(det.oddsratio, reps=10)
    sk = det.noise_sigma()
    plan = det.plan()
    t_abi, _ = timeit(lambda: plan.detector_odds(lc.clc, None, sk, k0=27, k1=4400 - 27), reps=20)
    t0 = time.perf_counter()
    ref, _ = orc.oddsratio(clc[0], np.zeros(4400), cts)
    t_cpu = time.perf_counter() - t0
    units = (4400 - 54) * det.ngrid_eval
    fl, n, mx = det.thresholder(lnO, 16.5, expand=1)
    rs, rn, rm = orc.thresholder(ref, 16.5, expand=1)
    return {"config": 1, "what": config1.__doc__, "N": 4400, "g_eval": det.ngrid_eval, "units": units,
            "api_first_call_s": t_first, "api_steady_s": t_api, "abi_s": t_abi,
            "units_per_s_api": units / t_api, "units_per_s_abi": units / t_abi,
            "cpu_oracle_s": t_cpu, "cpu_oracle_units_per_s": units / t_cpu, "cpu_cores": os.cpu_count(),
            "parity_max_scaled_err": scaled_err(lnO, ref), "detections_equal": [tuple(s) for s in fl] == [tuple(s) for s in rs],
            "n_detections": n}


def config2():
    """single short-cadence quarter (130k samples), flare vs noise/impulse/exp-decay odds, time-chunked with halos"""
    N = 130000
    cts, clc, truth = simulate_batch(1, N, dt=DT_SHORT, seed=22)
    clc = clc[0]
    # a few more injected flares so the series has detections along its length
    rng = np.random.default_rng(5)
    fl = bf.Flare(cts)
    for idx in rng.integers(1000, N - 1000, 12):
        lo, hi = idx - 200, idx + 900
        clc[lo:hi] += rng.uniform(4, 12) * fl.model({"t0": cts[idx], "amp": 1., "taugauss": rng.uniform(100, 400),
                                                     "tauexp": rng.uniform(400, 1500)}, ts=cts[lo:hi])
    lc = bf.Lightcurve(cts=cts, clc=clc)
    det = bf.OddsRatioDetector(lc)
    t_api, (lnO, ts) = timeit(det.oddsratio, reps=3)
    lnO = np.array(lnO)
    sk = det.noise_sigma()
    plan = det.plan()
    t_abi, whole = timeit(lambda: plan.detector_odds(clc, None, sk, k0=27, k1=N - 27), reps=10)
    # time chunks with halos through the device-resident entry point (what a multi-GPU split runs)
    import torch
    dev = torch.device("cuda:0")
    d_sk = torch.tensor([sk], dtype=torch.float64, device=dev)
    out = np.empty(N - 54)
    chunks = time_chunks(N, 8, 55, True)
    for k0, k1, s0, s1 in chunks:
        d_seg = torch.tensor(clc[s0:s1], dtype=torch.float64, device=dev)
        d_out = torch.empty(k1 - k0, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        plan.detector_odds_dev(d_seg.data_ptr(), 0, True, d_sk.data_ptr(), 1, N, s0, s1 - s0, k0, k1, d_out.data_ptr())
        plan.ctx.sync()
        out[k0 - 27:k1 - 27] = d_out.cpu().numpy()
    chunk_vs_whole = float(np.nanmax(np.abs(out - whole)))
    # parity on a bounded sample: the oracle on an 8,192-sample stretch with the same sigma
    a, b = 40000, 48192
    ref, _ = orc.oddsratio(clc[a:b], np.zeros(b - a), cts[a:b], sk=sk)
    # NB: templates are evaluated on the window time stamps at the series centre; a sub-series has its
    # own centre (last-bit differences in ts - t0), so the comparison uses a plan built on the sub-series
    det_sub = bf.OddsRatioDetector(bf.Lightcurve(cts=cts[a:b], clc=clc[a:b]))
    got = det_sub.plan().detector_odds(clc[a:b], None, sk, k0=27, k1=b - a - 27)
    units = (N - 54) * det.ngrid_eval
    segs, nseg, mx = det.thresholder(lnO, 16.5, expand=1)
    return {"config": 2, "what": config2.__doc__, "N": N, "g_eval": det.ngrid_eval, "units": units,
            "api_steady_s": t_api, "abi_s": t_abi, "units_per_s_api": units / t_api, "units_per_s_abi": units / t_abi,
            "chunks": len(chunks), "chunked_vs_whole_max_abs": chunk_vs_whole,
            "parity_sample": "8192-sample stretch", "parity_max_scaled_err": scaled_err(got, ref),
            "detections_above_16.5_equal": bool(np.array_equal(np.nonzero(got > 16.5)[0], np.nonzero(ref > 16.5)[0])),
            "n_detections": nseg}


def config4():
    """dense 64x64 tau_gauss x tau_exp flare grid + ParameterEstimationGrid posterior on an injected flare"""
    N = 4400
    rng = np.random.RandomState(44)
    cts = np.arange(N) * DT_LONG
    clc = rng.randn(N)
    t0i = 2000
    clc += bf.Flare(cts).model({"t0": cts[t0i], "amp": 20., "taugauss": 1060., "tauexp": 2168.})
    clc -= np.median(clc)
    lc = bf.Lightcurve(cts=cts, clc=clc)
    fp = {"taugauss": (0, 5400, 64), "tauexp": (1800, 10800, 64)}
    det = bf.OddsRatioDetector(lc, flareparams=fp)
    t_api, (lnO, ts) = timeit(det.oddsratio, reps=5)
    units = (N - 54) * det.ngrid_eval
    # parity of the dense grid on a bounded sample (N=512) against the oracle
    s = slice(1800, 2312)
    det_s = bf.OddsRatioDetector(bf.Lightcurve(cts=cts[s], clc=clc[s]), flareparams=fp, noiseestmethod="tailveto")
    got, _ = det_s.oddsratio()
    ref, _ = orc.oddsratio(clc[s], np.zeros(512), cts[s], flareparams=fp, noiseestmethod="tailveto")
    # property at full size: refining the grid changes the marginal only slightly and consistently
    det10 = bf.OddsRatioDetector(lc)
    lnO10, _ = det10.oddsratio()
    refine_shift = float(np.median(np.abs(np.array(lnO) - np.array(lnO10))))
    out = {"config": 4, "what": config4.__doc__, "N": N, "g_eval": det.ngrid_eval, "g_nominal": det.ngrid_nominal,
           "units": units, "api_steady_s": t_api, "units_per_s_api": units / t_api,
           "parity_sample": "512-sample stretch, 64x64 grid", "parity_max_scaled_err": scaled_err(got, ref),
           "median_abs_shift_vs_10x10_grid": refine_shift, "argmax_equal_10x10": int(np.argmax(lnO)) == int(np.argmax(lnO10))}
    # parameter-estimation grid, scripts/parameter_estimation_example.py at 64 x 64 x 40
    for bgorder in (3, 4):
        pe = bf.ParameterEstimationGrid('flare', lc)
        pe.lightcurve_chunk(t0i, 55)
        amprange = (0., 1.6 * (np.amax(pe.lightcurve.clc) - np.amin(pe.lightcurve.clc)), 40)
        pe.set_grid(ranges={'taugauss': (0., 3600., 64), 'tauexp': (0., 12000., 64), 'amp': amprange, 't0': (float(cts[t0i]),)})
        t_pe, _ = timeit(lambda: pe.calculate_posterior(bgorder=bgorder), reps=3)
        G = pe.posterior.size
        # parity of the WHOLE 64 x 64 x 40 grid against the oracle (the literal except_final recurrence)
        refp = orc.pe_posterior(pe.lightcurve.clc, pe.lightcurve.cts, "flare", pe.paramValues, pe.noiseSigma, bgorder=bgorder)
        same_inf = bool(np.array_equal(np.isfinite(pe.posterior), np.isfinite(refp)))
        fin = np.isfinite(refp) & np.isfinite(pe.posterior)
        pe_err = float(np.max(np.abs(pe.posterior[fin] - refp[fin]) / np.maximum(1.0, np.abs(refp[fin]))))
        mp, mpar = pe.maximum_posterior()
        out["pe_bgorder%d" % bgorder] = {"grid_points": int(G), "seconds": t_pe, "grid_points_per_s": G / t_pe,
                                         "max_posterior_params": {k: float(v) for k, v in mpar.items()},
                                         "pe_parity_max_scaled_err": pe_err, "pe_parity_points": int(fin.sum()),
                                         "pe_same_nonfinite_pattern": same_inf,
                                         "pe_argmax_equal": bool(int(np.argmax(pe.posterior)) == int(np.argmax(refp))),
                                         "finite_fraction": float(np.isfinite(pe.posterior).mean())}
        del refp
    return out


def config5(nreal=10000):
    """flare_detection_efficiency-style injection sweep: 10^4 realisations, scripts' detector (55-point impulse grid, tailveto)"""
    N = int(np.ceil(2893536. / DT_LONG))          # 1639 samples, scripts/flare_detection_efficiency.py:210-216
    t0 = time.perf_counter()
    cts, clc, truth = simulate_batch(nreal, N, seed=55, sinusoid_amp=3.0)
    t_sim = time.perf_counter() - t0
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0]), noiseestmethod="tailveto",
                               noiseimpulseparams={"t0": (0, 54 * DT_LONG, 55)})
    # host buffers in and out; the noise sigma of every realisation is estimated on the device (tail veto)
    t_abi, lnO = timeit(lambda: det.oddsratio_batch(clc), reps=3)
    units = nreal * (N - 54) * det.ngrid_eval
    # detection = any sample above threshold within +-2 samples of the injection (flare_detection_efficiency.py:560-600)
    thr = 16.5
    idx = truth["idx"] - 27
    hit = np.zeros(nreal, dtype=bool)
    for b in range(nreal):
        a, e = max(idx[b] - 2, 0), min(idx[b] + 3, lnO.shape[1])
        hit[b] = np.any(lnO[b, a:e] > thr)
    bins = np.linspace(0, 50, 51)
    tot, _ = np.histogram(truth["snr"], bins)
    det_h, _ = np.histogram(truth["snr"][hit], bins)
    # parity: four realisations against the oracle
    errs = []
    for b in (0, 1, 2, 3):
        ref, _ = orc.oddsratio(clc[b], np.zeros(N), cts, noiseestmethod="tailveto", noiseimpulseparams={"t0": (0, 54 * DT_LONG, 55)})
        errs.append(scaled_err(lnO[b], ref))
    # transit-model odds on one realisation (Bayes factor of a transit + background vs noise)
    tr = bf.Transit(cts, paramranges={"t0": (np.inf,), "sigmag": (1800., 10800., 4), "tauf": (1800., 14400., 4), "amp": (1.,)})
    B = bf.Bayes(bf.Lightcurve(cts=cts, clc=clc[0]), tr)
    t0 = time.perf_counter()
    B.bayes_factors_marg_poly_bgd(noiseestmethod="tailveto", halfrange=True)
    t_tr = time.perf_counter() - t0
    reft, _ = orc.bayes_factors(clc[0], np.zeros(N), cts, "transit",
                                {"t0": (np.inf,), "sigmag": (1800., 10800., 4), "tauf": (1800., 14400., 4), "amp": (1.,)},
                                det.noise_sigma(bf.Lightcurve(cts=cts, clc=clc[0])))
    # the whole sweep on the device: simulation (Philox), noise estimate, detector, matching, histograms
    from bayesflare_b200.sweep import detection_efficiency
    detection_efficiency(2048, N=N, seed=5, device_sim=True)          # warm-up (plan, buffers)
    t0 = time.perf_counter()
    sw = detection_efficiency(nreal, N=N, seed=5, device_sim=True)
    t_dev = time.perf_counter() - t0
    return {"config": 5, "what": config5.__doc__, "realisations": nreal, "N": N, "g_eval": det.ngrid_eval, "units": units,
            "simulate_host_s": t_sim, "abi_batch_s": t_abi, "units_per_s_abi": units / t_abi,
            "all_device_sweep_s": t_dev, "units_per_s_all_device_sweep": sw["units"] / t_dev,
            "all_device_efficiency_by_snr_decade": [float(sw["detected"][i:i + 10].sum() / max(sw["injected"][i:i + 10].sum(), 1))
                                                    for i in range(0, 50, 10)],
            "all_device_false_alarms": sw["false_alarms"],
            "efficiency_snr_gt_20": float(det_h[20:].sum() / max(tot[20:].sum(), 1)),
            "efficiency_by_snr_decade": [float(det_h[i:i + 10].sum() / max(tot[i:i + 10].sum(), 1)) for i in range(0, 50, 10)],
            "parity_sample": "4 realisations", "parity_max_scaled_err": max(errs),
            "transit_odds_s": t_tr, "transit_parity_max_scaled_err": scaled_err(B.lnBmargAmp, reft)}


def config6(B=512):
    """extra (not a BASELINE config): batch of long-cadence curves WITH per-sample errors, as real Kepler files
    carry (PDCSAP_FLUX_ERR) -- the weighted kernels; device noise estimate, host buffers in and out"""
    N = 4400
    cts, clc, _ = simulate_batch(B, N, seed=66)
    rng = np.random.default_rng(6)
    cle = 0.3 + 0.2 * rng.random((B, N)) + 0.05 * np.sin(np.arange(N) / 50.)[None, :]
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0], cle=cle[0]))
    ctx = det.plan().ctx
    t_api, (lnO, sk) = timeit(lambda: det.oddsratio_batch(clc, cle=cle, return_sk=True), reps=3)
    ctx.set_timing(True)
    det.oddsratio_batch(clc, cle=cle)
    ctx.sync()
    ms, nl = ctx.window_time()
    ctx.set_timing(False)
    units = B * (N - 54) * det.ngrid_eval
    errs = []
    for b in (0, 1):
        ref, _ = orc.oddsratio(clc[b], cle[b], cts, sk=float(sk[b]))
        errs.append(scaled_err(lnO[b], ref))
    return {"config": "6 (extra)", "what": config6.__doc__, "B": B, "N": N, "g_eval": det.ngrid_eval, "units": units,
            "api_batch_s": t_api, "units_per_s_api": units / t_api, "window_kernels_ms": ms, "window_kernel_launches": nl,
            "units_per_s_window_kernels": units / (ms * 1e-3), "parity_sample": "2 curves", "parity_max_scaled_err": max(errs)}


def config7(B=1024):
    """extra (not a BASELINE config): the config-3 batch with OTHER window lengths -- the split form is instantiated
    for bglen 21, 33, 55 and 101 (the lengths the reference's scripts use), anything else takes the single-kernel path"""
    N = 4400
    cts, clc, _ = simulate_batch(B, N, seed=100)
    out = {"config": "7 (extra)", "what": config7.__doc__, "B": B, "N": N, "by_bglen": {}}
    for bglen in (21, 33, 55, 101, 35):
        det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0]), bglen=bglen)
        ctx = det.plan().ctx
        sk = np.ones(B)
        det.oddsratio_batch(clc, sk=sk)
        ctx.set_timing(True)
        ctx.window_time()
        lnO = det.oddsratio_batch(clc, sk=sk)
        ms, nl = ctx.window_time()
        ctx.set_timing(False)
        units = B * (N - 2 * (bglen // 2)) * det.ngrid_eval
        ref, _ = orc.oddsratio(clc[1], np.zeros(N), cts, sk=1.0, bglen=bglen)
        out["by_bglen"][str(bglen)] = {"window_kernels_ms": ms, "launches": nl, "units_per_s_window_kernels": units / (ms * 1e-3),
                                       "shapes": det.plan().shape_count, "parity_max_scaled_err_curve1": scaled_err(lnO[1], ref)}
    return out


def main():
    which = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 5, 6, 7]
    fns = {1: config1, 2: config2, 4: config4, 5: config5, 6: config6, 7: config7}
    for w in which:
        if w == 3:
            print(json.dumps({"config": 3, "what": "see bench.py (the contract benchmark runs this configuration)"}))
            continue
        t0 = time.perf_counter()
        rec = fns[w]()
        rec["wall_s"] = time.perf_counter() - t0
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
