"""Window-stage time of the weighted path (k_wcorr + k_wsolve by default) through bfl_detector_odds with host buffers, for
several batch sizes: CUDA events around the stage (bfl_ctx_set_timing), five calls each.
    python tools/time_wsplit.py 256 1024"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import bayesflare_b200 as bf
from bayesflare_b200.synthetic import simulate_batch
N = 4400
for B in [int(x) for x in sys.argv[1:]] or [256]:
    cts, clc, _ = simulate_batch(B, N, seed=3)
    rng = np.random.default_rng(0)
    cle = 0.3 + 0.2 * rng.random((B, N))
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0]))
    plan = det.plan()
    sk = np.full(B, 1.0)
    plan.detector_odds(clc, cle, sk, k0=27, k1=N - 27); plan.ctx.sync()
    plan.ctx.set_timing(True); plan.ctx.window_time()
    for _ in range(5): plan.detector_odds(clc, cle, sk, k0=27, k1=N - 27)
    plan.ctx.sync()
    ms, n = plan.ctx.window_time()
    units = B * (N - 54) * det.ngrid_eval
    print("B %d window stage %.3f ms/call (%d timed launches) -> %.3e units/s" % (B, ms / 5, n, units / (ms / 5e3)))
