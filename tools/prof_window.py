"""Small driver for ncu: a few launches of the fused window kernel on a noise-only batch."""
import sys; sys.path.insert(0, '.')
import numpy as np
import bayesflare_b200 as bf
from bayesflare_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = 4400
rng = np.random.default_rng(0)
ts = np.arange(N) * 1765.55929
clc = rng.standard_normal((B, N))
det = bf.OddsRatioDetector(bf.Lightcurve(cts=ts, clc=clc[0]))
plan = det.plan()
sk = np.ones(B)
for it in range(3):
    lnO = plan.detector_odds(clc, None, sk)
print('ok', lnO.shape, float(np.nanmax(lnO)), plan.ctx.launches)
