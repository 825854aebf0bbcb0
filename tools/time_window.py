"""Time the fused window kernel alone (CUDA events inside libbfl) for a noise batch."""
import sys; sys.path.insert(0, '.')
import numpy as np
import bayesflare_b200 as bf
from bayesflare_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = 4400
rng = np.random.default_rng(0)
ts = np.arange(N) * 1765.55929
clc = rng.standard_normal((B, N))
det = bf.OddsRatioDetector(bf.Lightcurve(cts=ts, clc=clc[0]))
plan = det.plan()
ctx = plan.ctx
sk = np.ones(B)
peak, _ = _lib.fp64_peak(ctx, 2048)
for it in range(2):
    lnO = plan.detector_odds(clc, None, sk, k0=27, k1=N-27)
ctx.set_timing(True); ctx.window_time()
for it in range(5):
    lnO = plan.detector_odds(clc, None, sk, k0=27, k1=N-27)
ms, n = ctx.window_time()
ms /= 5   # per call (the pipelined host entry launches the kernel once per sub-batch)
units = B*(N-54)*92
fl = B*(N-54)*(92*132+575)
print('KT env', __import__('os').environ.get('BFL_KT'), 'window ms %.3f  units/s %.3e  TF(survey) %.2f  frac %.3f  strict %.3f  peak %.2f' % (ms, units/ms*1e3, fl/ms/1e9, fl/ms/1e9/peak, B*(N-54)*92*110/ms/1e9/peak, peak), 'chk', float(np.nanmax(lnO)))
