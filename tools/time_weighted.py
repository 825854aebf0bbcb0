import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import bayesflare_b200 as bf
from bayesflare_b200.synthetic import simulate_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 4400
cts, clc, _ = simulate_batch(B, N, seed=3)
rng = np.random.default_rng(0)
cle = 0.3 + 0.2 * rng.random((B, N))
det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0]))
plan = det.plan()
sk = np.full(B, 1.0)
def run():
    return plan.detector_odds(clc, cle, sk, k0=27, k1=N - 27)
out = run(); plan.ctx.sync()
plan.ctx.set_timing(True)
t0 = time.perf_counter()
for _ in range(3): out = run()
plan.ctx.sync()
dt = (time.perf_counter() - t0) / 3
ms, n = plan.ctx.window_time()
units = B * (N - 54) * det.ngrid_eval
# which kernels ran: BFL_NO_WSPLIT=1 -> k_window_wsep (one-kernel separable form), BFL_NO_WSEP=1 -> k_window_weighted (dense),
# BFL_NO_WEIGHTED=1 on top -> reference-order general path; default -> split form (k_wcorr + k_wsolve)
tag = ("general" if os.environ.get("BFL_NO_WEIGHTED") and os.environ.get("BFL_NO_WSEP") else
       "weighted" if os.environ.get("BFL_NO_WSEP") else "wsep" if os.environ.get("BFL_NO_WSPLIT") else "wsplit")
print(tag, "B", B, "s/call %.5f" % dt, "units/s %.3e" % (units / dt), "window kernel ms/call %.3f launches/call %d -> %.3e units/s" % (ms / 3, n // 3, units / (ms / 3e3) if ms else 0))
os.makedirs("gpurun_out", exist_ok=True); np.save("/tmp/wgt_%s_%d.npy" % (tag, B), out)
g = "/tmp/wgt_wsep_%d.npy" % B
if tag != "wsep" and os.path.exists(g):
    g = np.load(g)
    print("   max scaled diff vs the one-kernel form (wsep)", np.max(np.abs(g - out) / np.maximum(1, np.abs(g))))
