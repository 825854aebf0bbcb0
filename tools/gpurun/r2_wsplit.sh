# weighted path, split form (bfl_wsplit.cu): parity tests, then timing against the one-kernel form
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "weighted or kepler or varying" > gpurun_out/wsplit_tests.log 2>&1; tail -3 gpurun_out/wsplit_tests.log
BFL_NO_WSPLIT=1 timeout 300 python tools/time_weighted.py 512 2>&1 | tail -1
timeout 300 python tools/time_weighted.py 512 2>&1 | tail -2
timeout 300 python tools/time_wsplit.py 256 1024 2>&1 | tail -2
python - <<'PY'
import numpy as np, time, torch, sys, os
sys.path.insert(0, os.getcwd())
import bayesflare_b200 as bf
from bayesflare_b200.synthetic import simulate_batch
B, N = 256, 4400
cts, clc, _ = simulate_batch(B, N, seed=3)
cle = 0.3 + 0.2 * np.random.default_rng(0).random((B, N))
det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0])); plan = det.plan()
d_c, d_e, d_s = (torch.from_numpy(x).cuda() for x in (clc, cle, np.ones(B)))
d_o = torch.empty((B, N - 54), dtype=torch.float64, device="cuda")
def step(): plan.detector_odds_dev(d_c.data_ptr(), d_e.data_ptr(), False, d_s.data_ptr(), B, N, 0, N, 27, N - 27, d_o.data_ptr())
step(); plan.ctx.sync()
t0 = time.perf_counter()
for _ in range(10): step()
plan.ctx.sync(); dt = (time.perf_counter() - t0) / 10
print("whole device-resident call B=256: %.3f ms -> %.3e units/s" % (dt * 1e3, B * (N - 54) * det.ngrid_eval / dt))
PY
