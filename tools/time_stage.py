"""Time the window stage (device-resident batch, CUDA events inside libbfl) and check it against the oracle.

    python tools/time_stage.py [curves] [samples] [tag]

Prints one line: per-launch time of the window stage (all its kernels), of the whole step (prep + window +
region count) measured with torch events, and the maximum scaled error of two curves against the CPU oracle.
With a tag, the log odds of the first 8 curves are saved to gpurun_out/stage_<tag>.npy so two builds /
environment settings can be compared bit by bit.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bayesflare_b200 as bf
from bayesflare_b200 import _lib
from bayesflare_b200.bayes import background_models, hypothesis_from_model, window_times
from bayesflare_b200.synthetic import simulate_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4400
tag = sys.argv[3] if len(sys.argv) > 3 else None
iters = int(os.environ.get("ITERS", "20"))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = _lib.Context(0, stream.cuda_stream)
cts, clc, _ = simulate_batch(B, N, seed=100)
det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0]), device=0)
models = det.hypotheses(cts)
mts = np.asarray(window_times(cts, det.bglen, models[0][0].t0), dtype=float)
plan = _lib.Plan(ctx, det.bglen, background_models(det.bglen, det.bgorder), mts,
                 [hypothesis_from_model(m, mts, hr) for m, hr in models])
sk = _lib.noise_sigma(ctx, clc, det.noise_spec())
NB = max(2, min(8, int(300e6 / (B * N * 8)) + 1))          # rotate inputs so they do not sit in L2
d_in = [torch.from_numpy(np.roll(clc, 17 * i, axis=0)).to(dev) for i in range(NB)]
d_sk = [torch.from_numpy(np.roll(sk, 17 * i).copy()).to(dev) for i in range(NB)]
H = det.bglen // 2
K0, K1 = H, N - H
NK = K1 - K0
d_out = torch.empty((B, NK), dtype=torch.float64, device=dev)
d_cnt = torch.zeros(B, dtype=torch.int64, device=dev)
d_tot = torch.zeros(2, dtype=torch.int64, device=dev)


def step(i):
    j = i % NB
    plan.detector_odds_dev(d_in[j].data_ptr(), 0, True, d_sk[j].data_ptr(), B, N, 0, N, K0, K1, d_out.data_ptr())
    ctx.count_regions_dev(d_out.data_ptr(), B, NK, 16.5, d_cnt.data_ptr(), d_tot.data_ptr())


for i in range(3):
    step(i)
torch.cuda.synchronize()
ctx.set_timing(True)
ctx.window_time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    step(i)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / iters
wms, wn = ctx.window_time()
ctx.set_timing(False)
step(0)
torch.cuda.synchronize()
out = d_out.cpu().numpy()
err = -1.0
if os.environ.get("NO_ORACLE") is None:
    from oracle import oracle as orc
    err = 0.0
    for b in (0, B - 1):
        ref, _ = orc.oddsratio(clc[b], np.zeros(N), cts, sk=float(sk[b]))
        err = max(err, float(np.max(np.abs(out[b] - ref) / np.maximum(1.0, np.abs(ref)))))
if tag:
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/stage_%s.npy" % tag, out[:8])
units = B * NK * 93
print("stage B=%d N=%d EPI=%s KT=%s: window %.4f ms/launch, step %.4f ms, %.3e units/s (window) %.3e (step), max err vs oracle %.2e"
      % (B, N, os.environ.get("BFL_EPI", "2"), os.environ.get("BFL_KT", "-"), wms / max(wn, 1), tot,
         units / (wms / max(wn, 1)) * 1e3, units / tot * 1e3, err), flush=True)
