"""Time the pipelined batch search (bfl_pipe) from page-locked host buffers: ms per step, search mode and full mode.
    python tools/time_pipe.py [curves] [depth] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayesflare_b200 as bf
from bayesflare_b200 import _lib
from bayesflare_b200.bayes import background_models, hypothesis_from_model, window_times
from bayesflare_b200.synthetic import simulate_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
N = 4400
ctx = _lib.Context(0)
cts, clc, _ = simulate_batch(B, N, seed=100)
det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc[0]), device=0)
models = det.hypotheses(cts)
mts = np.asarray(window_times(cts, det.bglen, models[0][0].t0), dtype=float)
plan = _lib.Plan(ctx, det.bglen, background_models(det.bglen, det.bgorder), mts, [hypothesis_from_model(m, mts, hr) for m, hr in models])
K0, K1 = 27, N - 27
NK = K1 - K0
host_in = [torch.from_numpy(np.roll(clc, 17 * i, axis=0).copy()).pin_memory() for i in range(4)]
host_out = [torch.empty((B, NK), dtype=torch.float64).pin_memory() for _ in range(depth)]
host_cnt = [torch.zeros(B + 3, dtype=torch.int64).pin_memory() for _ in range(depth)]
host_reg = [torch.zeros(8 * B * 3, dtype=torch.float64).pin_memory() for _ in range(depth)]
pipe = _lib.Pipe(plan, B, N, K0, K1, noise=det.noise_spec(), depth=depth, max_regions=8 * B)
def run(n, full):
    jobs = []
    for i in range(n):
        s = i % depth
        if len(jobs) >= depth:
            pipe.wait(jobs[i - depth])
        jobs.append(pipe.submit(host_in[i % 4].data_ptr(), None, host_out[s].data_ptr() if full else None,
                                float("nan") if full else 16.5, None if full else host_reg[s].data_ptr(),
                                None if full else host_cnt[s].data_ptr()))
    for jb in jobs[-depth:]:
        pipe.wait(jb)
for full in (False, True):
    run(10, full)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps, full)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    # host-side cost of submit alone (no waiting): enqueue `depth` jobs into an idle pipe
    print("pipe B=%d depth=%d %s: %.4f ms/step, %.3e units/s, graphs %d, launches %d" % (B, depth, "full lnO" if full else "search", 1e3 * dt / steps, B * NK * 93 * steps / dt, pipe.graphs, pipe.launches), flush=True)
