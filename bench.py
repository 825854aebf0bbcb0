#!/usr/bin/env python
"""bench.py -- the headline benchmark: odds-ratio samples x grid-points per second.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port, host cores)

Workload (BASELINE.json config 3, per GPU): a batch of 1,024 simulated Kepler long-cadence light
curves (4,400 samples each, white noise + one injected flare per curve), the reference's default
OddsRatioDetector configuration (flare 10x10 grid + polynomial, impulse, exponential decay and rise
noise models; 93 grid points with finite prior, 108 nominal).  One step = one pass of the fused
detector over one batch.  Multi-GPU: one process per GPU, every rank owns its own batch (weak
scaling, no data-path collective); NCCL all-reduces only the per-step detection counts.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_EVAL, G_NOMINAL, N_TEMPLATES = 93, 108, 92
FLOP_PER_UNIT = 132.0                              # 2W + 2P + 12, constant-variance path
FLOP_SHARED_PER_SAMPLE = 575.0                     # data x background + forward substitution
THRESH = 16.5                                      # scripts/kepler_analysis_script.py:69-71


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curves", type=int, default=1024, help="light curves per GPU")
    ap.add_argument("--samples", type=int, default=4400, help="samples per light curve")
    ap.add_argument("--cpu-curves", type=int, default=64,
                    help="light curves in the CPU-baseline sample (the sample also stops after ~15 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""
    BITS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
            0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.BITS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.power:
            out["power_w_max"] = float(max(self.power))
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------
def cpu_oracle_throughput(cts, clc, max_seconds=None):
    """Time the CPU oracle (oracle/: C restatement of the reference's algorithm, OpenMP over grid
    points, all host cores) on the given curves.  Returns (units/s, seconds, lnO of curve 0)."""
    from oracle import oracle as orc
    orc.lib()
    t0 = time.perf_counter()
    first = None
    done = 0
    for b in range(clc.shape[0]):
        lnO, _ = orc.oddsratio(clc[b], np.zeros(clc.shape[1]), cts, noiseestmethod="powerspectrum", ignoreedges=False)
        if first is None:
            first = lnO
        done += 1
        if max_seconds is not None and time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return done * (clc.shape[1] - 54) * G_EVAL / dt, dt, first, done


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's algorithm on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from bayesflare_b200.synthetic import simulate_batch
    N = args.samples
    cts, clc, _ = simulate_batch(1, N, seed=3)
    # bounded sample: one step = one light curve of the workload, shortened if the host is slow
    ups, dt1, _, _ = cpu_oracle_throughput(cts[:1100], clc[:, :1100])
    n_step = N
    total_steps = args.steps + args.warmup
    while n_step > 256 and (n_step * G_EVAL / ups) * total_steps > 150.0:
        n_step //= 2
    cts_s, clc_s = cts[:n_step], clc[:, :n_step]
    for _ in range(args.warmup):
        cpu_oracle_throughput(cts_s, clc_s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_throughput(cts_s, clc_s)
    dt = time.perf_counter() - t0
    value = args.steps * (n_step - 54) * G_EVAL / dt
    cores = os.cpu_count()
    sample = "1 light curve x %d samples per step, default detector (G_eval=93), all samples incl. edges" % n_step
    line = {"impl": "reference", "metric": "odds-ratio samples x grid-points per second", "value": value,
            "unit": "samples*grid-points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: batch of simulated long-cadence light curves, default OddsRatioDetector "
                                   "(CPU arm: bounded sample)", "curves_per_gpu": args.curves, "samples": N,
                       "g_eval": G_EVAL, "g_nominal": G_NOMINAL},
            "cpu_baseline": {"value": value, "unit": "samples*grid-points/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "samples*grid-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import bayesflare_b200 as bf
    from bayesflare_b200 import _lib
    from bayesflare_b200.synthetic import simulate_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")           # both levels print NCCL's version banner on stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    B, N = args.curves, args.samples
    NB = 8                                        # rotating input batches: 8 x 36 MB > 126 MB L2

    # a dedicated (non-default) torch stream: libbfl enqueues on it and torch's CUDA events bracket it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    ctx = _lib.Context(local_rank, tstream.cuda_stream)

    # ---- inputs (outside the timed region) -------------------------------------------------
    cts, clc0, truth = simulate_batch(B, N, seed=100 + rank)
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc0[0]), device=local_rank)
    det.device = local_rank
    # the plan must live on the bench context (torch's stream) so CUDA events bracket the kernels
    from bayesflare_b200.bayes import background_models, hypothesis_from_model, window_times
    models = det.hypotheses(cts)
    mts = np.asarray(window_times(cts, det.bglen, models[0][0].t0), dtype=float)
    plan = _lib.Plan(ctx, det.bglen, background_models(det.bglen, det.bgorder), mts,
                     [hypothesis_from_model(m, mts, hr) for m, hr in models])
    g_eval = int(sum(np.isfinite(hypothesis_from_model(m, mts, hr)["logprior"]).sum() for m, hr in models)) + 1
    assert g_eval == G_EVAL, g_eval
    # noise sigma per curve (bayes.py:235-246), estimated on the device (bfl_noise_sigma); the
    # device-resident `value` loop takes it as an input, the e2e loop re-estimates it every step
    noise_spec = det.noise_spec()
    sk = _lib.noise_sigma(ctx, clc0, noise_spec)

    host = [torch.from_numpy(clc0).pin_memory()]
    for i in range(1, NB):
        host.append(torch.roll(host[0], shifts=(17 * i, 131 * i), dims=(0, 1)).contiguous().pin_memory())
    sk_rolled = [np.roll(sk, 17 * i) for i in range(NB)]
    d_in = [h.to(dev, non_blocking=True) for h in host]
    d_sk = [torch.from_numpy(s.copy()).to(dev) for s in sk_rolled]
    H = det.bglen // 2
    K0, K1 = H, N - H                            # ignoreedges=True (reference default): edges are dropped
    NK = K1 - K0
    d_out = torch.empty((B, NK), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int64, device=dev)
    d_tot = torch.zeros((args.steps + args.warmup + 1, 2), dtype=torch.int64, device=dev)
    d_red = torch.zeros_like(d_tot)              # all ranks' totals
    torch.cuda.synchronize()

    def step(i):
        j = i % NB
        plan.detector_odds_dev(d_in[j].data_ptr(), 0, True, d_sk[j].data_ptr(), B, N, 0, N, K0, K1, d_out.data_ptr())
        # per-curve region counts + this batch's totals (regions, curves with a detection) in row i
        ctx.count_regions_dev(d_out.data_ptr(), B, NK, THRESH, d_cnt.data_ptr(), d_tot[i].data_ptr())

    def drain():
        # NCCL over NVLink: detection counters only -- one all-reduce of the per-batch totals for the
        # whole job, issued on the kernels' stream inside the timed region
        d_red.copy_(d_tot)
        if world > 1:
            dist.all_reduce(d_red)

    peak_tf, _ = _lib.fp64_peak(ctx, 4096)        # roofline denominator, measured live (burst)

    for i in range(args.warmup):
        step(i)
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # rank 0's samples are the ones reported; the other ranks sample sparsely (NVML calls hold the GIL and
    # would otherwise compete with the launch loop on a box running 8 ranks)
    sampler = ClockSampler(physical_gpu_index(local_rank), period=0.004 if rank == 0 else 0.025)
    sampler.start()
    ctx.set_timing(True)
    ctx.window_time()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    drain()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    win_ms, win_n = ctx.window_time()
    ctx.set_timing(False)
    launches = ctx.launches - launches0
    clocks = sampler.stop()

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    units_per_step = float(B) * NK * G_EVAL       # samples actually produced x grid points with finite prior
    value = world * units_per_step * args.steps / (ms_max * 1e-3)

    # ---- e2e: the C-ABI call with HOST buffers (pinned), H2D + kernels + D2H every step ----
    out_host = torch.empty((B, NK), dtype=torch.float64).pin_memory()
    lib = ctx.lib
    import ctypes as C
    dp = C.POINTER(C.c_double)

    def e2e_step(i):
        # everything OddsRatioDetector.oddsratio does for a batch, from host memory to host memory:
        # H2D, noise sigma per curve (sk = NULL -> estimated on device), all hypotheses, D2H
        j = i % NB
        _lib.check(lib.bfl_detector_odds(ctx.handle, plan.handle, 1, C.cast(host[j].data_ptr(), dp), None,
                                         None, C.byref(noise_spec[0]), B, N, K0, K1,
                                         C.cast(out_host.data_ptr(), dp), None, None))
    for i in range(max(2, args.warmup // 2)):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, args.steps // 2)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 3)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_step * e2e_steps / float(t.item())

    # ---- sanity: the last e2e output against what the device path produced for that batch --
    lnO_host = out_host.numpy()
    n_det = int((d_cnt > 0).sum().item())
    job = d_red[args.warmup:args.warmup + args.steps].sum(dim=0).tolist()   # all ranks, timed steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_window_fast): FP64 FMA pipe --------------------
    interior = N - 2 * (det.bglen // 2)
    flops_launch = float(B) * interior * (N_TEMPLATES * FLOP_PER_UNIT + FLOP_SHARED_PER_SAMPLE)
    flops_strict = float(B) * interior * N_TEMPLATES * 2.0 * det.bglen
    # what the kernels really execute: the flare grid is held in separable form (rise x decay shapes), so
    # only `shapes` correlations of 2W flop are done per sample (k_corr_shapes), plus per template ~2 flop
    # to assemble z and ~34 flop of epilogue (E(z) = exp(phi(z)) from one 16-byte table row: node
    # derivatives, cubic in z - z_k, degree-5 series, one FMA into the linear-domain grid sum; k_epilogue)
    shapes = plan.shape_count if plan.shape_count > 0 else N_TEMPLATES
    flops_exec = float(B) * interior * (shapes * 2.0 * det.bglen + N_TEMPLATES * 36.0)
    win_avg_ms = win_ms / max(win_n, 1)
    achieved = flops_launch / (win_avg_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = float(B) * N * 8.0 + float(B) * NK * 8.0   # read the light curves, write the log odds
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_window_traffic_r1.json")))
        if B == 1024 and N == 4400:
            traffic = tj["traffic_bytes_per_launch"]      # dram read + write of one launch, ncu --set full
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "kernel": "window stage = k_corr_shapes<4> + k_epilogue<FUSED,2> (split form; both launches inside kernel_ms)", "kernel_ms": win_avg_ms,
                "kernel_share_of_step": win_ms / ms_total,
                "peak_source": "measured live by bfl_fp64_peak_probe (16 independent DFMA chains/thread, burst); "
                               "FP64 is not in MEASURED_PEAKS.json",
                "algorithmic_unit": "SURVEY.md 8(d): 132 flop per (sample, grid point) + 575 per sample "
                                    "(the reference's formulation of the work)",
                "achieved_dense_correlation_only": flops_strict / (win_avg_ms * 1e-3) / 1e12,
                "executed": {"tflops": flops_exec / (win_avg_ms * 1e-3) / 1e12, "frac": flops_exec / (win_avg_ms * 1e-3) / 1e12 / peak_tf,
                             "shapes_per_sample": shapes, "templates": N_TEMPLATES,
                             "note": "flops the kernel executes (separable flare grid: %d correlations per sample "
                                     "instead of %d) -- the algorithmic figure above counts the reference's work" % (shapes, N_TEMPLATES)},
                "flops_per_launch": flops_launch,
                "hbm": {"achieved": alg_bytes / (win_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}

    line = {"metric": "odds-ratio samples x grid-points per second", "value": value, "unit": "samples*grid-points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: batch of 1,024 simulated Kepler long-cadence light curves per GPU, "
                                   "default OddsRatioDetector (flare 10x10 + poly/impulse/expdecay/reverse noise models)",
                       "curves_per_gpu": B, "samples": N, "samples_out": NK, "ignoreedges": True, "bglen": det.bglen, "bgorder": det.bgorder,
                       "g_eval": G_EVAL, "g_nominal": G_NOMINAL, "threshold": THRESH,
                       "l2": "inputs rotate over %d distinct batches (%.0f MB) > 126 MB L2" % (NB, NB * B * N * 8 / 1e6),
                       "parallelism": "dp%d (one batch per rank, no data-path collective)" % world},
            "e2e": {"value": e2e_value, "unit": "samples*grid-points/s", "h2d_bytes_per_step": B * N * 8 + B * 8,
                    "d2h_bytes_per_step": B * NK * 8, "steps": e2e_steps,
                    "what": "bfl_detector_odds (C ABI) with pinned host buffers: H2D, noise sigma of every curve "
                            "estimated on device (SG detrend + periodogram), all hypotheses, D2H of the log odds"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "detections": {"curves_with_detection": n_det, "curves": B,
                           "job_regions_all_ranks": int(job[0]), "job_curves_with_detection_all_ranks": int(job[1])}}

    if not args.no_cpu_baseline and world == 1:
        nc = min(args.cpu_curves, B)
        ups, secs, ref0, done = cpu_oracle_throughput(cts, clc0[:nc], max_seconds=15.0)
        got0 = plan.detector_odds(clc0[0], None, sk[0], k0=K0, k1=K1)
        ref0 = ref0[K0:K1]
        err = float(np.max(np.abs(got0 - ref0) / np.maximum(1.0, np.abs(ref0))))
        line["cpu_baseline"] = {"value": ups, "unit": "samples*grid-points/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "%d light curves x %d samples of the same batch, %.1f s" % (done, N, secs)}
        line["parity_vs_oracle_curve0"] = err
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
