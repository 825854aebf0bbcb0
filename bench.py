#!/usr/bin/env python
"""bench.py -- the headline benchmark: odds-ratio samples x grid-points per second.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference's own NumPy path

Workload (BASELINE.json config 3): ONE batch of 1,024 simulated Kepler long-cadence light curves per step
(4,400 samples each, white noise + one injected flare per curve), the reference's default OddsRatioDetector
configuration (flare 10x10 grid + polynomial, impulse, exponential decay and rise noise models; 93 grid points
with finite prior, 108 nominal).  One step = one pass of the detector over that batch.

Multi-GPU (north_star: ">= 7x at 8 GPUs on the 1,024-light-curve batch"): STRONG scaling.  One process per GPU;
the 1,024 curves of a step are split evenly over the ranks (128 per GPU at N = 8), no data-path collective;
NCCL all-reduces only the detection counters, once per job.  Consecutive steps rotate over four CUDA streams
(four libbfl contexts), so the ramp-up and tail of one step's kernels are filled by its neighbours': on 128
curves they are a quarter of a kernel's run.  Each step's launch sequence is a CUDA graph.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
METRIC = "odds-ratio samples x grid-points per second"
UNIT = "samples*grid-points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curves", type=int, default=1024, help="light curves per step (the whole job; split over the GPUs)")
    ap.add_argument("--samples", type=int, default=4400, help="samples per light curve")
    ap.add_argument("--cpu-curves", type=int, default=64,
                    help="light curves in the CPU-baseline sample of the oracle port (also stops after ~12 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained-clock block")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling figure at N > 1")
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams (libbfl contexts) the steps alternate between")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons / power of one GPU with NVML while a timed region runs."""
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
            out["power_w_median"] = float(np.median(self.power))
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
# CPU arms
def set_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arms use every host core and say how many."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def cpu_oracle_throughput(cts, clc, max_seconds=None):
    """Time the CPU oracle (oracle/: C restatement of the reference's algorithm, OpenMP over grid
    points) on the given curves.  Returns (units/s, seconds, lnO of curve 0, curves done)."""
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


def reference_available():
    try:
        from oracle import ref_loader
        return ref_loader.available()
    except Exception:
        return False


def reference_oddsratio(cts, clc):
    """The reference's own path: ``OddsRatioDetector(lc).oddsratio()`` of the unmodified reference package
    (oracle/_ref/site, Python-3 patches applied in memory by oracle/ref_loader.py), ncpus=None = one worker
    process per host core (bayes.py:421-429).  Returns (lnO, seconds)."""
    import contextlib
    import io
    from oracle import ref_loader
    bf = ref_loader.load_reference()
    lc = ref_loader.RefLightcurve(cts, clc)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        det = bf.OddsRatioDetector(lc)
        lnO, _ = det.oddsratio()
    return np.asarray(lnO), time.perf_counter() - t0


def run_reference(args, rank, world):
    """CPU arm (rank 0 only): the reference's own NumPy/Cython implementation when oracle/_ref holds it (kind
    "reference"), else the oracle port (kind "port"); every host core; a bounded sample per step."""
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    from bayesflare_b200.synthetic import simulate_batch
    cores = set_host_threads()
    N = args.samples
    cts, clc, _ = simulate_batch(1, N, seed=100)
    total_steps = args.steps + args.warmup
    use_ref = reference_available()
    if use_ref:
        _, dt0 = reference_oddsratio(cts[:400], clc[0, :400])          # also pays the import
        _, dt1 = reference_oddsratio(cts[:700], clc[0, :700])
        per_sample = max(dt1 / 700.0, 1e-6)
        n_step = int(min(N, max(300, 140.0 / total_steps / per_sample)))
        kind = "reference"

        def one_step():
            reference_oddsratio(cts[:n_step], clc[0, :n_step])
        units = (n_step - 54) * G_EVAL          # find.py:846-851: the reference returns the interior samples
        what = "the reference's own OddsRatioDetector.oddsratio() (unmodified package from oracle/_ref/site, ncpus=None)"
    else:
        ups, _, _, _ = cpu_oracle_throughput(cts[:1100], clc[:, :1100])
        n_step = N
        while n_step > 256 and (n_step * G_EVAL / ups) * total_steps > 150.0:
            n_step //= 2
        kind = "port"

        def one_step():
            cpu_oracle_throughput(cts[:n_step], clc[:, :n_step])
        units = (n_step - 54) * G_EVAL
        what = "oracle port (C/OpenMP restatement of the reference's algorithm)"
    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    value = args.steps * units / dt
    sample = "1 light curve x %d samples per step, default detector (G_eval=93); %s" % (n_step, what)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: batch of 1,024 simulated long-cadence light curves, default OddsRatioDetector "
                                   "(CPU arm: bounded sample of it per step)", "curves": args.curves, "samples": N,
                       "g_eval": G_EVAL, "g_nominal": G_NOMINAL},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "threads": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if use_ref:      # the port beside it, same host, for scale
        ups, secs, _, done = cpu_oracle_throughput(cts, clc, max_seconds=8.0)
        line["cpu_baseline_port"] = {"value": ups, "unit": UNIT, "cores": cores, "threads": cores, "kind": "port",
                                     "sample": "%d light curve x %d samples, %.1f s" % (done, N, secs)}
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
    if args.curves % world:
        raise SystemExit("bench.py: --curves must be a multiple of the number of GPUs")
    B, N = args.curves // world, args.samples     # strong scaling: this rank's share of the batch

    # (non-default) torch streams, one libbfl context on each: consecutive steps rotate over them
    nstreams = max(1, min(4, args.streams))
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    torch.cuda.set_stream(streams[0])
    ctxs = [_lib.Context(local_rank, s.cuda_stream) for s in streams]

    # ---- inputs (outside the timed region) -------------------------------------------------
    # rank r owns curves [r B, (r + 1) B) of the batch; realisation i is the same whatever the number of ranks
    cts, clc0, truth = simulate_batch(B, N, seed=100, first=rank * B)
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=cts, clc=clc0[0]), device=local_rank)
    det.device = local_rank
    from bayesflare_b200.bayes import background_models, hypothesis_from_model, window_times
    models = det.hypotheses(cts)
    mts = np.asarray(window_times(cts, det.bglen, models[0][0].t0), dtype=float)
    hyps = [hypothesis_from_model(m, mts, hr) for m, hr in models]
    plan = _lib.Plan(ctxs[0], det.bglen, background_models(det.bglen, det.bgorder), mts, hyps)
    g_eval = int(sum(np.isfinite(h["logprior"]).sum() for h in hyps)) + 1
    assert g_eval == G_EVAL, g_eval
    # noise sigma per curve (bayes.py:235-246), estimated on the device (bfl_noise_sigma); the
    # device-resident `value` loop takes it as an input, the e2e loop re-estimates it every step
    noise_spec = det.noise_spec()
    sk = _lib.noise_sigma(ctxs[0], clc0, noise_spec)

    # rotating input batches: together larger than the 126 MB L2
    NB = max(4, int(300e6 / (B * N * 8)) + 1)
    NB += (-NB) % nstreams                       # step i uses batch i % NB on stream i % nstreams: one CUDA graph per batch
    base = torch.from_numpy(clc0).to(dev)
    d_in = [base] + [torch.roll(base, shifts=(17 * i, 131 * i), dims=(0, 1)).contiguous() for i in range(1, NB)]
    d_sk = [torch.from_numpy(np.roll(sk, 17 * i).copy()).to(dev) for i in range(NB)]
    H = det.bglen // 2
    K0, K1 = H, N - H                            # ignoreedges=True (reference default): edges are dropped
    NK = K1 - K0
    d_out = [torch.empty((B, NK), dtype=torch.float64, device=dev) for _ in range(nstreams)]
    d_cnt = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(nstreams)]
    d_tot = torch.zeros((4, 3), dtype=torch.int64, device=dev)   # rows: 0 timed job, 1 warm-up, 2 other loops, 3 scratch
    d_red = torch.zeros(3, dtype=torch.int64, device=dev)        # all ranks' totals of the timed job
    torch.cuda.synchronize()

    def launch_step(j, w, row):
        plan.detector_odds_dev(d_in[j].data_ptr(), 0, True, d_sk[j].data_ptr(), B, N, 0, N, K0, K1, d_out[w].data_ptr(),
                               ctx=ctxs[w])
        # per-curve region counts; the batch's totals (regions, curves with a detection) accumulate in d_tot[row]
        ctxs[w].count_regions_dev(d_out[w].data_ptr(), B, NK, THRESH, d_cnt[w].data_ptr(), d_tot[row].data_ptr())

    # The launch sequence of one step (k_prep_scalars, k_corr_shapes, k_epilogue2, memset, k_count_regions) is
    # captured once per input batch into a CUDA graph on the stream it runs on: at 128 curves per GPU a step is
    # 0.2 ms of GPU time, and 8 ranks' Python launch loops on one host do not keep 8 GPUs fed otherwise.
    for j in range(NB):                              # un-captured first: scratch is sized, attributes are set
        launch_step(j, j % nstreams, 3)
    torch.cuda.synchronize()
    graphs = {}
    launches_per_step = None
    for row in (0, 1, 2):
        for j in range(NB):
            w = j % nstreams
            l0 = ctxs[w].launches
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[w], capture_error_mode="thread_local"):
                launch_step(j, w, row)
            launches_per_step = ctxs[w].launches - l0
            graphs[(row, j)] = g
    torch.cuda.set_stream(streams[0])
    torch.cuda.synchronize()

    def step(i, row=0):
        j = i % NB
        with torch.cuda.stream(streams[j % nstreams]):
            graphs[(row, j)].replay()

    ev_join = [torch.cuda.Event() for _ in range(nstreams)]

    def join_streams():
        """stream 0 waits for everything enqueued on the other stream(s)"""
        for w in range(1, nstreams):
            ev_join[w].record(streams[w])
            streams[0].wait_event(ev_join[w])

    def fork_streams():
        """the other stream(s) wait for everything enqueued on stream 0 (the start event)"""
        ev_join[0].record(streams[0])
        for w in range(1, nstreams):
            streams[w].wait_event(ev_join[0])

    def drain():
        # NCCL over NVLink: detection counters only -- ONE all-reduce of the per-step totals for the whole
        # job, issued on stream 0 inside the timed region after both streams' kernels
        join_streams()
        d_red.copy_(d_tot[0])
        if world > 1:
            dist.all_reduce(d_red)

    peak_tf, _ = _lib.fp64_peak(ctxs[0], 4096)    # roofline denominator, measured live (burst)

    # rank 0's samples are the ones reported; the other ranks sample sparsely (NVML calls of 8 processes
    # serialise in the driver)
    sampler = ClockSampler(physical_gpu_index(local_rank), period=0.002 if rank == 0 else 0.05)
    sampler.start()
    align = torch.zeros(1, dtype=torch.int64, device=dev)
    for i in range(args.warmup):
        step(i, row=1)
    drain()

    def timed_region():
        """EXACTLY K steps between a barrier + synchronize on both sides; device time, max over ranks"""
        d_tot[0].zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            # Host threads leave the barrier up to milliseconds apart (8 Python processes on one box); a 20-step
            # region is only ~4 ms long and ends in a collective that waits for the last rank.  So the ranks are
            # aligned once more ON THE DEVICE: an all-reduce enqueued right in front of the start event -- every
            # rank's region starts when the last rank's GPU gets there, with the K steps already queued behind it.
            dist.all_reduce(align)
        e0.record(streams[0])
        fork_streams()
        for i in range(args.steps):
            step(args.warmup + i)
        drain()
        e1.record(streams[0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # The K-step region is a few milliseconds long (4 ms at 8 GPUs): one scheduling hiccup on one rank is a tenth of
    # it.  It is therefore measured three times, each a complete barrier-to-barrier measurement of exactly K steps,
    # and the MEDIAN is reported; all three are in the line (`ms_per_step_repeats`).
    repeats = sorted(timed_region() for _ in range(3))
    ms_total = repeats[1]
    launches = launches_per_step * args.steps       # kernels of libbfl inside the timed region (graph replays)
    clocks = sampler.stop()
    ms_max = ms_total
    units_per_step = float(B) * NK * G_EVAL       # this rank: samples actually produced x grid points with finite prior
    value = world * units_per_step * args.steps / (ms_max * 1e-3)
    job = d_red.tolist()                          # all ranks, timed steps
    n_det = int((d_cnt[(args.warmup + args.steps - 1) % nstreams] > 0).sum().item())

    # ---- the collective alone: the job's one all-reduce of the detection counters -----------
    coll_us = None
    if world > 1:
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dist.all_reduce(d_red)
        torch.cuda.synchronize()
        dist.barrier()
        ce0.record(streams[0])
        for _ in range(20):
            dist.all_reduce(d_red)
        ce1.record(streams[0])
        torch.cuda.synchronize()
        coll_us = 1e3 * ce0.elapsed_time(ce1) / 20.0

    # ---- window-stage kernels alone, one stream, CUDA events around every launch (roofline numerator) ----
    ctxs[0].set_timing(True)
    ctxs[0].window_time()
    ksteps = min(args.steps, 20)
    ke0, ke1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ke0.record(streams[0])
    for i in range(ksteps):
        launch_step(i % NB, 0, 3)
    ke1.record(streams[0])
    torch.cuda.synchronize()
    win_ms, win_n = ctxs[0].window_time()
    ctxs[0].set_timing(False)
    serial_ms = ke0.elapsed_time(ke1) / ksteps
    win_avg_ms = win_ms / max(win_n, 1)

    # ---- sustained: the same step loop for >= 2 s with the clock sampler on; the FP64 probe likewise ----
    sustained = None
    if not args.no_sustained:
        nsus = int(max(args.steps, 2200.0 / max(ms_max / args.steps, 1e-3)))
        if world > 1:
            dist.barrier()
        sm2 = ClockSampler(physical_gpu_index(local_rank), period=0.01)
        sm2.start()
        se0, se1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        se0.record(streams[0])
        fork_streams()
        for i in range(nsus):
            step(i, row=2)
        join_streams()
        se1.record(streams[0])
        torch.cuda.synchronize()
        sus_ms = se0.elapsed_time(se1)
        c_run = sm2.stop()
        t = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms_max = float(t.item())
        sus_value = world * units_per_step * nsus / (sus_ms_max * 1e-3)
        sm3 = ClockSampler(physical_gpu_index(local_rank), period=0.01)
        sm3.start()
        probes, t0 = [], time.perf_counter()
        while time.perf_counter() - t0 < 2.0:
            probes.append(_lib.fp64_peak(ctxs[0], 4096)[0])
        c_probe = sm3.stop()
        peak_sus = float(np.median(probes[len(probes) // 2:]))
        sustained = {"value": sus_value, "unit": UNIT, "seconds": sus_ms_max * 1e-3, "steps": nsus,
                     "ms_per_step": sus_ms_max / nsus, "vs_burst": sus_value / value, "clocks": c_run,
                     "fp64_peak_sustained_tflops": peak_sus, "fp64_peak_burst_tflops": peak_tf, "probe_clocks": c_probe}

    # ---- e2e: host buffers -> host buffers through the C ABI's pipelined batch search (bfl_pipe) ----
    # every step: H2D of the step's light curves from page-locked memory, noise sigma of every curve
    # estimated on device, all hypotheses, and D2H of the step's result.  Six jobs are in flight, so the
    # copies of one step overlap the kernels of its neighbours; every step's copies are inside the region.
    NBH = 4
    host_in = [torch.from_numpy(clc0).pin_memory()]
    for i in range(1, NBH):
        host_in.append(torch.roll(host_in[0], shifts=(17 * i, 131 * i), dims=(0, 1)).contiguous().pin_memory())
    depth = 6
    MAXREG = 8 * B
    host_out = [torch.empty((B, NK), dtype=torch.float64).pin_memory() for _ in range(depth)]
    host_cnt = [torch.zeros(B + 3, dtype=torch.int64).pin_memory() for _ in range(depth)]
    host_reg = [torch.zeros(MAXREG * 3, dtype=torch.float64).pin_memory() for _ in range(depth)]   # 24-byte records
    pipe = _lib.Pipe(plan, B, N, K0, K1, noise=noise_spec, depth=depth, max_regions=MAXREG)

    def e2e_run(nsteps, full):
        """`full`: the log odds come back (what oddsratio() returns); else only the above-threshold regions
        (what a search keeps: thresholder's segments with their maxima)."""
        jobs = []
        for i in range(nsteps):
            s = i % depth
            if len(jobs) >= depth:
                pipe.wait(jobs[i - depth])          # the slot's host buffers are free again
            jobs.append(pipe.submit(host_in[i % NBH].data_ptr(), None,
                                    host_out[s].data_ptr() if full else None,
                                    float("nan") if full else THRESH,
                                    None if full else host_reg[s].data_ptr(),
                                    None if full else host_cnt[s].data_ptr()))
        nreg = 0
        for jb in jobs[-depth:]:
            nreg = pipe.wait(jb)
        return nreg

    def e2e_time(full):
        """median of 3 complete wall-clock measurements of exactly K jobs (barrier + synchronize before, all jobs waited for)"""
        e2e_run(max(3, args.warmup), full)
        vals, nl = [], 0
        for _ in range(3):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            l0 = pipe.launches
            t0 = time.perf_counter()
            e2e_run(args.steps, full)
            dt = time.perf_counter() - t0
            nl = pipe.launches - l0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vals.append(world * units_per_step * args.steps / float(t.item()))
        return sorted(vals)[1], nl

    e2e_value, _ = e2e_time(True)
    e2e_search_value, e2e_launches = e2e_time(False)
    last_cnt = host_cnt[(args.steps - 1) % depth].numpy().copy()
    pipe.wait(pipe.submit(host_in[0].data_ptr(), None, host_out[0].data_ptr(), float("nan"), None, None))
    lnO_host = host_out[0].numpy().copy()         # the unrolled batch once more, for the parity line

    # ---- secondary: weak scaling (1,024 curves on EVERY GPU), device-resident, a few steps ----
    weak = None
    if world > 1 and not args.no_weak:
        Bw = args.curves
        _, clcw, _ = simulate_batch(Bw, N, seed=100, first=rank * Bw)
        skw = torch.from_numpy(_lib.noise_sigma(ctxs[0], clcw, noise_spec)).to(dev)
        dw = [torch.from_numpy(clcw).to(dev)]
        dw += [torch.roll(dw[0], shifts=(17 * i, 131 * i), dims=(0, 1)).contiguous() for i in range(1, 8)]
        ow = [torch.empty((Bw, NK), dtype=torch.float64, device=dev) for _ in range(nstreams)]

        def wstep(i):
            w = i % nstreams
            plan.detector_odds_dev(dw[i % 8].data_ptr(), 0, True, skw.data_ptr(), Bw, N, 0, N, K0, K1, ow[w].data_ptr(),
                                   ctx=ctxs[w])
        for i in range(3):
            wstep(i)
        torch.cuda.synchronize()
        dist.barrier()
        we0, we1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        we0.record(streams[0])
        fork_streams()
        for i in range(10):
            wstep(i)
        join_streams()
        we1.record(streams[0])
        torch.cuda.synchronize()
        t = torch.tensor([we0.elapsed_time(we1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        weak = {"value": world * float(Bw) * NK * G_EVAL * 10 / (float(t.item()) * 1e-3), "unit": UNIT,
                "curves_per_gpu": Bw, "steps": 10, "scaling": "weak"}
        del dw, ow

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the window stage (k_corr_shapes + k_epilogue2): FP64 FMA pipe ----------
    interior = N - 2 * (det.bglen // 2)
    flops_launch = float(B) * interior * (N_TEMPLATES * FLOP_PER_UNIT + FLOP_SHARED_PER_SAMPLE)
    flops_strict = float(B) * interior * N_TEMPLATES * 2.0 * det.bglen
    # what the kernels really execute: the flare grid is held in separable form (rise x decay shapes), so only
    # `shapes` correlations of 2W flop are done per sample (k_corr_shapes), plus per template 26 FP64
    # instructions (~50 flop: index, reciprocal refinement, derivatives of phi, two short series; k_epilogue2)
    shapes = plan.shape_count if plan.shape_count > 0 else N_TEMPLATES
    flops_exec = float(B) * interior * (shapes * 2.0 * det.bglen + N_TEMPLATES * 50.0)
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
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_window_traffic_r2.json")))
        traffic = tj.get("traffic_bytes_per_launch", {}).get(str(B))      # dram read + write of one launch, ncu --set full
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic,
                "kernel": "window stage = k_corr_shapes<4> + k_epilogue2 (split form; both launches inside kernel_ms)",
                "kernel_ms": win_avg_ms, "kernel_share_of_step": win_avg_ms / serial_ms,
                "how": "CUDA events around every launch of the stage on its own stream, %d steps back to back on ONE stream "
                       "after the timed region (the timed region alternates two streams, whose launches overlap)" % ksteps,
                "serial_ms_per_step": serial_ms,
                "peak_source": "measured live by bfl_fp64_peak_probe (16 independent DFMA chains/thread, burst); "
                               "FP64 is not in MEASURED_PEAKS.json",
                "algorithmic_unit": "SURVEY.md 8(d): 132 flop per (sample, grid point) + 575 per sample "
                                    "(the reference's formulation of the work)",
                "achieved_dense_correlation_only": flops_strict / (win_avg_ms * 1e-3) / 1e12,
                "executed": {"tflops": flops_exec / (win_avg_ms * 1e-3) / 1e12,
                             "frac": flops_exec / (win_avg_ms * 1e-3) / 1e12 / peak_tf,
                             "shapes_per_sample": shapes, "templates": N_TEMPLATES,
                             "note": "flops the kernels execute (separable flare grid: %d correlations per sample "
                                     "instead of %d) -- the algorithmic figure above counts the reference's work" % (shapes, N_TEMPLATES)},
                "flops_per_launch": flops_launch,
                "hbm": {"achieved": alg_bytes / (win_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
    if sustained is not None:
        sustained["frac_of_sustained_peak"] = (world * flops_launch / (sustained["ms_per_step"] * 1e-3) / 1e12 /
                                               (world * sustained["fp64_peak_sustained_tflops"]))

    line = {"metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "ms_per_step_repeats": [r / args.steps for r in repeats],
            "timing": "median of 3 complete measurements of exactly K steps (barrier + synchronize on both sides, CUDA events, max over ranks)",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: ONE batch of %d simulated Kepler long-cadence light curves per step, split evenly "
                                   "over the GPUs (%d per GPU), default OddsRatioDetector (flare 10x10 + "
                                   "poly/impulse/expdecay/reverse noise models)" % (args.curves, B),
                       "curves": args.curves, "curves_per_gpu": B, "samples": N, "samples_out": NK, "ignoreedges": True,
                       "bglen": det.bglen, "bgorder": det.bgorder,
                       "g_eval": G_EVAL, "g_nominal": G_NOMINAL, "threshold": THRESH,
                       "l2": "inputs rotate over %d distinct batches (%.0f MB) > 126 MB L2" % (NB, NB * B * N * 8 / 1e6),
                       "streams": nstreams,
                       "parallelism": "dp%d: strong scaling, the step's curves sharded over the ranks, no data-path collective" % world},
            "collective": {"op": "ncclAllReduce(sum, int64) of the per-step detection counters, once per job, inside the timed region",
                           "bytes": int(d_red.numel() * 8), "us": coll_us, "ranks": world},
            # The step's RESULT is what a flare search keeps of a batch (scripts/kepler_analysis_script.py:340-455:
            # oddsratio, then thresholder): the above-threshold regions with their maxima, and per-curve counts.
            "e2e": {"value": e2e_search_value, "unit": UNIT, "h2d_bytes_per_step": world * B * N * 8,
                    "d2h_bytes_per_step": world * ((B + 3) * 8 + MAXREG * 24), "steps": args.steps,
                    "gpu_launches": int(e2e_launches), "cuda_graphs": int(pipe.graphs),
                    "what": "bfl_pipe_submit / bfl_pipe_wait (C ABI) with page-locked host buffers, 6 jobs in flight: H2D of the "
                            "step's light curves (float64), noise sigma of every curve estimated on device (SG detrend + "
                            "periodogram), all hypotheses, region extraction, D2H of the step's region records (start, stop, "
                            "maximum, argmax) and per-curve counts -- every step's copies inside the timed region",
                    "regions_last_step_rank0": int(last_cnt[B]), "curves_with_region_last_step_rank0": int(last_cnt[B + 1])},
            # The same with the whole log-odds array coming back (what oddsratio() itself returns).  On an 8-GPU box
            # this one is bound by the host: 8 ranks' concurrent H2D + D2H get ~8 GB/s per direction per GPU
            # (tools/pcie_probe.py), 36 MB each way per step.
            "e2e_full_lnO": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * B * N * 8,
                             "d2h_bytes_per_step": world * B * NK * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "detections": {"curves_with_detection_last_step_rank0": n_det, "curves_per_gpu": B,
                           "job_regions_all_ranks": int(job[0]), "job_curves_with_detection_all_ranks": int(job[1])}}
    if sustained is not None:
        line["sustained"] = sustained
    if weak is not None:
        line["weak_scaling"] = weak

    if world == 1:
        # ---- the WEIGHTED path (row f4: every real Kepler file carries per-sample errors, so v_j = sk^2 + cle_j^2 varies
        # along the window): the same curves with synthetic flux errors through k_wcorr + k_wsolve, device-resident
        Bw = min(256, B)
        rngw = np.random.default_rng(6)
        cle_w = 0.3 + 0.2 * rngw.random((Bw, N)) + 0.05 * np.sin(np.arange(N) / 50.)[None, :]
        d_clew = torch.from_numpy(cle_w).to(dev)
        d_outw = torch.empty((Bw, NK), dtype=torch.float64, device=dev)

        def weighted_step():
            plan.detector_odds_dev(d_in[0].data_ptr(), d_clew.data_ptr(), False, d_sk[0].data_ptr(), Bw, N, 0, N, K0, K1,
                                   d_outw.data_ptr(), ctx=ctxs[0])
        with torch.cuda.stream(streams[0]):
            for _ in range(2):
                weighted_step()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nrep = 5
            ev0.record(streams[0])
            for _ in range(nrep):
                weighted_step()
            ev1.record(streams[0])
        torch.cuda.synchronize()
        ms_w = ev0.elapsed_time(ev1) / nrep
        lnO_w = d_outw[0].cpu().numpy()
        line["weighted_path"] = {"value": Bw * NK * G_EVAL / (ms_w * 1e-3), "unit": UNIT, "curves": Bw, "samples": N, "ms_per_call": ms_w,
                                 "what": "bfl_detector_odds_dev with per-sample flux errors (k_prep_whiten, k_wcorr, k_wsolve with the hypotheses folded in), "
                                         "inputs resident, CUDA events around %d calls" % nrep}

    if not args.no_cpu_baseline and world == 1:
        cores = set_host_threads()
        from oracle import oracle as orc_w
        ref_w, _ = orc_w.oddsratio(clc0[0], cle_w[0], cts, sk=float(sk[0]))
        ref_w = np.asarray(ref_w)
        if ref_w.shape[0] == N:
            ref_w = ref_w[K0:K1]                 # (the oracle drops the truncated-window samples itself by default)
        line["weighted_path"]["parity_vs_oracle_curve0"] = float(np.max(np.abs(lnO_w - ref_w) / np.maximum(1.0, np.abs(ref_w))))
        nc = min(args.cpu_curves, B)
        ups, secs, ref0, done = cpu_oracle_throughput(cts, clc0[:nc], max_seconds=12.0)
        got0 = plan.detector_odds(clc0[0], None, sk[0], k0=K0, k1=K1)
        ref0 = ref0[K0:K1]
        err = float(np.max(np.abs(got0 - ref0) / np.maximum(1.0, np.abs(ref0))))
        err_e2e = float(np.max(np.abs(lnO_host[0] - ref0) / np.maximum(1.0, np.abs(ref0))))
        port = {"value": ups, "unit": UNIT, "cores": cores, "threads": cores, "kind": "port",
                "sample": "%d light curves x %d samples of the same batch, %.1f s (C/OpenMP oracle)" % (done, N, secs)}
        line["parity_vs_oracle"] = {"curve0_device_path": err, "curve0_e2e_path_device_sigma": err_e2e,
                                    "note": "max |d| / max(1, |ref|) over ONE curve of the batch; the -m gpu tests hold the full parity suite"}
        if reference_available():
            n_ref = 1100
            lnO_ref, dt_ref = reference_oddsratio(cts[:n_ref], clc0[0, :n_ref])
            line["cpu_baseline"] = {"value": (n_ref - 54) * G_EVAL / dt_ref, "unit": UNIT, "cores": cores, "threads": cores,
                                    "kind": "reference",
                                    "sample": "the reference's own OddsRatioDetector.oddsratio() (unmodified package, ncpus=None) on 1 "
                                              "light curve x %d samples of the same batch, %.1f s" % (n_ref, dt_ref)}
            line["cpu_baseline_port"] = port
        else:
            line["cpu_baseline"] = port
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
