"""Injection sweep: detection efficiency versus signal-to-noise ratio and false-alarm count, the
workload of the reference's ``scripts/flare_detection_efficiency.py`` (:399-651; BASELINE config 5).

Every realisation is an independent light curve, so realisations are sharded across ranks
(`torchrun`, one process per GPU) with no data-path communication; each rank simulates its block,
runs the batched detector (noise estimate included) on its GPU, matches detections to injections
within +-2 samples (script :521-612) and histograms them; the integer counters are summed with one
all-reduce (NCCL on GPUs).  With ``--device-sim`` the simulation (Philox streams per realisation),
the matching and the histograms run on the device as well and only the counters leave it.

    python -m bayesflare_b200.sweep --realisations 10000
    python -m torch.distributed.run --nproc-per-node 8 -m bayesflare_b200.sweep --realisations 10000
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

from .finder import OddsRatioDetector
from .lightcurve import Lightcurve
from .parallel import allreduce_counts, shard_range
from .synthetic import DT_LONG, simulate_batch

__all__ = ["detection_efficiency", "scripts_detector_kwargs", "score_host"]


def scripts_detector_kwargs(dt=DT_LONG, bglen=55):
    """The detector configuration of scripts/flare_detection_efficiency.py:495-507: tail-veto noise
    estimate and an impulse model marginalised over every position of the window."""
    return dict(bglen=bglen, bgorder=4, noiseestmethod="tailveto", tvsigma=1.0,
                noiseimpulseparams={"t0": (0, (bglen - 1) * dt, bglen)})


def score_host(lnO, inj_idx, snr, threshold, edges):
    """The script's bookkeeping for a batch (flare_detection_efficiency.py:521-612) on the host:
    above-threshold runs expanded by +-2 samples and merged when they touch; the region holding the
    injection index is a detection, every other region a false positive.  Returns
    ``(injected, detected, false_alarms, regions)``; ``bfl_sweep_score_dev`` is the device version."""
    nb, nk = lnO.shape
    nbins = len(edges) - 1
    above = lnO > threshold
    hit = np.zeros(nb, dtype=bool)
    false_alarms = regions = 0
    for b in np.nonzero(above.any(axis=1))[0]:
        d = np.diff(np.concatenate(([0], above[b].astype(np.int8), [0])))
        starts, stops = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
        rs = re = -1
        segs = []
        for a, e in zip(starts, stops):
            xs, xe = max(a - 2, 0), min(e + 2, nk)
            if rs >= 0 and xs <= re:
                re = max(re, xe)
            else:
                if rs >= 0:
                    segs.append((rs, re))
                rs, re = xs, xe
        segs.append((rs, re))
        regions += len(segs)
        for a, e in segs:
            if inj_idx[b] >= 0 and a <= inj_idx[b] < e:
                hit[b] = True
            else:
                false_alarms += 1
    have = inj_idx >= 0
    injected = np.histogram(snr[have], edges)[0].astype(np.int64)
    detected = np.histogram(snr[have & hit], edges)[0].astype(np.int64)
    assert len(injected) == nbins
    return injected, detected, int(false_alarms), int(regions)


def detection_efficiency(nreal, N=1639, dt=DT_LONG, threshold=16.5, nbins=50, snr_max=50., seed=0, sinusoid_amp=0.0,
                         detector_kwargs=None, batch=2048, rank=0, world=1, device=None, comm_device="cpu", group=None,
                         device_sim=False, recipe="nstd", inject="flare"):
    """Run the sweep for the realisations owned by ``rank`` and reduce the counters over ranks.

    ``device_sim=False``: realisations are simulated on the host (NumPy streams seeded per realisation),
    the detector runs on the device, the bookkeeping on the host.  ``device_sim=True``: everything stays on
    the device -- simulation (``bfl_sweep_simulate_dev``, counter-based Philox streams per realisation),
    noise estimate, detector, matching and histograms (``bfl_sweep_score_dev``); only the counters return.

    ``recipe="script"`` follows the reference script to the letter: ``inject`` = ``"flare"`` or ``"transit"``
    (``--injtransit``, :291-312, 444-453), the injection is rescaled with the noise sigma ESTIMATED from the base
    curve by the tail veto (:429-434, 483-487) -- on the device: ``bfl_sweep_simulate_dev`` (base curve + drawn
    parameters), ``bfl_noise_sigma_dev``, ``bfl_sweep_inject_dev``.  ``recipe="nstd"`` (flares only) scales with the
    true noise level and removes the median, the bench workloads' convention.

    Returns a dict with ``edges`` (SNR bin edges), ``injected`` / ``detected`` (per-bin counts),
    ``efficiency``, ``false_alarms`` (regions that do not hold the injection) and timing."""
    kw = scripts_detector_kwargs(dt) if detector_kwargs is None else dict(detector_kwargs)
    lo, hi = shard_range(nreal, world, rank)
    edges = np.linspace(0., snr_max, nbins + 1)
    injected = np.zeros(nbins, dtype=np.int64)
    detected = np.zeros(nbins, dtype=np.int64)
    false_alarms = 0
    t_sim = t_gpu = 0.0
    units = 0
    cts = np.arange(N, dtype=np.float64) * dt
    det = OddsRatioDetector(Lightcurve(cts=cts, clc=np.zeros(N)), device=device, **kw)
    h = det.bglen // 2
    if device_sim and hi > lo:
        import torch
        plan = det.plan()
        ctx = plan.ctx
        dev = torch.device("cuda", ctx.device if ctx.device >= 0 else torch.cuda.current_device())
        nbm = min(batch, hi - lo)
        d_clc = torch.empty((nbm, N), dtype=torch.float64, device=dev)
        d_truth = torch.empty((nbm, 4), dtype=torch.float64, device=dev)
        d_sk = torch.empty(nbm, dtype=torch.float64, device=dev)
        d_lnO = torch.empty((nbm, N - 2 * h), dtype=torch.float64, device=dev)
        d_inj = torch.zeros(nbins, dtype=torch.int64, device=dev)
        d_det = torch.zeros(nbins, dtype=torch.int64, device=dev)
        d_cnt = torch.zeros(2, dtype=torch.int64, device=dev)
        spec = det.noise_spec()
        from ._lib import make_noise_spec
        spec_script = make_noise_spec("tailveto", det.bglen, det.bgorder, tvsigma=1.0)      # script :429-434
        d_skb = torch.empty(nbm, dtype=torch.float64, device=dev)
        kind = 2 if inject == "transit" else 1
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for first in range(lo, hi, batch):
            nb = min(batch, hi - first)
            if recipe == "script":
                ctx.sweep_simulate_dev(seed, first, nb, N, h, dt, 1.0, sinusoid_amp, 0., snr_max, 1 + kind,
                                       d_clc.data_ptr(), d_truth.data_ptr())
                ctx.noise_sigma_dev(spec_script, d_clc.data_ptr(), nb, N, d_skb.data_ptr())
                # (for a transit the kernel then clears the injection index: the script counts EVERY segment of the flare
                # detector as a false positive there, :616-623)
                ctx.sweep_inject_dev(kind, nb, N, dt, d_skb.data_ptr(), d_truth.data_ptr(), d_clc.data_ptr())
            else:
                ctx.sweep_simulate_dev(seed, first, nb, N, h, dt, 1.0, sinusoid_amp, 0., snr_max, 1,
                                       d_clc.data_ptr(), d_truth.data_ptr())
            ctx.noise_sigma_dev(spec, d_clc.data_ptr(), nb, N, d_sk.data_ptr())
            plan.detector_odds_dev(d_clc.data_ptr(), 0, True, d_sk.data_ptr(), nb, N, 0, N, h, N - h, d_lnO.data_ptr(),
                                   noisepoly=det.noisepoly)
            ctx.sweep_score_dev(d_lnO.data_ptr(), nb, N - 2 * h, h, d_truth.data_ptr(), threshold, nbins, snr_max,
                                d_inj.data_ptr(), d_det.data_ptr(), d_cnt.data_ptr())
            units += nb * (N - 2 * h) * det.ngrid_eval
        ctx.sync()
        t_gpu = time.perf_counter() - t0
        injected, detected = d_inj.cpu().numpy(), d_det.cpu().numpy()
        false_alarms = int(d_cnt[0].item())
    elif hi > lo:
        for first in range(lo, hi, batch):
            nb = min(batch, hi - first)
            t0 = time.perf_counter()
            _, clc, truth = simulate_batch(nb, N, dt=dt, seed=seed, sinusoid_amp=sinusoid_amp, first=first,
                                           snr_range=(0., snr_max), recipe=recipe,
                                           inject=(inject if recipe == "script" else True))
            t_sim += time.perf_counter() - t0
            t0 = time.perf_counter()
            lnO = det.oddsratio_batch(clc)                 # [nb, N - 2h], noise sigma estimated on the device
            t_gpu += time.perf_counter() - t0
            units += nb * (N - 2 * h) * det.ngrid_eval
            inj_idx = truth["idx"] - h
            if recipe == "script" and inject == "transit":
                inj_idx = np.full(nb, -1)                   # script :616-623: every segment is a false positive
            inj, dete, fa, _ = score_host(lnO, inj_idx, truth["snr"], threshold, edges)
            injected += inj
            detected += dete
            false_alarms += fa
    packed = np.concatenate([injected, detected, [false_alarms, units]]).astype(np.int64)
    packed = allreduce_counts(packed, device=comm_device, group=group)
    injected, detected = packed[:nbins], packed[nbins:2 * nbins]
    with np.errstate(invalid="ignore", divide="ignore"):
        eff = detected / injected
    return {"edges": edges, "injected": injected, "detected": detected, "efficiency": eff,
            "false_alarms": int(packed[-2]), "units": int(packed[-1]), "t_simulate_s": t_sim, "t_detector_s": t_gpu,
            "realisations": int(nreal), "local_realisations": int(hi - lo), "device_sim": bool(device_sim)}


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--realisations", type=int, default=10000)
    ap.add_argument("--samples", type=int, default=1639)
    ap.add_argument("--threshold", type=float, default=16.5)
    ap.add_argument("--sinusoid-amp", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device-sim", action="store_true", help="simulate, match and histogram on the device too")
    ap.add_argument("--recipe", default="nstd", choices=["nstd", "script"],
                    help="script: injection scaled with the sigma estimated from the base curve, as the reference script does")
    ap.add_argument("--injtransit", action="store_true", help="inject transits instead of flares (script recipe)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm_device = "cpu"
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")           # no NCCL version banner in front of the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm_device = "cuda:%d" % local_rank
    t0 = time.perf_counter()
    res = detection_efficiency(args.realisations, N=args.samples, threshold=args.threshold, seed=args.seed,
                               sinusoid_amp=args.sinusoid_amp, rank=rank, world=world, device=local_rank,
                               comm_device=comm_device, device_sim=args.device_sim,
                               recipe="script" if args.injtransit else args.recipe,
                               inject="transit" if args.injtransit else "flare")
    wall = time.perf_counter() - t0
    if rank == 0:
        out = {"realisations": res["realisations"], "n_gpus": world, "device_sim": res["device_sim"],
               "recipe": "script" if args.injtransit else args.recipe, "injection": "transit" if args.injtransit else "flare", "wall_s": wall, "t_detector_s_rank0": res["t_detector_s"],
               "t_simulate_s_rank0": res["t_simulate_s"], "units": res["units"],
               "units_per_s_detector": res["units"] / world / max(res["t_detector_s"], 1e-9) * world,
               "false_alarms": res["false_alarms"], "false_alarms_per_realisation": res["false_alarms"] / max(res["realisations"], 1),
               "efficiency_by_snr_decade": [float(res["detected"][i:i + 10].sum() / max(res["injected"][i:i + 10].sum(), 1))
                                            for i in range(0, 50, 10)],
               "injected": res["injected"].tolist(), "detected": res["detected"].tolist()}
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
