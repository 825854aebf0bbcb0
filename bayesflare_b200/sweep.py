"""Injection sweep: detection efficiency versus signal-to-noise ratio and false-alarm count, the
workload of the reference's ``scripts/flare_detection_efficiency.py`` (:399-651; BASELINE config 5).

Every realisation is an independent light curve, so realisations are sharded across ranks
(`torchrun`, one process per GPU) with no data-path communication; each rank simulates its block,
runs the batched detector (noise estimate included) on its GPU, matches detections to injections
within +-2 samples (script :560-600) and histograms them; the integer counters are summed with one
all-reduce (NCCL on GPUs).

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

__all__ = ["detection_efficiency", "scripts_detector_kwargs"]


def scripts_detector_kwargs(dt=DT_LONG, bglen=55):
    """The detector configuration of scripts/flare_detection_efficiency.py:495-507: tail-veto noise
    estimate and an impulse model marginalised over every position of the window."""
    return dict(bglen=bglen, bgorder=4, noiseestmethod="tailveto", tvsigma=1.0,
                noiseimpulseparams={"t0": (0, (bglen - 1) * dt, bglen)})


def detection_efficiency(nreal, N=1639, dt=DT_LONG, threshold=16.5, nbins=50, snr_max=50., seed=0, sinusoid_amp=0.0,
                         detector_kwargs=None, batch=2048, rank=0, world=1, device=None, comm_device="cpu", group=None):
    """Run the sweep for the realisations owned by ``rank`` and reduce the counters over ranks.

    Returns a dict with ``edges`` (SNR bin edges), ``injected`` / ``detected`` (per-bin counts),
    ``efficiency``, ``false_alarms`` (above-threshold regions away from the injection) and timing."""
    kw = scripts_detector_kwargs(dt) if detector_kwargs is None else dict(detector_kwargs)
    lo, hi = shard_range(nreal, world, rank)
    edges = np.linspace(0., snr_max, nbins + 1)
    injected = np.zeros(nbins, dtype=np.int64)
    detected = np.zeros(nbins, dtype=np.int64)
    false_alarms = 0
    det = None
    t_sim = t_gpu = 0.0
    units = 0
    for first in range(lo, hi, batch):
        nb = min(batch, hi - first)
        t0 = time.perf_counter()
        cts, clc, truth = simulate_batch(nb, N, dt=dt, seed=seed, sinusoid_amp=sinusoid_amp, first=first,
                                         snr_range=(0., snr_max))
        t_sim += time.perf_counter() - t0
        if det is None:
            det = OddsRatioDetector(Lightcurve(cts=cts, clc=clc[0]), device=device, **kw)
        t0 = time.perf_counter()
        lnO = det.oddsratio_batch(clc)                 # [nb, N - 2h], noise sigma estimated on the device
        t_gpu += time.perf_counter() - t0
        h = det.bglen // 2
        units += nb * (N - 2 * h) * det.ngrid_eval
        above = lnO > threshold
        idx = truth["idx"] - h
        near = np.zeros_like(above)
        cols = np.arange(lnO.shape[1])[None, :]
        near[:] = np.abs(cols - idx[:, None]) <= 2
        hit = np.any(above & near, axis=1)
        injected += np.histogram(truth["snr"], edges)[0]
        detected += np.histogram(truth["snr"][hit], edges)[0]
        # false alarms: starts of above-threshold regions that do not touch the injection window
        starts = above & ~np.concatenate([np.zeros((nb, 1), bool), above[:, :-1]], axis=1)
        region_id = np.cumsum(starts, axis=1) * above
        for b in np.nonzero(above.any(axis=1))[0]:
            ids = np.unique(region_id[b][above[b]])
            good = np.unique(region_id[b][above[b] & near[b]])
            false_alarms += len(ids) - len(good)
    packed = np.concatenate([injected, detected, [false_alarms, units]]).astype(np.int64)
    packed = allreduce_counts(packed, device=comm_device, group=group)
    injected, detected = packed[:nbins], packed[nbins:2 * nbins]
    with np.errstate(invalid="ignore", divide="ignore"):
        eff = detected / injected
    return {"edges": edges, "injected": injected, "detected": detected, "efficiency": eff,
            "false_alarms": int(packed[-2]), "units": int(packed[-1]), "t_simulate_s": t_sim, "t_detector_s": t_gpu,
            "realisations": int(nreal), "local_realisations": int(hi - lo)}


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--realisations", type=int, default=10000)
    ap.add_argument("--samples", type=int, default=1639)
    ap.add_argument("--threshold", type=float, default=16.5)
    ap.add_argument("--sinusoid-amp", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm_device = "cpu"
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm_device = "cuda:%d" % local_rank
    t0 = time.perf_counter()
    res = detection_efficiency(args.realisations, N=args.samples, threshold=args.threshold, seed=args.seed,
                               sinusoid_amp=args.sinusoid_amp, rank=rank, world=world, device=local_rank,
                               comm_device=comm_device)
    wall = time.perf_counter() - t0
    if rank == 0:
        out = {"realisations": res["realisations"], "n_gpus": world, "wall_s": wall, "t_detector_s_rank0": res["t_detector_s"],
               "t_simulate_s_rank0": res["t_simulate_s"], "units": res["units"],
               "units_per_s_detector": res["units"] / world / max(res["t_detector_s"], 1e-9) * world,
               "false_alarms": res["false_alarms"],
               "efficiency_by_snr_decade": [float(res["detected"][i:i + 10].sum() / max(res["injected"][i:i + 10].sum(), 1))
                                            for i in range(0, 50, 10)],
               "injected": res["injected"].tolist(), "detected": res["detected"].tolist()}
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
