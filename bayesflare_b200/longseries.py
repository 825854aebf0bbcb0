"""One long series over several GPUs: time chunks with window halos.

A sample of the log odds needs only the ``+-bglen/2`` neighbourhood of the INPUT, so a long short-cadence
series is cut into contiguous output ranges, each rank uploads its range plus a halo from the host buffer,
runs the detector on it (``bfl_detector_odds_dev`` with the segment descriptor) and nothing is exchanged
between ranks on the data path.  The noise sigma is a property of the whole series: every rank estimates it
from the full curve (or takes it as an argument).  The pieces are collected with one all-gather.

    python -m torch.distributed.run --nproc-per-node 8 -m bayesflare_b200.longseries --samples 1040000
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

from .finder import OddsRatioDetector
from .lightcurve import Lightcurve
from .parallel import time_chunks

__all__ = ["oddsratio_time_chunks"]


def oddsratio_time_chunks(det, clc, sk=None, rank=0, world=1, gather=True):
    """Log odds of the series ``clc`` (on ``det``'s time base) computed as ``world`` halo'd time chunks, this
    process doing chunk ``rank``.  Returns ``(lnO, seconds)``: with ``gather`` the full edge-trimmed series on
    every rank (all-gather over the initialised process group), otherwise this rank's piece."""
    import torch
    clc = np.ascontiguousarray(clc, dtype=np.float64)
    N = len(clc)
    plan = det.plan()
    ctx = plan.ctx
    dev = torch.device("cuda", ctx.device if ctx.device >= 0 else torch.cuda.current_device())
    if sk is None:
        sk = float(det.noise_sigma_batch(clc)[0])
    chunks = time_chunks(N, world, det.bglen, bool(det.ignoreedges))
    k0, k1, s0, s1 = chunks[rank]
    d_sk = torch.tensor([sk], dtype=torch.float64, device=dev)
    pinned = torch.from_numpy(clc[s0:s1].copy()).pin_memory()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    d_seg = pinned.to(dev, non_blocking=False)
    d_out = torch.empty(k1 - k0, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(dev)
    plan.detector_odds_dev(d_seg.data_ptr(), 0, True, d_sk.data_ptr(), 1, N, s0, s1 - s0, k0, k1, d_out.data_ptr(),
                           noisepoly=det.noisepoly)
    ctx.sync()
    secs = time.perf_counter() - t0
    if not gather or world == 1:
        return d_out.cpu().numpy(), secs
    import torch.distributed as dist
    nmax = max(c[1] - c[0] for c in chunks)
    pad = torch.zeros(nmax, dtype=torch.float64, device=dev)
    pad[:k1 - k0] = d_out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    lnO = np.concatenate([b[:c[1] - c[0]].cpu().numpy() for b, c in zip(bufs, chunks)])
    return lnO, secs


def main():
    from .synthetic import DT_SHORT
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--samples", type=int, default=1040000, help="series length (a year of short cadence: 8 x 130,000)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = args.samples
    from .models import Flare
    rng = np.random.default_rng(17)                      # the same series on every rank: unit noise + a flare
    x = rng.standard_normal(N)                           # every ~20,000 samples
    cts = np.arange(N, dtype=np.float64) * DT_SHORT
    tpl = Flare(cts[:1200], amp=1.).model({"t0": cts[200], "amp": 1., "taugauss": 150., "tauexp": 600.})
    for pos in range(5000, N - 2000, 20011):
        x[pos - 200:pos + 1000] += rng.uniform(6., 14.) * tpl
    det = OddsRatioDetector(Lightcurve(cts=cts, clc=x), device=local_rank)
    sk = float(det.noise_sigma_batch(x)[0])
    oddsratio_time_chunks(det, x, sk=sk, rank=rank, world=world)                # warm-up: plan, buffers
    if world > 1:
        dist.barrier()
    lnO, secs = oddsratio_time_chunks(det, x, sk=sk, rank=rank, world=world)
    t = torch.tensor([secs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        oddsratio_time_chunks(det, x, sk=sk, rank=0, world=1)
        whole, secs1 = oddsratio_time_chunks(det, x, sk=sk, rank=0, world=1)
        units = (N - 2 * (det.bglen // 2)) * det.ngrid_eval
        print(json.dumps({"samples": N, "n_gpus": world, "chunk_seconds_max_over_ranks": float(t.item()),
                          "whole_series_on_one_gpu_seconds": secs1, "units": units,
                          "units_per_s_chunked": units / float(t.item()), "units_per_s_one_gpu": units / secs1,
                          "chunked_vs_whole_max_abs": float(np.max(np.abs(lnO - whole))),
                          "regions_above_16.5": int(np.sum((lnO[1:] > 16.5) & ~(lnO[:-1] > 16.5)))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
