"""One long series over several GPUs: time chunks with window halos.

A sample of the log odds needs only the ``+-bglen/2`` neighbourhood of the INPUT, so a long short-cadence
series is cut into contiguous output ranges, each rank uploads its range plus a halo from the host buffer,
runs the detector on it (``bfl_detector_odds_dev`` with the segment descriptor) and nothing is exchanged
between ranks on the data path.  The noise sigma is a property of the whole series: every rank estimates it
from the full curve (or takes it as an argument).

What the ranks exchange is what north_star prescribes -- candidate detections: every rank extracts the
above-threshold regions of its chunk on the device (``bfl_batch_regions_dev``), and the region records
``(start, stop, argmax, max)`` in series coordinates are all-gathered (two KB-scale NCCL collectives,
:func:`bayesflare_b200.parallel.gather_detections`); a region cut by a chunk boundary comes back as two
records that touch and is joined on the host.  The log odds themselves stay on the rank that made them.

A rank's chunk is itself cut into sub-chunks (each with its halo) that rotate over a few streams, so the
upload of one overlaps the kernels of the previous.

    python -m torch.distributed.run --nproc-per-node 8 -m bayesflare_b200.longseries --samples 1040000
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

from . import _lib
from .finder import OddsRatioDetector
from .lightcurve import Lightcurve
from .parallel import gather_detections, time_chunks

__all__ = ["oddsratio_time_chunks", "regions_time_chunks", "join_regions"]


def _subchunks(k0, k1, N, h, nsub):
    """split the output range [k0, k1) into nsub pieces, each with the input segment it needs"""
    out = []
    n = k1 - k0
    for i in range(nsub):
        a, b = k0 + n * i // nsub, k0 + n * (i + 1) // nsub
        if b > a:
            out.append((a, b, max(0, a - h), min(N, b + h)))
    return out


class _ChunkRunner:
    """Device state of one rank for one series length: sub-chunk buffers, streams, contexts (reused across calls)."""

    def __init__(self, det, N, rank, world, nsub=4, nstreams=2, max_regions=4096):
        import torch
        self.torch = torch
        self.det, self.N = det, int(N)
        self.plan = det.plan()
        ctx0 = self.plan.ctx
        self.dev = torch.device("cuda", ctx0.device if ctx0.device >= 0 else torch.cuda.current_device())
        self.h = det.bglen // 2
        self.chunks = time_chunks(N, world, det.bglen, bool(det.ignoreedges))
        self.k0, self.k1, self.s0, self.s1 = self.chunks[rank]
        self.subs = _subchunks(self.k0, self.k1, N, self.h, max(1, min(nsub, (self.k1 - self.k0) // 4096 or 1)))
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(min(nstreams, len(self.subs)))]
        self.ctxs = [_lib.Context(self.dev.index, s.cuda_stream) for s in self.streams]
        self.d_sk = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.d_seg = [torch.empty(s1 - s0, dtype=torch.float64, device=self.dev) for (_, _, s0, s1) in self.subs]
        self.d_out = torch.empty(self.k1 - self.k0, dtype=torch.float64, device=self.dev)
        self.max_regions = int(max_regions)
        self.d_rec = torch.zeros(self.max_regions * 3, dtype=torch.float64, device=self.dev)      # 24-byte records
        self.d_cnt = torch.zeros(4, dtype=torch.int64, device=self.dev)
        self.pinned = None
        self.ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in self.subs]

    def run(self, clc, sk, thresh=None):
        """Upload + detector (+ region extraction) of this rank's chunk.  Returns (seconds, upload_ms_sum)."""
        torch = self.torch
        if self.pinned is None:
            self.pinned = torch.empty(self.s1 - self.s0, dtype=torch.float64).pin_memory()
        self.pinned.copy_(torch.from_numpy(clc[self.s0:self.s1]))      # host staging, outside the timed region
        self.d_sk.fill_(float(sk))
        torch.cuda.synchronize(self.dev)
        t0 = time.perf_counter()
        for i, (a, b, s0, s1) in enumerate(self.subs):
            w = i % len(self.streams)
            with torch.cuda.stream(self.streams[w]):
                e0, e1 = self.ev[i]
                e0.record()
                self.d_seg[i].copy_(self.pinned[s0 - self.s0:s1 - self.s0], non_blocking=True)
                e1.record()
                self.plan.detector_odds_dev(self.d_seg[i].data_ptr(), 0, True, self.d_sk.data_ptr(), 1, self.N, s0, s1 - s0,
                                            a, b, self.d_out[a - self.k0:].data_ptr(), noisepoly=self.det.noisepoly,
                                            ctx=self.ctxs[w])
        for s in self.streams[1:]:
            self.streams[0].wait_stream(s)
        if thresh is not None:
            with torch.cuda.stream(self.streams[0]):
                self.ctxs[0].batch_regions_dev(self.d_out.data_ptr(), 1, self.k1 - self.k0, thresh, self.d_rec.data_ptr(),
                                               self.max_regions, self.d_cnt.data_ptr())
        torch.cuda.synchronize(self.dev)
        secs = time.perf_counter() - t0
        return secs, float(sum(e0.elapsed_time(e1) for e0, e1 in self.ev))

    def regions(self):
        """this chunk's region records as float64 rows (start, stop, argmax, max) in series coordinates"""
        cnt = self.d_cnt.cpu().numpy()
        n = int(cnt[3])
        if n > self.max_regions:
            raise RuntimeError("more than %d regions in one chunk: raise max_regions" % self.max_regions)
        raw = self.d_rec.cpu().numpy().tobytes()
        rec = np.frombuffer(raw, dtype=_lib.REGION_DTYPE)[:n]
        rec = np.sort(rec, order="start")
        off = self.k0 - (self.h if self.det.ignoreedges else 0)     # index into the edge-trimmed series, as oddsratio() returns it
        out = np.empty((n, 4))
        out[:, 0], out[:, 1], out[:, 2], out[:, 3] = rec["start"] + off, rec["stop"] + off, rec["argmax"] + off, rec["maxval"]
        return out


_runners = {}


def _runner(det, N, rank, world, **kw):
    key = (id(det), int(N), rank, world)
    if key not in _runners:
        _runners[key] = _ChunkRunner(det, N, rank, world, **kw)
    return _runners[key]


def join_regions(records):
    """Join records that touch (a region cut by a chunk boundary): rows (start, stop, argmax, max) sorted by start."""
    records = records[np.argsort(records[:, 0], kind="stable")] if len(records) else records
    out = []
    for r in records:
        if out and out[-1][1] == r[0]:
            prev = out[-1]
            prev[1] = r[1]
            if r[3] > prev[3]:                 # the earlier piece wins ties, as np.argmax does
                prev[2], prev[3] = r[2], r[3]
        else:
            out.append(list(r))
    return np.array(out).reshape(-1, 4)


def regions_time_chunks(det, clc, thresh, sk=None, rank=0, world=1):
    """Above-threshold regions of the series ``clc`` computed as ``world`` halo'd time chunks (this process does
    chunk ``rank``) and gathered over the initialised process group: rows ``(start, stop, argmax, max lnO)`` in the
    coordinates of the edge-trimmed series ``oddsratio()`` returns.  Returns ``(regions, seconds, upload_ms)``."""
    clc = np.ascontiguousarray(clc, dtype=np.float64)
    if sk is None:
        sk = float(det.noise_sigma_batch(clc)[0])
    r = _runner(det, len(clc), rank, world)
    secs, up_ms = r.run(clc, sk, thresh=thresh)
    t0 = time.perf_counter()
    mine = r.regions()
    allrec = gather_detections(mine, device=r.dev) if world > 1 else mine
    secs += time.perf_counter() - t0
    return join_regions(allrec), secs, up_ms


def oddsratio_time_chunks(det, clc, sk=None, rank=0, world=1, gather=True):
    """Log odds of the series ``clc`` (on ``det``'s time base) computed as ``world`` halo'd time chunks, this
    process doing chunk ``rank``.  Returns ``(lnO, seconds)``: with ``gather`` the full edge-trimmed series on
    every rank (one padded all-gather of the log odds -- for checks; a search uses :func:`regions_time_chunks`),
    otherwise this rank's piece."""
    import torch
    clc = np.ascontiguousarray(clc, dtype=np.float64)
    if sk is None:
        sk = float(det.noise_sigma_batch(clc)[0])
    r = _runner(det, len(clc), rank, world)
    secs, _ = r.run(clc, sk)
    if not gather or world == 1:
        return r.d_out.cpu().numpy(), secs
    import torch.distributed as dist
    nmax = max(c[1] - c[0] for c in r.chunks)
    pad = torch.zeros(nmax, dtype=torch.float64, device=r.dev)
    pad[:r.k1 - r.k0] = r.d_out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    lnO = np.concatenate([b[:c[1] - c[0]].cpu().numpy() for b, c in zip(bufs, r.chunks)])
    return lnO, secs


def main():
    from .synthetic import DT_SHORT
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--samples", type=int, default=1040000, help="series length (a year of short cadence: 8 x 130,000)")
    ap.add_argument("--threshold", type=float, default=16.5)
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
    for _ in range(2):                                                           # warm-up: plan, buffers, NCCL
        regions_time_chunks(det, x, args.threshold, sk=sk, rank=rank, world=world)
    best, best_up = None, None
    for _ in range(5):
        if world > 1:
            dist.barrier()
        reg, secs, up_ms = regions_time_chunks(det, x, args.threshold, sk=sk, rank=rank, world=world)
        t = torch.tensor([secs, up_ms * 1e-3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if best is None or float(t[0]) < best:
            best, best_up = float(t[0]), float(t[1])
    if rank == 0:
        one = OddsRatioDetector(Lightcurve(cts=cts, clc=x), device=local_rank)
        for _ in range(2):
            reg1, secs1, up1 = regions_time_chunks(one, x, args.threshold, sk=sk, rank=0, world=1)
        units = (N - 2 * (det.bglen // 2)) * det.ngrid_eval
        print(json.dumps({"samples": N, "n_gpus": world, "seconds_max_over_ranks": best, "of_which_upload_seconds": best_up,
                          "one_gpu_seconds": secs1, "one_gpu_upload_seconds": up1 * 1e-3, "units": units,
                          "units_per_s": units / best, "units_per_s_one_gpu": units / secs1,
                          "speedup_vs_one_gpu": secs1 / best, "regions": int(len(reg)),
                          "regions_identical_to_one_gpu": bool(np.array_equal(reg, reg1)),
                          "exchanged_bytes_per_rank": int(len(reg) * 32),
                          "what": "upload of the rank's chunk (+ halo), detector, region extraction, all-gather of the region "
                                  "records; the log odds stay on the device"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
