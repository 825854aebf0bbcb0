"""Sharding and result aggregation for multi-GPU runs (one process per GPU, ``torch.distributed``).

The odds-ratio path needs **no data-path collective** (SURVEY.md 8e): light curves are independent
and a sample depends only on its +-bglen/2 neighbourhood of the *inputs*, so

* batches are split into contiguous blocks of curves (:func:`shard_range`),
* long series are split into time chunks that carry a halo (:func:`time_chunks`), and
* only results travel: detection records and efficiency / false-alarm counts
  (:func:`gather_detections`, :func:`allreduce_counts`) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

__all__ = ["gather_objects", "shard_range", "time_chunks", "gather_detections", "allreduce_counts"]


def shard_range(n, world, rank):
    """Contiguous block ``[lo, hi)`` of ``n`` items owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def time_chunks(N, nchunks, bglen, ignoreedges=True):
    """Split samples of one series into ``nchunks`` output ranges ``[k0, k1)`` with the input segment
    ``[seg0, seg1)`` each needs (its range plus a ``bglen//2`` halo, clipped to the series).  With
    ``ignoreedges`` the first/last ``bglen//2`` samples are not produced (find.py:846-851)."""
    h = int(bglen) // 2
    first, last = (h, N - h) if ignoreedges else (0, N)
    out = []
    for c in range(int(nchunks)):
        lo, hi = shard_range(last - first, nchunks, c)
        k0, k1 = first + lo, first + hi
        if k1 > k0:
            out.append((k0, k1, max(0, k0 - h), min(N, k1 + h)))
    return out


def _dist():
    import torch.distributed as dist
    return dist


def gather_detections(records, device="cpu", group=None):
    """All-gather variable-length detection records.

    ``records``: float64 array ``[n_local, ncol]`` (e.g. curve id, start, stop, argmax, max lnO).
    Returns the concatenation over ranks (rank order) on every rank.  Two collectives: the counts,
    then one padded all-gather -- message sizes are KB-scale, latency bound."""
    import torch
    dist = _dist()
    records = np.ascontiguousarray(records, dtype=np.float64)
    if records.ndim != 2:
        raise ValueError("records must be a 2-D array")
    if not (dist.is_available() and dist.is_initialized()):
        return records.copy()
    world = dist.get_world_size(group)
    ncol = records.shape[1]
    n_local = torch.tensor([records.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    pad = torch.zeros((nmax, ncol), dtype=torch.float64, device=device)
    if records.shape[0]:
        pad[:records.shape[0]] = torch.from_numpy(records).to(device)
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    parts = [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, ncol))


def allreduce_counts(counts, device="cpu", group=None):
    """Sum integer counters (efficiency histogram, false-alarm counts) over ranks."""
    import torch
    dist = _dist()
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()):
        return counts.copy()
    t = torch.from_numpy(counts.copy()).to(device)
    dist.all_reduce(t, group=group)
    return t.cpu().numpy()


def gather_objects(obj, group=None):
    """All-gather one picklable object per rank (e.g. the result dictionary of a sharded search); returns the
    list in rank order on every rank, or ``[obj]`` without a process group."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out
