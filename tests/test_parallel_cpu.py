"""CPU tests of the N>1 host logic: sharding helpers and result aggregation with a world_size-2
gloo process group (the same code runs over NCCL on the GPUs)."""
import os
import socket

import numpy as np
import pytest

from bayesflare_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = parallel.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [parallel.shard_range(n, world, r)[1] - parallel.shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("ignoreedges", [True, False])
def test_time_chunks_cover_series_with_halos(ignoreedges):
    N, W = 130000, 55
    chunks = parallel.time_chunks(N, 8, W, ignoreedges)
    h = W // 2
    first, last = (h, N - h) if ignoreedges else (0, N)
    assert chunks[0][0] == first and chunks[-1][1] == last
    for (k0, k1, s0, s1), nxt in zip(chunks, chunks[1:] + [None]):
        assert s0 == max(0, k0 - h) and s1 == min(N, k1 + h)      # every window is inside its segment
        if nxt:
            assert nxt[0] == k1
    assert parallel.time_chunks(100, 200, 55)[-1][1] == 100 - 27      # more chunks than samples: empties dropped


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "detects" a different number of flares in its shard of 10 curves
        lo, hi = parallel.shard_range(10, world, rank)
        rec = np.array([[c, 100 + c, 105 + c, 102 + c, 17.0 + c] for c in range(lo, hi) if c % 3 != 1], dtype=float)
        rec = rec.reshape(-1, 5)
        allrec = parallel.gather_detections(rec)
        hist = np.zeros(50, dtype=np.int64)
        hist[rank::7] = rank + 1
        tot = parallel.allreduce_counts(hist)
        empty = parallel.gather_detections(np.zeros((0, 5)))
        # per-rank result dictionaries of a sharded flare search, merged as rank 0 would
        from bayesflare_b200.search import merge_results, new_result
        mine = new_result(16.5, nstars=hi - lo)
        mine["Star list"] = list(range(lo, hi)); mine["No. stars analysed"] = hi - lo
        mine["Flaring stars"] = [c for c in range(lo, hi) if c % 4 == 0]
        mine["No. flares"] = len(mine["Flaring stars"])
        for c in mine["Flaring stars"]:
            mine["Flaring star data"]["KPLR%09d" % c] = {"Kepler ID": c, "No. flares": 1}
        merged = merge_results(parallel.gather_objects(mine))
        q.put((rank, allrec, tot, empty.shape, merged))
    finally:
        dist.destroy_process_group()


def test_gather_and_reduce_world2_gloo():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([[c, 100 + c, 105 + c, 102 + c, 17.0 + c] for c in range(10) if c % 3 != 1], dtype=float)
    hist = np.zeros(50, dtype=np.int64)
    hist[0::7] += 1
    hist[1::7] += 2
    for rank, allrec, tot, eshape, merged in res:
        np.testing.assert_array_equal(allrec, expect)      # identical on every rank, rank order preserved
        np.testing.assert_array_equal(tot, hist)
        assert merged["Star list"] == list(range(10)) and merged["No. stars analysed"] == 10
        assert merged["Flaring stars"] == [0, 4, 8] and merged["No. flares"] == 3
        assert sorted(merged["Flaring star data"]) == ["KPLR000000000", "KPLR000000004", "KPLR000000008"]
        assert merged["Total no. of stars"] == 10
        assert eshape == (0, 5)


def test_single_process_is_a_no_op():
    rec = np.arange(10.0).reshape(2, 5)
    np.testing.assert_array_equal(parallel.gather_detections(rec), rec)
    np.testing.assert_array_equal(parallel.allreduce_counts([1, 2, 3]), [1, 2, 3])


def test_sweep_bookkeeping_matches_the_scripts_rules():
    """scripts/flare_detection_efficiency.py:521-612 on hand-made log odds: runs are expanded by +-2 samples
    and merged when they touch; the region holding the injection is the detection, the rest false positives"""
    import numpy as np
    from bayesflare_b200.sweep import score_host
    nk = 60
    lnO = np.full((5, nk), -1.0)
    lnO[0, 10:13] = 20.          # one run; injection at 14 -> inside the expanded region [8, 15)
    lnO[1, 10:13] = 20.          # injection at 15 -> just outside: a false positive, no detection
    lnO[2, 10:12] = 20.; lnO[2, 16:18] = 20.     # [8,14) and [14,20) touch -> ONE region holding 18
    lnO[3, 0] = 20.; lnO[3, nk - 1] = 20.; lnO[3, 30] = 20.      # clamped at the ends; injection at 30
    lnO[4, 5] = 20.              # no injection recorded: always a false positive
    inj = np.array([14, 15, 18, 30, -1])
    snr = np.array([5., 15., 25., 50., 35.])
    edges = np.linspace(0., 50., 6)
    injected, detected, fa, regions = score_host(lnO, inj, snr, 16.5, edges)
    assert injected.tolist() == [1, 1, 1, 0, 1]      # 50.0 falls in the last (inclusive) bin; curve 4 not injected
    assert detected.tolist() == [1, 0, 1, 0, 1]
    assert fa == 1 + 0 + 2 + 1 and regions == 1 + 1 + 1 + 3 + 1
