"""Pinned-host <-> device copy bandwidth of this box, one process per GPU (run plain or under torchrun).
Prints GB/s for H2D alone, D2H alone and both directions at once, for 4.5 MB and 36 MB buffers."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mb in (4.5, 36.0):
    n = int(mb * 1e6 / 8)
    h_in = torch.randn(n, dtype=torch.float64).pin_memory()
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    d_a = torch.empty(n, dtype=torch.float64, device="cuda")
    d_b = torch.randn(n, dtype=torch.float64, device="cuda")
    res = {}
    for mode in ("h2d", "d2h", "both"):
        reps = 20
        for timed in (False, True):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res[mode] = reps * n * 8 / dt / 1e9
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("pcie world=%d %.1f MB: h2d %.1f GB/s, d2h %.1f GB/s, both %.1f GB/s per direction (min over ranks)"
              % (world, mb, t[0].item(), t[1].item(), t[2].item()), flush=True)
if world > 1:
    dist.destroy_process_group()
