#!/usr/bin/env python
"""Host -> device copy floor of the end-to-end path:  torchrun --nproc-per-node N tools/bench_h2d.py [GiB]
Every rank copies its own pinned buffer to its own GPU with bare asynchronous copies (torch's copy_ = one cudaMemcpyAsync), all
ranks at once between two barriers; prints per-rank and aggregate GB/s (aggregate = all bytes / the slowest rank's time).
bench.py's e2e figure moves 1 byte per base, so this aggregate is the ceiling of `e2e` in bases/s on the same box."""
import os
import sys
import time

import torch
import torch.distributed as dist

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
n = int(gib * (1 << 30))
host = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
host.fill_(65)
dev = torch.empty((n,), dtype=torch.uint8, device="cuda")
dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
reps = 4
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mine = torch.tensor([dt], dtype=torch.float64, device="cuda")
slowest = mine.clone()
if world > 1:
    dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
else:
    every = [mine]
if rank == 0:
    per_rank = [n * reps / float(t.item()) / 1e9 for t in every]
    print("h2d floor, %d rank(s), %.1f GiB x %d per rank: per rank GB/s %s, aggregate %.1f GB/s (cpu count %s)"
          % (world, gib, reps, ["%.1f" % v for v in per_rank], world * n * reps / float(slowest.item()) / 1e9, os.cpu_count()), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
