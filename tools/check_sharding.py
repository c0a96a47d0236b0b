#!/usr/bin/env python
"""Sharding invariance on real GPUs:  torchrun --nproc-per-node N tools/check_sharding.py [contigs]
A fixed synthetic workload is cut into N contiguous shards balanced by bases (parallel.balanced_partition); every rank counts and
scores its shard, the scores are gathered over NCCL, and rank 0 compares the gathered vector, element for element, with the vector
it gets by scoring the whole workload alone."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phamers_b200 import ops, parallel, pipeline  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
scorer = pipeline.ContigScorer()
# offsets of the whole workload (cheap: lengths only come from the generator's own table)
seq_all, off_all = ops.synth_contigs(20260101, 0, n_total)
bounds = parallel.balanced_partition(off_all.cpu().numpy(), world)
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
seq, off = ops.synth_contigs(20260101, lo, hi - lo)                       # the shard, generated independently of the whole
assert torch.equal(seq[:int(off[-1])], seq_all[int(off_all[lo]):int(off_all[hi])])
_, local = scorer.score_device(seq, off)
counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
gathered = parallel.gather_scores(local, counts) if world > 1 else local
if rank == 0:
    _, whole = scorer.score_device(seq_all, off_all)
    same = torch.equal(gathered, whole)
    print("world %d: shards %s contigs, gathered == single-GPU scores: %s" % (world, counts, same))
    assert same
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
