"""Host-side logic of the multi-GPU layout on CPU: world-size-2 gloo processes.

The data path has one collective, the gather of per-contig scores; everything else is a contiguous partition of the contigs
balanced by bases.  These tests need no GPU: the per-rank "scores" are a deterministic function of the global contig index,
so the gathered vector must be identical for every world size and every rank.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phamers_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _lengths(n, seed=7):
    rng = np.random.default_rng(seed)
    return np.clip(np.round(np.exp(rng.normal(np.log(10000.0), 1.0, size=n))), 1000, 100000).astype(np.int64)


def test_balanced_partition_properties():
    lengths = _lengths(5000)
    offsets = np.concatenate(([0], np.cumsum(lengths)))
    for world in (1, 2, 3, 4, 8):
        bounds = parallel.balanced_partition(offsets, world)
        assert bounds[0] == 0 and bounds[-1] == len(lengths) and len(bounds) == world + 1
        assert np.all(np.diff(bounds) >= 0)
        per_rank = np.array([offsets[bounds[r + 1]] - offsets[bounds[r]] for r in range(world)])
        assert per_rank.sum() == offsets[-1]
        # no rank is off the ideal share by more than one (longest) contig
        assert np.max(np.abs(per_rank - offsets[-1] / world)) <= lengths.max()


def test_balanced_partition_edge_cases():
    assert list(parallel.balanced_partition(np.array([0]), 4)) == [0, 0, 0, 0, 0]                 # no contigs
    assert list(parallel.balanced_partition(np.array([0, 10]), 2))[-1] == 1                        # fewer contigs than ranks
    b = parallel.balanced_partition(np.array([0, 0, 0, 5, 5, 9]), 2)                               # empty contigs inside
    assert b[0] == 0 and b[-1] == 5 and b[1] in (3, 4)


def _worker(rank, world, port, n_contigs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lengths = _lengths(n_contigs)
        offsets = np.concatenate(([0], np.cumsum(lengths)))
        bounds = parallel.balanced_partition(offsets, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        # stand-in for the per-contig score of this rank's shard: depends on the global contig index only
        local = torch.from_numpy(np.tanh((np.arange(lo, hi) % 97 - 48) / 30.0) + np.sign(np.arange(lo, hi) % 5 - 2.5))
        counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
        gathered = parallel.gather_scores(local, counts)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), gathered.numpy())
        # equal shards (what bench.py has: a fixed number of contigs per GPU) take the path without padding
        same = torch.arange(rank * 50, (rank + 1) * 50, dtype=torch.float64)
        equal = parallel.gather_scores(same, [50] * world)
        assert torch.equal(equal, torch.arange(0, 50 * world, dtype=torch.float64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_scores_is_world_size_invariant(tmp_path, world):
    n = 1237
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    idx = np.arange(n)
    expect = np.tanh((idx % 97 - 48) / 30.0) + np.sign(idx % 5 - 2.5)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert got.dtype == np.float64 and np.array_equal(got, expect)
