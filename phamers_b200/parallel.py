"""
Multi-GPU layout of the hot path: one process per GPU, contigs sharded, reference features replicated, and a single
collective -- the gather of per-contig scores.  Contigs are independent in all three stages
(reference loops scripts/kmer.py:137-139, scripts/phamer.py:251-255), so there is no other exchange.

torch.distributed is the plumbing: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def balanced_partition(offsets, world_size):
    """Contiguous contig ranges with ~equal BASES per rank (lengths are lognormal, so equal contig counts would not
    balance).  offsets: int64[n+1] (host).  Returns int64[world_size+1] contig boundaries."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets.shape[0] - 1
    total = int(offsets[-1] - offsets[0])
    targets = offsets[0] + (total * np.arange(1, world_size, dtype=np.float64) / world_size)
    cuts = np.searchsorted(offsets, targets, side="left")
    bounds = np.concatenate(([0], np.clip(cuts, 0, n), [n])).astype(np.int64)
    return np.maximum.accumulate(bounds)


def gather_scores(local_scores, counts_per_rank, group=None):
    """All ranks contribute their shard's scores; every rank gets the full vector in contig order.
    One all_gather over equally padded shards (payload 8 B per contig)."""
    world = dist.get_world_size(group)
    counts_per_rank = [int(c) for c in counts_per_rank]
    width = max(counts_per_rank) if counts_per_rank else 0
    if counts_per_rank and min(counts_per_rank) == width and local_scores.is_contiguous():
        out = torch.empty((world * width,), dtype=local_scores.dtype, device=local_scores.device)     # equal shards: no padding, no copy
        dist.all_gather_into_tensor(out, local_scores, group=group)
        return out
    padded = torch.zeros((width,), dtype=local_scores.dtype, device=local_scores.device)
    padded[:local_scores.numel()] = local_scores
    out = torch.empty((world * width,), dtype=local_scores.dtype, device=local_scores.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * width:r * width + counts_per_rank[r]] for r in range(world)]
    return torch.cat(pieces) if pieces else out
