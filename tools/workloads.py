"""Synthetic workloads shared by bench.py and the GPU parity tests.  Bench / test infrastructure: the
product package never imports this module."""
import numpy as np
import torch

SEED = 20260101


def enlarged_references(pos, neg, n_refs, seed=SEED, depth=20000.0):
    """BASELINE configs[4]: `n_refs` synthetic reference rows, the first half labelled phage.  Each row
    is a shipped reference row re-sampled as ~`depth` 4-mers (Poisson counts) and normalised.  pos / neg: the shipped features
    (numpy float64 [., 256]).  Returns (refs float64 CUDA tensor [n_refs, 256], n_positive)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    half = n_refs // 2
    parts = []
    for src_rows, m in ((pos, half), (neg, n_refs - half)):
        src_t = torch.from_numpy(np.ascontiguousarray(src_rows)).cuda()
        pick = torch.randint(0, src_t.shape[0], (m,), generator=g, device="cuda")
        big = torch.empty((m, 256), dtype=torch.float64, device="cuda")
        for lo in range(0, m, 1 << 17):
            lam = (src_t[pick[lo:lo + (1 << 17)]] * depth).float()
            c = torch.poisson(lam, generator=g).double()
            big[lo:lo + (1 << 17)] = c / c.sum(dim=1, keepdim=True).clamp_min(1.0)
        parts.append(big)
    return torch.cat(parts), half
