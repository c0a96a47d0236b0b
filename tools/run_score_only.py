#!/usr/bin/env python
"""Small driver for profiling the scoring kernels alone: N synthetic contigs counted once, scored R times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phamers_b200 import ops, pipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scorer = pipeline.ContigScorer()
seq, off = ops.synth_contigs(20260101, 0, n)
_, freq = ops.count_cuda(seq, off, 4, counts=False, freq=True)
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.score_cuda(freq, scorer.refs, scorer.n_positive, scorer.cent_pos, scorer.cent_neg, 3)
    e1.record()
    torch.cuda.synchronize()
    print("score %d rows: %.3f ms  %s" % (n, e0.elapsed_time(e1), ops.score_stats()))
