#!/usr/bin/env python
"""GPU experiment: shortlist-margin fallback rate and ranking error of the tensor-core scorer on different query
families (bench workload, reference-like resamples at several depths)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phamers_b200 import _lib, kmer, ops, references  # noqa: E402


def main():
    pos, neg = references.load_reference_features(equalize=True)
    cents = references.reference_centroids(pos, neg)
    refs = torch.from_numpy(np.vstack((pos, neg))).cuda()
    cp = torch.from_numpy(np.ascontiguousarray(cents[0])).cuda()
    cn = torch.from_numpy(np.ascontiguousarray(cents[1])).cuda()
    _, pos_c, _, neg_c = references.load_reference_counts()
    both = np.vstack((pos_c[:len(pos)], neg_c[:len(neg)])).astype(np.float64)
    rng = np.random.default_rng(0)

    def run(name, pts):
        ops.set_score_path("tc")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ops.score_cuda(pts, refs, len(pos), cp, cn, 3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = ops.score_stats()
        fb, rel = st["fallback_rows"], st["max_bound_usage"]
        print("%-34s n=%7d  fallback %7d (%.3f%%)  remeasured %7d  max err %.3g  rel 2^%.1f  %.1f ms"
              % (name, pts.shape[0], fb, 100.0 * fb / pts.shape[0], st["rows_remeasured"], st["max_rank_error"],
                 np.log2(max(rel, 1e-30)), dt * 1e3))
        ops.set_score_path("auto")
        return out

    _lib.set_option("score_stats", int(os.environ.get("TC_STATS", "1")))
    seq, off = ops.synth_contigs(20260101, 0, 200000)
    _, freq = ops.count_cuda(seq, off, 4, counts=False, freq=True)
    run("bench workload (iid contigs)", freq)
    run("bench workload (iid contigs) again", freq)
    for depth in (1000, 5000, 20000, 100000, 1000000):
        rows = np.stack([rng.multinomial(depth, both[i] / both[i].sum()) for i in rng.integers(0, len(both), size=20000)])
        run("reference resample depth %d" % depth, torch.from_numpy(kmer.normalize_counts(rows)).cuda())
    run("exact reference rows", refs)


if __name__ == "__main__":
    main()
