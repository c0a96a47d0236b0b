#!/usr/bin/env python
"""Host scikit-learn fit against the device Lloyd iterations on the shipped reference sets (k = 86) and on an enlarged set."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phamers_b200 import kmer, references
_, pos_c, _, neg_c = references.load_reference_counts()
pos = kmer.normalize_counts(pos_c)
out = {}
for name, data in (("positive_2255", pos), ("resampled_50000", None)):
    if data is None:
        rng = np.random.default_rng(0)
        rows = pos_c[rng.integers(0, len(pos_c), size=50000)].astype(np.float64)
        data = kmer.normalize_counts(rng.poisson(rows / rows.sum(axis=1, keepdims=True) * 20000))
    references.kmeans_assign_device(data[:500], 5)                      # warm-up (library, allocator)
    t0 = time.perf_counter(); a = references.kmeans_assign_device(data, 86); t1 = time.perf_counter()
    b = references.kmeans_assign(data, 86); t2 = time.perf_counter()
    out[name] = {"device_s": t1 - t0, "host_sklearn_s": t2 - t1, "labels_equal": bool(np.array_equal(a, b))}
print(json.dumps(out))
