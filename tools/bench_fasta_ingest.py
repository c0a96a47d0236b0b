#!/usr/bin/env python
"""Throughput of the device FASTA tokeniser (phm_fasta_index + phm_fasta_extract) on a synthetic multi-record file with
60-column lines, against the host (numpy) tokeniser on the same bytes.   python tools/bench_fasta_ingest.py [contigs]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phamers_b200 import fileIO, ops  # noqa: E402

n_contigs = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
seq, off = ops.synth_contigs(20260101, 0, n_contigs)
h_seq, h_off = seq.cpu().numpy(), off.cpu().numpy()
parts = []
for i in range(n_contigs):
    body = h_seq[h_off[i]:h_off[i + 1]]
    pad = (-len(body)) % 60
    lines = np.concatenate((body, np.full(pad, 32, dtype=np.uint8))).reshape(-1, 60)
    lines = np.concatenate((lines, np.full((lines.shape[0], 1), 10, dtype=np.uint8)), axis=1).reshape(-1)
    parts.append(np.frombuffer(b">SuperContig_%d_length_%d_ID_%d\n" % (i, len(body), i), dtype=np.uint8))
    parts.append(lines)
raw = np.concatenate(parts)
n = raw.shape[0]
dev = torch.zeros(((n + 15) // 16 * 16,), dtype=torch.uint8, device="cuda")
dev[:n] = torch.from_numpy(raw).cuda()
for _ in range(2):
    out = ops.fasta_scan_cuda(dev[:n])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    d_seq, d_off, d_hpos, odd = ops.fasta_scan_cuda(dev[:n])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
assert not odd and torch.equal(d_off, off) and torch.equal(d_seq[:int(h_off[-1])], seq[:int(h_off[-1])])
t0 = time.perf_counter()
ids, s2, o2 = fileIO.split_fasta_bytes(raw.tobytes())
host_s = time.perf_counter() - t0
assert np.array_equal(o2, h_off)
# the whole flow from a file in the page cache: mmap -> pinned staging -> device tokeniser -> count + score -> host scores
import tempfile
from phamers_b200 import pipeline
scorer = pipeline.ContigScorer()
with tempfile.NamedTemporaryFile(suffix=".fasta", delete=False) as fh:
    fh.write(raw.tobytes())
    path = fh.name
try:
    scorer.score_fasta(path)
    t0 = time.perf_counter()
    file_ids, file_scores = scorer.score_fasta(path)
    file_s = time.perf_counter() - t0
finally:
    os.unlink(path)
assert len(file_ids) == n_contigs and np.isfinite(file_scores).all()
print(json.dumps({"fasta_file_to_scores_s": file_s, "fasta_file_to_scores_GBps": n / file_s / 1e9, "file_bytes": int(n), "records": n_contigs, "bases": int(h_off[-1]), "device_ms": ms,
                  "device_file_GBps": n / ms / 1e6, "algorithmic_GBps_read_plus_write": (n + int(h_off[-1])) / ms / 1e6,
                  "host_numpy_tokeniser_s": host_s, "host_file_GBps": n / host_s / 1e9}))
