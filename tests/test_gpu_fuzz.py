"""Randomised parity rounds for the counting kernels against the C oracle (reference scripts/kmer.py:42-50 restated in
oracle/kmer_oracle.c): record lengths around every threshold the kernel has (16-byte chunks, 512-byte warp steps, the 131 072-base
splitting threshold, 65 536-base tiles), arbitrary alignment of the records in the buffer, blank bytes at record edges, on tile
boundaries and in runs.  PHM_FUZZ_ROUNDS scales the number of rounds (default 4, a few seconds)."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import phamers_oracle as po

pytestmark = pytest.mark.gpu

ROUNDS = int(os.environ.get("PHM_FUZZ_ROUNDS", "4"))
BLANKS = np.frombuffer(b"Natgcn-\n\x00\xffRYKM*", dtype=np.uint8)


def _lengths(rng):
    parts = [rng.integers(0, 40, size=rng.integers(5, 40)),                                   # shorter than a few windows
             512 * rng.integers(1, 9, size=rng.integers(2, 10)) + rng.integers(-3, 4, size=1),  # around whole warp steps
             16 * rng.integers(1, 60, size=rng.integers(2, 10)) + rng.integers(-2, 3, size=1),  # around whole chunks
             rng.integers(1000, 60000, size=rng.integers(3, 12))]
    for _ in range(int(rng.integers(0, 4))):                                                  # long records: tiled
        base = int(rng.choice([131072, 196608, 262144, 65536 * int(rng.integers(2, 7))]))
        parts.append(np.array([max(0, base + int(rng.integers(-4, 5)) + int(rng.choice([0, 0, 511, 512, 513, 7])))]))
    if rng.random() < 0.5:
        parts.append(np.array([131071, 131072, 131073]))
    lengths = np.concatenate([np.asarray(p, dtype=np.int64).ravel() for p in parts])
    lengths = np.clip(lengths, 0, None)
    rng.shuffle(lengths)
    return lengths


def _workload(rng):
    lengths = _lengths(rng)
    off = np.concatenate(([0], np.cumsum(lengths)))
    skew = rng.dirichlet([2.0, 2.0, 2.0, 2.0])
    seq = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(off[-1]), p=skew)
    total = int(off[-1])
    if total:
        for _ in range(int(rng.integers(0, 60))):                                             # single blank bytes anywhere
            seq[int(rng.integers(0, total))] = rng.choice(BLANKS)
        for i in range(len(lengths)):
            lo, hi = int(off[i]), int(off[i + 1])
            if hi - lo >= 2 and rng.random() < 0.25:                                          # ... at the edges of records
                seq[lo if rng.random() < 0.5 else hi - 1] = rng.choice(BLANKS)
            if hi - lo >= 131072:                                                             # ... on and next to tile boundaries
                for t in range(1, (hi - lo) // 65536 + 1):
                    edge = ((lo + t * 65536) // 512) * 512 + int(rng.integers(-6, 7))
                    if lo <= edge < hi and rng.random() < 0.6:
                        seq[edge] = rng.choice(BLANKS)
            if hi - lo >= 3000 and rng.random() < 0.2:                                        # ... in runs (soft-masked stretches)
                a = int(rng.integers(lo, hi - 1000))
                seq[a:a + int(rng.integers(1, 900))] = rng.choice(BLANKS)
            if hi - lo >= 70000 and rng.random() < 0.2:                                       # a homopolymer: one bin takes every count
                a = int(rng.integers(lo, hi - 66000))
                seq[a:a + 66000] = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8))
    return seq, off


def _device(seq, off):
    total = int(off[-1])
    buf = torch.zeros(((total + 15) // 16 * 16 + 16,), dtype=torch.uint8)
    buf[:total] = torch.from_numpy(seq[:total].copy())
    return buf.cuda(), torch.from_numpy(np.asarray(off, dtype=np.int64)).cuda()


def _u32(t):
    return t.cpu().numpy().view(np.uint32).astype(np.int64)


@pytest.mark.parametrize("round_", range(ROUNDS))
def test_random_rounds_against_the_oracle(round_):
    from phamers_b200 import ops
    rng = np.random.default_rng(77000 + round_)
    seq, off = _workload(rng)
    d_seq, d_off = _device(seq, off)
    for k in rng.permutation([1, 2, 3, 4, 5, 6])[: 3 if round_ % 2 else 6]:
        k = int(k)
        want = c_oracle.count(seq, off, k)
        counts, freq = ops.count_cuda(d_seq, d_off, k, freq=True)
        assert np.array_equal(_u32(counts), want), (round_, k)
        assert np.array_equal(freq.cpu().numpy(), c_oracle.normalize(want), equal_nan=True), (round_, k)
        _, only_freq = ops.count_cuda(d_seq, d_off, k, counts=False, freq=True)             # split rows accumulate inside the feature rows
        assert np.array_equal(only_freq.cpu().numpy(), c_oracle.normalize(want), equal_nan=True), (round_, k)
        canon, cfreq = ops.count_cuda(d_seq, d_off, k, canonical=True, freq=True)
        want_c = po.canonical_fold(want, k)
        assert np.array_equal(_u32(canon), want_c), (round_, k, "canonical")
        assert np.array_equal(cfreq.cpu().numpy(), c_oracle.normalize(want_c), equal_nan=True), (round_, k, "canonical")


@pytest.mark.parametrize("round_", range(max(1, ROUNDS // 2)))
def test_random_rounds_fused_count_score(round_):
    """The fused call (histogram kernel emitting the scorer's operands, tiles finished by hist_finish_kernel) against the two-call
    path on the same random workloads: counts bit for bit, scores bit for bit."""
    from phamers_b200 import ops, pipeline
    rng = np.random.default_rng(88000 + round_)
    seq, off = _workload(rng)
    d_seq, d_off = _device(seq, off)
    scorer = pipeline.ContigScorer()
    want = c_oracle.count(seq, off, 4)
    counts, knn, km, combo = ops.count_score_cuda(d_seq, d_off, scorer.refs, scorer.n_positive, scorer.cent_pos, scorer.cent_neg, 3)
    assert np.array_equal(_u32(counts), want)
    plain, _ = ops.count_cuda(d_seq, d_off, 4)
    two = ops.score_cuda(plain, scorer.refs, scorer.n_positive, scorer.cent_pos, scorer.cent_neg, 3)            # int32 counts in
    for a, b in zip((knn, km, combo), two):
        assert np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True)
