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


def _scoring_case(rng):
    """Random shapes and composition families: (reference counts, n_positive, query counts, centroid counts x 2, k)."""
    n_fam = int(rng.integers(1, 9))
    alpha = float(rng.choice([0.05, 0.3, 0.7, 3.0, 50.0]))
    base = rng.dirichlet(np.full(256, alpha), size=n_fam)
    if rng.random() < 0.3:
        base[0] = np.full(256, 1.0 / 256)                               # the uniform composition: degenerate centred rows

    def sample(n, depth):
        fam = rng.integers(0, n_fam, size=n)
        return np.stack([rng.multinomial(int(rng.integers(max(1, depth // 2), depth * 2)), base[f]) for f in fam]).astype(np.int64)

    depth = int(rng.choice([60, 2000, 20000, 3000000]))
    n_refs = int(rng.choice([3, 5, 127, 128, 129, 300, 1000, 2500]))
    kn = int(rng.choice([1, 3, 5]))
    n_refs = max(n_refs, kn)
    n_pos = int(rng.integers(1, n_refs)) if n_refs > 1 else 1
    n_pts = int(rng.choice([1, 31, 255, 256, 257, 600]))
    n_cp, n_cn = int(rng.choice([1, 2, 86, 130])), int(rng.choice([1, 3, 86]))
    ref_counts = sample(n_refs, depth)
    ref_counts[ref_counts.sum(axis=1) == 0, 0] = 1                      # a reference row is never empty
    if n_refs >= 8:
        dup = rng.integers(0, n_refs, size=4)
        ref_counts[dup[2:]] = ref_counts[dup[:2]]                       # duplicated references, possibly with opposite labels
    q_counts = sample(n_pts, depth)
    hit = rng.integers(0, n_pts, size=max(1, n_pts // 10))
    q_counts[hit] = ref_counts[rng.integers(0, n_refs, size=len(hit))]   # queries sitting exactly on references
    if n_pts > 3:
        q_counts[1, :] = 0                                              # an empty contig
    cen_p, cen_n = sample(n_cp, depth * 10), sample(n_cn, depth * 10)
    cen_p[cen_p.sum(axis=1) == 0, 0] = 1
    cen_n[cen_n.sum(axis=1) == 0, 0] = 1
    tag = str((alpha, depth, n_pts, n_refs, n_pos, n_cp, n_cn, kn))
    return ref_counts, n_pos, q_counts, cen_p, cen_n, kn, tag


@pytest.mark.parametrize("round_", range(max(1, ROUNDS // 2)))
def test_random_rounds_scoring_paths_agree(round_):
    """Tensor-core path (first pass, list pass, decisions) against the exhaustive float64 kernel on random shapes and random
    composition families: sparse to flat rows, shallow to very deep counts, queries that ARE references, duplicated references
    with opposite labels (ties settled by the lower reference index on every path), near-uniform rows, empty contigs."""
    from phamers_b200 import kmer, ops
    rng = np.random.default_rng(99000 + round_)
    ref_counts, n_pos, q_counts, cen_p, cen_n, kn, tag = _scoring_case(rng)
    refs = torch.from_numpy(kmer.normalize_counts(ref_counts)).cuda()
    cp = torch.from_numpy(kmer.normalize_counts(cen_p)).cuda()
    cn = torch.from_numpy(kmer.normalize_counts(cen_n)).cuda()
    counts = torch.from_numpy(q_counts.astype(np.int32)).cuda()
    feats = ops.normalize_cuda(counts)
    try:
        ops.set_score_path("tc")
        t_c = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, n_pos, cp, cn, kn)]
        t_f = [t.cpu().numpy() for t in ops.score_cuda(feats, refs, n_pos, cp, cn, kn)]
        ops.set_score_path("exact")
        e = [t.cpu().numpy() for t in ops.score_cuda(feats, refs, n_pos, cp, cn, kn)]
    finally:
        ops.set_score_path("auto")
    tag = str(round_) + " " + tag
    # Shallow counts give EXACT rational ties between different references; which of them a float64 sum calls nearer depends on
    # the order of the additions (direct differences in the decision kernels, the norm expansion in the exhaustive kernel -- and a
    # third order in scikit-learn), so the vote of such a row is not defined.  Identical reference ROWS tie bit for bit on every
    # path and stay in the comparison: the lower index wins everywhere.
    X, R = feats.cpu().numpy(), refs.cpu().numpy()
    live = ~np.isnan(X[:, 0])
    d2 = (X[live] ** 2).sum(axis=1)[:, None] + (R ** 2).sum(axis=1)[None, :] - 2.0 * X[live] @ R.T
    order = np.argsort(d2, axis=1, kind="stable")
    decided = np.ones(len(X), dtype=bool)
    if R.shape[0] > kn:
        rows = np.arange(order.shape[0])
        a, b = order[:, kn - 1], order[:, kn]
        close = np.abs(d2[rows, b] - d2[rows, a]) <= 1e-10 * (np.abs(d2[rows, a]) + 1e-30)
        same_row = (R[a] == R[b]).all(axis=1)
        decided[np.flatnonzero(live)[close & ~same_row]] = False
    for got in (t_c, t_f):
        assert np.array_equal(got[0][decided], e[0][decided], equal_nan=True), tag        # votes
        ok = ~np.isnan(e[2])
        assert np.array_equal(np.isnan(got[2]), np.isnan(e[2])), tag
        assert np.max(np.abs(got[1][ok] - e[1][ok]), initial=0.0) <= 1e-12, tag
        assert np.max(np.abs(got[2][ok & decided] - e[2][ok & decided]), initial=0.0) <= 1e-12, tag


@pytest.mark.parametrize("round_", range(max(1, ROUNDS // 2)))
def test_random_rounds_host_pipeline(round_):
    """ContigScorer.score_host (ranges of contigs uploaded on a copy stream while the previous range is counted and scored) against the
    device-resident call on the same random workloads, with ranges small enough that every call is cut: long records that fill
    several ranges' worth of bases on their own, empty records at the cuts, unaligned range starts."""
    from phamers_b200 import pipeline
    rng = np.random.default_rng(66000 + round_)
    seq, off = _workload(rng)
    scorer = pipeline.ContigScorer()
    scorer.HOST_CHUNKS = int(rng.integers(2, 9))
    scorer.HOST_CHUNK_MIN_BASES = int(rng.choice([1000, 30000, 200000]))
    d_seq, d_off = _device(seq, off)
    for method in ("combo", "knn", "kmeans")[: 1 + round_ % 3]:
        _, want = scorer.score_device(d_seq, d_off, method=method, return_counts=False)
        got = scorer.score_host(seq, off, method=method)
        assert np.array_equal(got, want.cpu().numpy(), equal_nan=True), (round_, method)
        pinned = torch.from_numpy(seq.copy()).pin_memory()
        assert np.array_equal(scorer.score_host(pinned, torch.from_numpy(off.copy()), method=method), got, equal_nan=True)
