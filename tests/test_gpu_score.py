"""GPU parity tests for stage 3 (kNN vote + nearest-centroid proximity + combo) against the reference's golden
scores.  Tolerance from BASELINE.json: |dscore| <= 1e-5 with identical sign and identical kNN vote."""
import os

import numpy as np
import pytest

from oracle import phamers_oracle as po

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def scoring(golden_dir):
    g = np.load(os.path.join(golden_dir, "scoring_golden.npz"))
    from phamers_b200 import kmer, references
    _, pos_c, _, neg_c = references.load_reference_counts()
    n_ref = int(g["n_ref"])
    pos = kmer.normalize_counts(pos_c[:n_ref])
    neg = kmer.normalize_counts(neg_c[:n_ref])
    return g, pos, neg


def test_scores_match_reference_golden(scoring, monkeypatch):
    from phamers_b200 import kmer, phamer, references
    g, pos, neg = scoring
    pts = kmer.normalize_counts(g["query_counts"])
    # golden centroids = what the reference's own scikit-learn call produced when the fixture was made
    monkeypatch.setattr(references, "reference_centroids", lambda p, n, k=86: (g["centroids_pos"], g["centroids_neg"]))
    knn = phamer.score_points(pts, pos, neg, method="knn")
    assert np.array_equal(knn, g["scores_knn"])                       # identical vote
    km = phamer.score_points(pts, pos, neg, method="kmeans")
    assert np.max(np.abs(km - g["scores_kmeans"])) <= TOL
    combo = phamer.score_points(pts, pos, neg)
    assert combo.dtype == np.float64 and combo.shape == (len(pts),)
    assert np.max(np.abs(combo - g["scores_combo"])) <= TOL
    assert np.array_equal(np.sign(combo), np.sign(g["scores_combo"]))
    assert np.array_equal(combo >= 0, g["scores_combo"] >= 0)         # classification threshold (analysis.py:112)


def test_scores_with_host_kmeans(scoring):
    """End to end as a user runs it: centroids from this machine's scikit-learn (cached), still within tolerance."""
    from phamers_b200 import kmer, phamer
    g, pos, neg = scoring
    pts = kmer.normalize_counts(g["query_counts"])
    combo = phamer.score_points(pts, pos, neg)
    assert np.max(np.abs(combo - g["scores_combo"])) <= TOL
    again = phamer.score_points(pts, pos, neg)
    assert np.array_equal(combo, again)


def test_scorer_object_and_learning_knn(scoring):
    from phamers_b200 import kmer, learning, phamer
    g, pos, neg = scoring
    pts = kmer.normalize_counts(g["query_counts"][:64])
    scorer = phamer.phamer_scorer()
    assert (scorer.scoring_method, scorer.kmer_length, scorer.k_clusters, scorer.k_neighbors) == ("combo", 4, 86, 3)
    scorer.data_points, scorer.positive_data, scorer.negative_data = pts, pos, neg[:-100]
    scorer.equalize_reference_data()
    assert scorer.positive_data.shape == scorer.negative_data.shape
    scorer.scoring_method = "svm"
    with pytest.raises(NotImplementedError):
        scorer.score_points()
    train = np.vstack((pos, neg))
    labels = np.append(np.ones(len(pos)), np.zeros(len(neg)))
    votes = learning.knn(pts, train, labels, k=3)
    assert np.array_equal(votes, g["scores_knn"][:64])
    for k in (1, 5):
        assert np.array_equal(learning.knn(pts, train, labels, k=k), po.knn_scores(pts, train, labels, k=k))
    # shuffled label order: positives are regrouped internally
    perm = np.random.default_rng(0).permutation(len(train))
    assert np.array_equal(learning.knn(pts, train[perm], labels[perm], k=3), g["scores_knn"][:64])


def test_random_references_match_oracle():
    from phamers_b200 import kmer, phamer
    rng = np.random.default_rng(42)
    pos = kmer.normalize_counts(rng.integers(1, 80, size=(300, 256)))
    neg = kmer.normalize_counts(rng.integers(1, 80, size=(257, 256)) + (rng.integers(0, 30, size=256))[None, :])
    pts = kmer.normalize_counts(rng.integers(1, 80, size=(333, 256)))
    for method in ("knn", "kmeans", "combo"):
        want = po.score_points(pts, pos, neg, method=method)
        got = phamer.score_points(pts, pos, neg, method=method)
        assert np.max(np.abs(got - want)) <= TOL, method
        assert np.array_equal(np.sign(got), np.sign(want))


def test_nan_rows_score_nan(scoring):
    from phamers_b200 import phamer
    g, pos, neg = scoring
    pts = np.full((3, 256), 1.0 / 256)
    pts[1, :] = np.nan
    out = phamer.score_points(pts, pos, neg)
    assert np.isnan(out[1]) and np.isfinite(out[0]) and np.isfinite(out[2])


def test_tensor_core_path_agrees_with_exact_path(scoring):
    """The tcgen05 shortlist + float64 re-rank must give the same votes and the same scores (to rounding) as the
    exhaustive float64 kernel, on the golden queries and on a larger mixed set; the margin proof should rarely fail."""
    import torch
    from phamers_b200 import _lib, kmer, ops, references
    g, pos, neg = scoring
    rng = np.random.default_rng(3)
    _, pos_c, _, neg_c = references.load_reference_counts()
    both = np.vstack((pos_c, neg_c)).astype(np.float64)
    rows = []
    for _ in range(3000):
        row = both[int(rng.integers(0, both.shape[0]))]
        rows.append(rng.multinomial(int(rng.choice([3000, 16000, 100000])), row / row.sum()))
    rows.extend(list(rng.integers(0, 120, size=(1500, 256))))
    rows.extend(list(g["query_counts"]))
    counts = np.stack(rows).astype(np.int64)
    counts[7, :] = 0                                                      # one empty contig -> NaN row
    pts = torch.from_numpy(kmer.normalize_counts(counts)).cuda()
    refs = torch.from_numpy(np.vstack((pos, neg))).cuda()
    cp = torch.from_numpy(np.ascontiguousarray(g["centroids_pos"])).cuda()
    cn = torch.from_numpy(np.ascontiguousarray(g["centroids_neg"])).cuda()
    try:
        ops.set_score_path("exact")
        e_knn, e_km, e_combo = [t.cpu().numpy() for t in ops.score_cuda(pts, refs, len(pos), cp, cn, 3)]
        ops.set_score_path("tc")
        _lib.set_option("score_stats", 1)
        t_knn, t_km, t_combo = [t.cpu().numpy() for t in ops.score_cuda(pts, refs, len(pos), cp, cn, 3)]
        stats = ops.score_stats()
    finally:
        _lib.set_option("score_stats", 0)
        ops.set_score_path("auto")
    assert np.isnan(t_combo[7]) and np.isnan(e_combo[7])
    ok = ~np.isnan(e_combo)
    assert np.array_equal(t_knn[ok], e_knn[ok])
    assert np.max(np.abs(t_km[ok] - e_km[ok])) <= 1e-12
    assert np.max(np.abs(t_combo[ok] - e_combo[ok])) <= 1e-12
    n_gold = len(g["query_counts"])
    assert np.array_equal(t_knn[-n_gold:], g["scores_knn"])
    assert np.max(np.abs(t_combo[-n_gold:] - g["scores_combo"])) <= TOL
    print("tensor-core path on %d rows: %s" % (len(counts), stats))
    assert stats["fallback_rows"] <= len(counts) // 10
    assert 0.0 < stats["max_bound_usage"] < 0.9              # true ranking values stay inside the proven intervals


def test_scoring_from_counts_equals_scoring_from_features(scoring):
    """phm_score_counts (stages 2 + 3 fused: count / row total formed inside the kernels) must give bit-identical results
    to phm_normalize_counts followed by phm_score, including the NaN of an empty contig."""
    import torch
    from phamers_b200 import ops
    g, pos, neg = scoring
    rng = np.random.default_rng(11)
    counts = np.vstack((g["query_counts"], rng.integers(0, 90, size=(700, 256)))).astype(np.int64)
    counts[5, :] = 0
    d_counts = torch.from_numpy(counts.astype(np.int32)).cuda()
    refs = torch.from_numpy(np.vstack((pos, neg))).cuda()
    cp = torch.from_numpy(np.ascontiguousarray(g["centroids_pos"])).cuda()
    cn = torch.from_numpy(np.ascontiguousarray(g["centroids_neg"])).cuda()
    feats = ops.normalize_cuda(d_counts)
    a = [t.cpu().numpy() for t in ops.score_cuda(feats, refs, len(pos), cp, cn, 3)]
    b = [t.cpu().numpy() for t in ops.score_cuda(d_counts, refs, len(pos), cp, cn, 3)]
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    assert np.isnan(b[2][5])
    n_gold = len(g["query_counts"])
    assert np.array_equal(b[0][:n_gold][np.arange(n_gold) != 5], g["scores_knn"][np.arange(n_gold) != 5])
    # shapes the tensor cores do not take are refused, not silently handled elsewhere
    with pytest.raises(Exception):
        ops.score_cuda(d_counts, refs, len(pos), cp, cn, 7)


def test_full_path_properties_at_scale(scoring):
    """BASELINE configs[1] at its full size (1 M synthetic contigs, 16 Gbases, one B200): properties that
    do not need an oracle run of the whole thing -- shard invariance (the multi-GPU layout must not change a single score),
    the tensor-core path against the exhaustive float64 kernel on a sample of rows, the kNN term being exactly +-1, the
    centroid term inside tanh's range for this metric, and a small prefix against the CPU oracle end to end."""
    import torch
    from phamers_b200 import ops, pipeline, references
    from oracle import c_oracle
    g, pos, neg = scoring
    cents = (np.ascontiguousarray(g["centroids_pos"]), np.ascontiguousarray(g["centroids_neg"]))
    scorer = pipeline.ContigScorer(pos, neg, centroids=cents)
    n = 1000000
    seq, off = ops.synth_contigs(20260101, 0, n)
    counts, combo = scorer.score_device(seq, off)
    assert ops.score_stats()["fallback_rows"] <= 100                # read before another path reuses the workspace
    knn, km, combo2 = ops.score_cuda(counts, scorer.refs, scorer.n_positive, scorer.cent_pos, scorer.cent_neg, 3)
    assert torch.equal(combo, combo2)
    assert bool(((knn == 1.0) | (knn == -1.0)).all())
    assert float(km.abs().max()) <= np.tanh(1.0) + 1e-12
    assert torch.equal(combo, knn + km)

    # shard invariance: three uneven shards scored separately give the same bits as the whole
    pieces = []
    for lo, hi in ((0, 270001), (270001, 270002), (270002, n)):
        o = (off[lo:hi + 1] - off[lo]).contiguous()
        s = seq[int(off[lo]):int(off[hi])]
        if s.data_ptr() % 16:                                       # the C-ABI wants a 16-byte aligned sequence buffer,
            buf = torch.empty(((s.numel() + 15) // 16 * 16 + 16,), dtype=torch.uint8, device="cuda")   # readable to the next multiple of 16
            buf[:s.numel()] = s
            s = buf[:s.numel()]
        pieces.append(scorer.score_device(s, o)[1])
    assert torch.equal(torch.cat(pieces), combo)

    # a sample of rows through the exhaustive float64 kernel
    idx = torch.from_numpy(np.random.default_rng(5).choice(n, size=3000, replace=False)).cuda()
    feats = ops.normalize_cuda(counts[idx].contiguous())
    try:
        ops.set_score_path("exact")
        e_knn, e_km, e_combo = ops.score_cuda(feats, scorer.refs, scorer.n_positive, scorer.cent_pos, scorer.cent_neg, 3)
    finally:
        ops.set_score_path("auto")
    assert torch.equal(e_knn, knn[idx])
    assert float((e_combo - combo[idx]).abs().max()) <= 1e-12

    # the 10 000-contig prefix of SURVEY.md 8(d) config 2 against the CPU oracle, end to end (bit-exact counts from the C
    # restatement, scores within the stated tolerance from the numpy / scikit-learn restatement)
    m = 10000
    end = int(off[m].item())
    want_counts = c_oracle.count(seq[:end].cpu().numpy(), off[:m + 1].cpu().numpy(), 4)
    assert np.array_equal(counts[:m].cpu().numpy().view(np.uint32).astype(np.int64), want_counts)
    want = po.score_points(po.normalize_counts(want_counts), pos, neg, centroids=cents)
    got = combo[:m].cpu().numpy()
    assert np.max(np.abs(got - want)) <= TOL and np.array_equal(np.sign(got), np.sign(want))


@pytest.mark.parametrize("n_rows", [300, 2500])
def test_fallback_kernels_match_exact_path(scoring, n_rows):
    """Both exhaustive fallback kernels (column slices across CTAs for few rows, 32 rows per CTA sharing the reference stream
    for many) against the exhaustive float64 scorer: option score_force_fallback makes the tensor-core epilogue keep no candidates, so
    every row goes down the fallback road."""
    import torch
    from phamers_b200 import _lib, ops, references
    g, pos, neg = scoring
    rng = np.random.default_rng(n_rows)
    _, pos_c, _, neg_c = references.load_reference_counts()
    both = np.vstack((pos_c, neg_c)).astype(np.float64)
    rows = []
    for _ in range(n_rows - 1):
        src = both[int(rng.integers(0, both.shape[0]))]
        rows.append(rng.multinomial(12000, src / src.sum()))
    rows.append(np.zeros(256, dtype=np.int64))                            # one empty contig
    counts = torch.from_numpy(np.stack(rows).astype(np.int32)).cuda()
    refs = torch.from_numpy(np.vstack((pos, neg))).cuda()
    cp = torch.from_numpy(np.ascontiguousarray(g["centroids_pos"])).cuda()
    cn = torch.from_numpy(np.ascontiguousarray(g["centroids_neg"])).cuda()
    try:
        _lib.set_option("score_force_fallback", 1)
        _lib.set_option("score_list_pass", 0)                            # otherwise the list pass would take these rows (see below)
        f_knn, f_km, f_combo = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, len(pos), cp, cn, 3)]
        assert ops.score_stats()["fallback_rows"] == n_rows - 1          # the NaN row never asks for a fallback
        # same rows through the list pass: no candidate kept means an infinite threshold, i.e. every reference is listed
        _lib.set_option("score_list_pass", 1)
        l_knn, l_km, l_combo = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, len(pos), cp, cn, 3)]
        st = ops.score_stats()
        # the pool holds 2048 listed references per possible row (at least 4 M): 299 x 4510 fit, 2499 x 4510 do not, and the rows
        # without room take the exhaustive kernels
        assert st["rows_listed"] + st["fallback_rows"] == n_rows - 1
        assert (st["fallback_rows"] == 0) == (n_rows == 300) and st["rows_listed"] >= 1000 * (n_rows > 300)
    finally:
        _lib.set_option("score_force_fallback", 0)
        _lib.set_option("score_list_pass", 1)
    t_knn, t_km, t_combo = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, len(pos), cp, cn, 3)]
    try:
        ops.set_score_path("exact")
        e_knn, e_km, e_combo = [t.cpu().numpy() for t in ops.score_cuda(ops.normalize_cuda(counts), refs, len(pos), cp, cn, 3)]
    finally:
        ops.set_score_path("auto")
    ok = ~np.isnan(e_combo)
    assert ok.sum() == n_rows - 1 and np.isnan(f_combo[-1]) and np.isnan(t_combo[-1])
    assert np.array_equal(f_knn[ok], e_knn[ok]) and np.array_equal(t_knn[ok], e_knn[ok]) and np.array_equal(l_knn[ok], e_knn[ok])
    assert np.array_equal(l_combo[ok], f_combo[ok])
    assert np.max(np.abs(f_km[ok] - e_km[ok])) <= 1e-12
    assert np.array_equal(f_km[ok], t_km[ok])                            # centroid term: same arithmetic on every path


def test_list_pass_settles_overflowed_rows(scoring):
    """Reference sets that are clouds of near-identical rows with MIXED labels overflow the 10-slot candidate buffers of the
    first tensor-core pass.  Such rows have a final threshold, so a second pass lists every reference under it (count, prefix
    sum, fill: lists of any length share one pool -- the 2000-row cloud gives lists far longer than the others) and the exact
    decision is taken over the list (rows_listed).  With the list pass switched off the same rows take the exhaustive kernels.
    Every road must give the exhaustive float64 scorer's votes and scores."""
    import torch
    from phamers_b200 import _lib, kmer, ops, references
    g, pos, neg = scoring
    rng = np.random.default_rng(42)
    _, pos_c, _, neg_c = references.load_reference_counts()
    both = np.vstack((pos_c, neg_c)).astype(np.float64)
    src_a = both[rng.choice(both.shape[0], size=40, replace=False)]
    src_b = both[int(rng.integers(0, both.shape[0]))]
    # row totals differ from row to row: with one common total the distances are quantised and tie exactly at the k-th place
    cloud_a = [rng.multinomial(int(rng.integers(15000, 25000)), s / s.sum()) for s in src_a for _ in range(250)]
    cloud_b = [rng.multinomial(int(rng.integers(40000000, 60000000)), src_b / src_b.sum()) for _ in range(2000)]
    ref_counts = np.stack(cloud_a + cloud_b).astype(np.int64)
    ref_counts = ref_counts[rng.permutation(len(ref_counts))]             # labels (first half positive) mixed inside every cloud
    refs = torch.from_numpy(kmer.normalize_counts(ref_counts)).cuda()
    n_pos = len(ref_counts) // 2
    queries = [rng.multinomial(int(rng.integers(15000, 25000)), s / s.sum()) for s in src_a for _ in range(15)]
    queries += [rng.multinomial(int(rng.integers(40000000, 60000000)), src_b / src_b.sum()) for _ in range(50)]
    queries += list(g["query_counts"])
    counts = torch.from_numpy(np.stack(queries).astype(np.int32)).cuda()
    cp = torch.from_numpy(np.ascontiguousarray(g["centroids_pos"])).cuda()
    cn = torch.from_numpy(np.ascontiguousarray(g["centroids_neg"])).cuda()
    try:
        ops.set_score_path("tc")
        t_knn, t_km, t_combo = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, n_pos, cp, cn, 3)]
        stats = ops.score_stats()
        k5 = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, n_pos, cp, cn, 5)]
        k1 = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, n_pos, cp, cn, 1)]
        _lib.set_option("score_list_pass", 0)
        n_knn, n_km, n_combo = [t.cpu().numpy() for t in ops.score_cuda(counts, refs, n_pos, cp, cn, 3)]
        stats_off = ops.score_stats()
        ops.set_score_path("exact")
        e_knn, e_km, e_combo = [t.cpu().numpy() for t in ops.score_cuda(ops.normalize_cuda(counts), refs, n_pos, cp, cn, 3)]
        e5 = [t.cpu().numpy() for t in ops.score_cuda(ops.normalize_cuda(counts), refs, n_pos, cp, cn, 5)]
        e1 = [t.cpu().numpy() for t in ops.score_cuda(ops.normalize_cuda(counts), refs, n_pos, cp, cn, 1)]
    finally:
        _lib.set_option("score_list_pass", 1)
        ops.set_score_path("auto")
    print("list pass on: %s   off: %s" % (stats, stats_off))
    assert stats["rows_listed"] >= 50 and stats_off["rows_listed"] == 0
    assert stats["fallback_rows"] == 0
    assert stats["rows_listed"] + stats["fallback_rows"] == stats_off["fallback_rows"]
    assert np.array_equal(t_knn, e_knn) and np.array_equal(n_knn, e_knn)
    assert np.array_equal(k5[0], e5[0]) and np.array_equal(k1[0], e1[0])
    assert np.max(np.abs(t_combo - e_combo)) <= 1e-12 and np.max(np.abs(n_combo - e_combo)) <= 1e-12
    assert np.array_equal(t_km, n_km)


def test_shapes_around_every_tile_boundary():
    """Tensor-core path (first pass, list pass, exact decisions) against the exhaustive float64 kernel over shapes that sit on
    and next to every tile size: 128-row reference tiles, 256-row contig tiles, tiny reference sets (fewer tiles than list-pass
    slices), one centroid per class, more than a tile of centroids, and k = 1 / 3 / 5; queries are counts or features."""
    import torch
    from phamers_b200 import kmer, ops
    rng = np.random.default_rng(2026)
    base = rng.dirichlet(np.full(256, 0.7), size=6)                    # a few composition families: dense neighbourhoods
    def sample(n, depth):
        fam = rng.integers(0, len(base), size=n)
        return np.stack([rng.multinomial(int(rng.integers(depth // 2, depth * 2)), base[f]) for f in fam]).astype(np.int64)
    shapes = [(1, 5, 3, 1, 1, 1), (1, 5, 5, 1, 1, 5), (255, 127, 60, 1, 2, 3), (256, 128, 64, 86, 86, 3), (257, 129, 1, 129, 3, 5),
              (700, 1000, 999, 130, 86, 1), (513, 384, 100, 86, 256, 3), (40, 4510, 2255, 86, 86, 5)]
    for n_pts, n_refs, n_pos, n_cp, n_cn, kn in shapes:
        refs = torch.from_numpy(kmer.normalize_counts(sample(n_refs, 20000))).cuda()
        cp = torch.from_numpy(kmer.normalize_counts(sample(n_cp, 200000))).cuda()
        cn = torch.from_numpy(kmer.normalize_counts(sample(n_cn, 200000))).cuda()
        q_counts = sample(n_pts, 8000)
        if n_pts > 3:
            q_counts[2, :] = 0                                          # an empty contig somewhere
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
        tag = str((n_pts, n_refs, n_pos, n_cp, n_cn, kn))
        for got in (t_c, t_f):
            assert np.array_equal(got[0], e[0], equal_nan=True), tag    # votes
            ok = ~np.isnan(e[2])
            assert np.array_equal(np.isnan(got[2]), np.isnan(e[2])), tag
            assert np.max(np.abs(got[1][ok] - e[1][ok]), initial=0.0) <= 1e-12, tag
            assert np.max(np.abs(got[2][ok] - e[2][ok]), initial=0.0) <= 1e-12, tag


def test_enlarged_reference_parity(scoring):
    """BASELINE configs[4] at its full size: 1 M synthetic reference rows (the generator bench.py uses), where the quick reject,
    the list pass and the 32 M-entry pool actually matter.  2 201 queries, 600 of them sitting exactly on reference rows: the
    tensor-core path must give the exhaustive float64 kernel's votes and scores (|d| <= 1e-12) without a single row reaching
    the exhaustive fallback, and the oracle's scores (scikit-learn brute-force kNN over the same 1 M rows, as
    scripts/learning.py:118-128 calls it) within the stated tolerance with identical signs on a 500-query subset."""
    import torch
    from phamers_b200 import kmer, ops, references
    from tools import workloads
    g, pos, neg = scoring
    n_refs = 1000000
    refs, n_pos = workloads.enlarged_references(pos, neg, n_refs)
    assert refs.shape == (n_refs, 256) and n_pos == n_refs // 2
    rng = np.random.default_rng(404)
    _, pos_c, _, neg_c = references.load_reference_counts()
    both = np.vstack((pos_c, neg_c)).astype(np.float64)
    rows = []
    for _ in range(800):                                                  # contig-like rows: shipped compositions at contig depths
        src = both[int(rng.integers(0, both.shape[0]))]
        rows.append(rng.multinomial(int(rng.choice([3000, 16000, 100000])), src / src.sum()))
    rows.extend(list(rng.integers(0, 120, size=(200, 256))))              # rows far from every reference: long candidate lists
    rows.append(np.zeros(256, dtype=np.int64))                            # an empty contig -> NaN
    feats = kmer.normalize_counts(np.stack(rows).astype(np.int64))
    seq, off = ops.synth_contigs(workloads.SEED, 0, 600)                  # config-2 contigs, counted on the device
    synth_counts, _ = ops.count_cuda(seq, off, 4)
    on_refs = refs[torch.from_numpy(rng.choice(n_refs, size=600, replace=False)).cuda()]
    pts = torch.cat((torch.from_numpy(feats).cuda(), ops.normalize_cuda(synth_counts), on_refs)).contiguous()
    n = pts.shape[0]
    assert n >= 2000
    cp = torch.from_numpy(np.ascontiguousarray(g["centroids_pos"])).cuda()
    cn = torch.from_numpy(np.ascontiguousarray(g["centroids_neg"])).cuda()
    try:
        ops.set_score_path("tc")
        t_knn, t_km, t_combo = [t.cpu().numpy() for t in ops.score_cuda(pts, refs, n_pos, cp, cn, 3)]
        stats = ops.score_stats()
        ops.set_score_path("exact")
        e_knn, e_km, e_combo = [t.cpu().numpy() for t in ops.score_cuda(pts, refs, n_pos, cp, cn, 3)]
    finally:
        ops.set_score_path("auto")
    print("enlarged reference, %d queries x %d rows: %s" % (n, n_refs, stats))
    assert stats["fallback_rows"] == 0
    nan_row = len(rows) - 1
    ok = np.arange(n) != nan_row
    assert np.isnan(t_combo[nan_row]) and np.isnan(e_combo[nan_row]) and not np.isnan(e_combo[ok]).any()
    assert np.array_equal(t_knn[ok], e_knn[ok])                           # identical votes
    assert np.max(np.abs(t_km[ok] - e_km[ok])) <= 1e-12
    assert np.max(np.abs(t_combo[ok] - e_combo[ok])) <= 1e-12
    # the oracle on a subset that covers every kind of row (scikit-learn on the host: 500 x 1 M distances)
    pick = np.concatenate((np.arange(0, 1000, 5), np.arange(1001, 1601, 6), np.arange(1601, n, 3)))[:500]
    host_refs = refs.cpu().numpy()
    want = po.score_points(pts[torch.from_numpy(pick).cuda()].cpu().numpy(), host_refs[:n_pos], host_refs[n_pos:],
                           centroids=(g["centroids_pos"], g["centroids_neg"]))
    assert np.max(np.abs(t_combo[pick] - want)) <= TOL
    assert np.array_equal(np.sign(t_combo[pick]), np.sign(want))


def test_learning_distances_and_closest_to(scoring):
    """learning.distances / learning.closest_to (reference scripts/learning.py:47-66) on the device against the oracle's numpy
    restatement: distances to every centroid, and the nearest centroid returned as the row itself."""
    from phamers_b200 import kmer, learning
    g, pos, neg = scoring
    cents = np.ascontiguousarray(g["centroids_pos"])
    pts = kmer.normalize_counts(g["query_counts"][:25])
    for p in pts:
        got = learning.distances(p, cents)
        want = po.distances(p, cents)
        assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-15
        assert np.array_equal(learning.closest_to(p, cents), po.closest_to(p, cents))
    two_d = learning.distances(pts[:1], cents)                            # a [1, dim] point is accepted as well (:53-55)
    assert np.array_equal(two_d, learning.distances(pts[0], cents))
    assert np.array_equal(learning.closest_to(cents[7], cents), cents[7])
