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
