"""Device Lloyd iterations (phm_kmeans_lloyd) against scikit-learn's KMeans(n_clusters, random_state=10)
-- the call learning.kmeans makes (reference scripts/learning.py:138) -- label for label, and the centroids that follow."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sets():
    from phamers_b200 import kmer, references
    _, pos_c, _, neg_c = references.load_reference_counts()
    pos, neg = kmer.normalize_counts(pos_c), kmer.normalize_counts(neg_c)
    rng = np.random.default_rng(5)
    fold = rng.permutation(2255) % 20 != 3                                # a cross-validation training subset
    return [("positive", pos[:2255], 86), ("negative", neg[:2255], 86), ("fold", pos[:2255][fold], 86),
            ("small", neg[:300], 7), ("k1", pos[:50], 1)]


def test_labels_equal_scikit_learn():
    from phamers_b200 import references
    for name, data, k in _sets():
        want = references.kmeans_assign(data, k)
        got = references.kmeans_assign_device(data, k)
        assert got.shape == want.shape and np.array_equal(got, want), name
        assert np.array_equal(references.get_centroids(data, got), references.get_centroids(data, want)), name


def test_scores_with_device_kmeans_match_golden(golden_dir):
    """phamer.score_points with the clustering on the device gives the reference's golden scores."""
    import os
    from phamers_b200 import kmer, phamer, references
    g = np.load(os.path.join(golden_dir, "scoring_golden.npz"))
    pos, neg = references.load_reference_features(equalize=True)
    pts = kmer.normalize_counts(g["query_counts"])
    references.clear_cache()
    references.kmeans_on_device = True
    try:
        got = phamer.score_points(pts, pos, neg)
    finally:
        references.kmeans_on_device = False
        references.clear_cache()
    assert np.max(np.abs(got - g["scores_combo"])) <= 1e-5 and np.array_equal(np.sign(got), np.sign(g["scores_combo"]))
