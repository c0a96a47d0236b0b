"""Cross-validation drop-in against vectors generated from the unmodified reference
scripts/cross_validate.py (tests/golden/make_cv_golden.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-5                                       # north-star tolerance on scores


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "cross_validate_golden.npz"))


def test_metric_helpers_match_reference(golden):
    """ROC / truth-table helpers (host side) on the reference's own cross-validated scores."""
    from phamers_b200 import learning
    ps, ns = golden["positive_scores"], golden["negative_scores"]
    fpr, tpr, area = learning.predictor_performance(ps, ns)
    assert np.array_equal(fpr, golden["fpr"]) and np.array_equal(tpr, golden["tpr"]) and area == float(golden["auc"])
    metrics = learning.get_predictor_metrics(ps, ns, threshold=0)
    assert list(metrics.index) == list(golden["metric_names"])
    assert np.array_equal(metrics.values.astype(float), golden["metric_values"])
    assert learning.get_truth_table(ps, ns) == tuple(golden["metric_values"][4:8])
    assert learning.get_truth_table(np.zeros(0), np.zeros(0)) == (0, 0, 1, 1)


@pytest.mark.gpu
def test_cross_validation_matches_reference(golden, tmp_path):
    import sklearn
    from phamers_b200 import cross_validate, kmer, references
    if sklearn.__version__ != str(golden["sklearn_version"]):
        pytest.skip("k-means centroids are tied to scikit-learn %s" % golden["sklearn_version"])
    n_rows = int(golden["n_rows"])
    _, pos_c, _, neg_c = references.load_reference_counts()
    v = cross_validate.cross_validator()
    v.method = "combo"
    v.N = int(golden["n_fold"])
    v.positive_data = kmer.normalize_counts(pos_c[:n_rows])
    v.negative_data = kmer.normalize_counts(neg_c[:n_rows])
    v.positive_ids = np.array(["p%d" % i for i in range(n_rows)])
    v.negative_ids = np.array(["n%d" % i for i in range(n_rows)])
    v.output_directory = str(tmp_path)
    np.random.seed(int(golden["seed"]))
    ps, ns = v.cross_validate()
    for got, want in ((ps, golden["positive_scores"]), (ns, golden["negative_scores"])):
        assert np.max(np.abs(got - want)) <= TOL
        assert np.array_equal(np.sign(got), np.sign(want))
    assert abs(v.roc()[2] - float(golden["auc"])) <= 1e-9
    # the same folds with every k-means fit iterating on the device: identical clusterings, hence identical scores
    v2 = cross_validate.cross_validator()
    v2.method, v2.N, v2.kmeans_on_device = "combo", v.N, True
    v2.positive_data, v2.negative_data = v.positive_data, v.negative_data
    np.random.seed(int(golden["seed"]))
    ps_dev, ns_dev = v2.cross_validate()
    assert np.array_equal(ps_dev, ps) and np.array_equal(ns_dev, ns)
    v.make_metrics_file()
    v.make_summary_file()
    lines = open(v.get_metric_filename()).read().splitlines()
    assert lines[0] == "# Cross Validation Performance Metrics"
    values = dict(line.split("\t") for line in lines[2:])
    assert float(values["acc"]) == float(golden["metric_values"][-1])
    summary = open(v.get_summary_filename()).read().splitlines()
    assert summary[0] == "# Cross Validation Scores" and len(summary) == 1 + n_rows
    scores = [float(line.split("\t")[1]) for line in summary[1:]]
    assert scores == sorted(scores)
    # equalising truncates the longer class to the shorter one (scripts/cross_validate.py:64-73)
    w = cross_validate.cross_validator()
    w.method = "knn"
    w.N = 2
    w.equalize_reference = True
    w.positive_data, w.negative_data = v.positive_data[:120], v.negative_data[:90]
    w.positive_ids, w.negative_ids = v.positive_ids[:120], v.negative_ids[:90]
    ps2, ns2 = w.cross_validate()
    assert ps2.shape == (90,) and ns2.shape == (90,) and set(np.unique(np.concatenate((ps2, ns2)))) <= {-1.0, 1.0}
