"""
Drop-in for the hot-path part of the reference's scripts/learning.py.

    knn            :118   brute-force k nearest neighbours + vote       -> CUDA (phm_score)
    distances      :47, closest_to :59   per-contig nearest-centroid search -> fused into phm_score (no host
                          equivalent is provided: there is no CPU arithmetic path in this package)
    get_centroids  :69, kmeans :131   reference-set preprocessing         -> host scikit-learn, see references.py

    predictor_performance :185, get_truth_table :199, get_predictor_metrics :224   ROC and truth-table metrics (host)

The other evaluation helpers of learning.py (dbscan, silhouettes, get_density) are outside the hot path.
"""
import numpy as np

from . import references

kmeans_seed = references.KMEANS_SEED


def knn(queries, ref_data, ref_labels, k=3):
    """scripts/learning.py:118-128: 2 * (KNeighborsClassifier(k).fit(refs, labels).predict(queries) - 0.5), computed on
    the GPU.  Labels must be 0 / 1; rows are reordered so that the 1-labelled references come first."""
    import torch
    from . import ops, _lib
    _lib.require_cuda()
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    ref_data = np.asarray(ref_data, dtype=np.float64)
    ref_labels = np.asarray(ref_labels)
    if not np.all((ref_labels == 0) | (ref_labels == 1)):
        raise NotImplementedError("knn on the device supports the reference's 0/1 labels")
    order = np.argsort(ref_labels == 0, kind="stable")          # positives first, original order kept inside each class
    refs = np.ascontiguousarray(ref_data[order])
    n_pos = int(np.sum(ref_labels == 1))
    empty = torch.empty((0, queries.shape[1]), dtype=torch.float64, device="cuda")
    votes, _, _ = ops.score_cuda(torch.from_numpy(queries).cuda(), torch.from_numpy(refs).cuda(), n_pos, empty, empty, k)
    return votes.cpu().numpy()


def get_centroids(data, assignment):
    """scripts/learning.py:69-81."""
    return references.get_centroids(np.asarray(data), np.asarray(assignment))


def kmeans(data, k, verbose=False, sort_by_size=False):
    """scripts/learning.py:131-146 (host scikit-learn, random_state = kmeans_seed)."""
    if sort_by_size:
        raise NotImplementedError("sort_by_size is outside the hot path")
    return references.kmeans_assign(np.asarray(data), k)


# ---- evaluation helpers used by cross_validate.py (host side: a handful of scalars per run) --------------------------
def predictor_performance(positive_scores, negative_scores):
    """scripts/learning.py:185-196: (false positive rate, true positive rate, ROC area under the curve)."""
    from sklearn.metrics import auc, roc_curve
    truth = np.append(np.ones(len(positive_scores)), np.zeros(len(negative_scores))).astype(bool)
    predictions = np.append(positive_scores, negative_scores)
    false_positive_rate, true_positive_rate, _ = roc_curve(truth, predictions)
    return false_positive_rate, true_positive_rate, auc(false_positive_rate, true_positive_rate)


def get_truth_table(positive_scores, negative_scores, threshold=0):
    """scripts/learning.py:199-221: (TPR, FPR, FNR, TNR); a score >= threshold is a positive call."""
    tp = np.sum(positive_scores >= threshold)
    fp = np.sum(negative_scores >= threshold)
    fn = np.sum(positive_scores < threshold)
    tn = np.sum(negative_scores < threshold)
    tpr = float(tp) / (tp + fn) if tp + fn != 0 else 0
    fpr = float(fp) / (fp + tn) if fp + tn != 0 else 0
    return tpr, fpr, 1 - tpr, 1 - fpr


def get_predictor_metrics(positive_scores, negative_scores, threshold=0):
    """scripts/learning.py:224-245: pandas Series tp fp fn tn tpr fpr fnr tnr ppv npv fdr acc."""
    import pandas as pd
    metrics = ["tp", "fp", "fn", "tn", "tpr", "fpr", "fnr", "tnr", "ppv", "npv", "fdr", "acc"]
    series = pd.Series(index=metrics, dtype=float)
    series["tp"] = np.sum(positive_scores >= threshold)
    series["fp"] = np.sum(negative_scores >= threshold)
    series["fn"] = np.sum(positive_scores < threshold)
    series["tn"] = np.sum(negative_scores < threshold)
    series["tpr"], series["fpr"], series["fnr"], series["tnr"] = get_truth_table(positive_scores, negative_scores, threshold=threshold)
    series["ppv"] = float(series["tp"]) / (series["tp"] + series["fp"])
    series["npv"] = float(series["tn"]) / (series["tn"] + series["fn"])
    series["fdr"] = 1 - series["ppv"]
    series["acc"] = float(series["tp"] + series["tn"]) / (series["tp"] + series["fp"] + series["fn"] + series["tn"])
    return series
