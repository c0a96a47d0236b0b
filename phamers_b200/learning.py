"""
Drop-in for the hot-path part of the reference's scripts/learning.py.

    knn            :118   brute-force k nearest neighbours + vote       -> CUDA (phm_score)
    distances      :47, closest_to :59   per-contig nearest-centroid search -> fused into phm_score (no host
                          equivalent is provided: there is no CPU arithmetic path in this package)
    get_centroids  :69, kmeans :131   reference-set preprocessing         -> host scikit-learn, see references.py

The evaluation helpers of learning.py (dbscan, silhouettes, get_density, ROC metrics) are outside the hot path.
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
