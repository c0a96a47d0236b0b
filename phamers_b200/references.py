"""
Reference feature sets and their (host-side, cached) clustering.

The shipped phage / bacteria tables (reference data/reference_features/{positive,negative}_features.csv, raw 4-mer
counts in 'ATGC' order) travel with the package as data/reference_features.npz (tools/pack_reference_features.py).

Clustering of the reference sets is reference-only preprocessing, independent of the contigs being scored
(scripts/phamer.py:245-248 runs it twice per scoring call, ~1 s each).  It stays on the host with the same
scikit-learn call the reference makes (scripts/learning.py:138: KMeans(n_clusters=k, random_state=10)) so that the
centroids are the reference's centroids, and it is cached per reference set instead of being redone on every call.
"""
import hashlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_NPZ = os.path.join(_HERE, "data", "reference_features.npz")
KMEANS_SEED = 10                      # scripts/learning.py:21
_centroid_cache = {}


def load_reference_counts():
    """(positive_ids, positive_counts int64[2255, 256], negative_ids, negative_counts int64[2418, 256])."""
    z = np.load(_NPZ)
    return (z["positive_ids"], z["positive_counts"].astype(np.int64),
            z["negative_ids"], z["negative_counts"].astype(np.int64))


def load_reference_features(equalize=True):
    """Normalised reference features as phamer.load_data builds them (scripts/phamer.py:112,119: read_feature_file(...,
    normalize=True)); equalize truncates both to min(nP, nN) rows (scripts/phamer.py:159-175)."""
    from . import kmer
    _, pos, _, neg = load_reference_counts()
    if equalize:
        n = min(len(pos), len(neg))
        pos, neg = pos[:n], neg[:n]
    return kmer.normalize_counts(pos), kmer.normalize_counts(neg)


def _digest(arr):
    arr = np.ascontiguousarray(arr)
    return hashlib.sha1(arr.view(np.uint8)).hexdigest() + str(arr.shape)


def kmeans_assign(data, k):
    """scripts/learning.py:131-146 (host, scikit-learn, seed 10)."""
    from sklearn.cluster import KMeans
    return np.asarray(KMeans(n_clusters=k, random_state=KMEANS_SEED).fit(data).labels_)


def get_centroids(data, assignment):
    """scripts/learning.py:69-81."""
    labels = sorted(set(assignment) - {-1})
    return np.array([np.mean(data[assignment == c], axis=0) for c in labels])


def cluster_centroids(data, k):
    key = (_digest(data), int(k))
    hit = _centroid_cache.get(key)
    if hit is None:
        hit = get_centroids(data, kmeans_assign(data, k))
        _centroid_cache[key] = hit
    return hit


def reference_centroids(positive, negative, k_clusters=86):
    """scripts/phamer.py:245-248: centroids of k-means (k = 86) on each reference set."""
    return cluster_centroids(positive, k_clusters), cluster_centroids(negative, k_clusters)


def clear_cache():
    _centroid_cache.clear()
