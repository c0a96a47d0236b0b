"""
Drop-in for the hot-path part of the reference's scripts/phamer.py: the `phamer_scorer` object and the
`score_points(scoring_data, positive_training_data, negative_training_data, method=None)` facade (:451-468).

Scoring methods on the device: 'knn' (:268-273), 'kmeans' (:240-256) and the default 'combo' (:303-313, knn + kmeans,
range +-1.7616 -- not [-1, 1] as the reference README says).  The other methods of the reference ('dbscan', 'svm',
'density', 'silhouette') are evaluation experiments that are not data-parallel per contig and are out of scope
(SURVEY.md section 2); asking for one raises NotImplementedError.

Differences a caller can observe, all deliberate:
  * k-means on the reference sets is cached per reference set (references.py) instead of re-run on every call.
  * a query row with NaN features (contig with no countable k-mer) scores NaN instead of making scikit-learn raise.
"""
import logging
import os

import numpy as np

from . import fileIO, kmer, references

logger = logging.getLogger(__name__)
logger.setLevel(logging.WARNING)


class phamer_scorer(object):
    """scripts/phamer.py:42-449 (hot-path attributes and methods only)."""

    def __init__(self):
        self.input_directory = None
        self.features_file = None
        self.fasta_file = None
        self.output_directory = None

        self.data_ids = None
        self.data_points = None
        self.positive_ids = None
        self.positive_data = None
        self.negative_ids = None
        self.negative_data = None

        self.length_requirement = 5000                       # :68
        self.scoring_method = "combo"                        # :70
        self.all_scoring_methods = ["dbscan", "kmeans", "knn", "svm", "density", "silhouette", "combo"]   # :71
        self.kmer_length = 4                                 # :77
        self.k_clusters = 86                                 # :78
        self.k_neighbors = 3                                 # :79
        self.scores = None

    # ---- data -------------------------------------------------------------------------------------------
    def load_reference_data(self, positive_features_file=None, negative_features_file=None):
        """:110-123.  Reference sets come from feature CSVs (the reference's FASTA fallback is dead code,
        SURVEY.md 8(b)); with no files given the shipped tables are used."""
        if positive_features_file and negative_features_file:
            self.positive_ids, self.positive_data = fileIO.read_feature_file(positive_features_file, normalize=True)
            self.negative_ids, self.negative_data = fileIO.read_feature_file(negative_features_file, normalize=True)
        else:
            pid, pos, nid, neg = references.load_reference_counts()
            self.positive_ids, self.negative_ids = pid, nid
            self.positive_data, self.negative_data = kmer.normalize_counts(pos), kmer.normalize_counts(neg)

    def load_data(self, length_requirement=None):
        """:125-142: a features CSV wins over counting the FASTA; freshly counted features are cached next to the
        FASTA as <fasta>_features.csv; then normalise and screen by length."""
        if self.features_file is not None and os.path.exists(self.features_file):
            self.data_ids, self.data_points = fileIO.read_feature_file(self.features_file)
        elif self.fasta_file is not None and os.path.exists(self.fasta_file):
            self.data_ids, self.data_points = kmer.count_file(self.fasta_file, self.kmer_length, normalize=False)
            self.features_file = "{base}_features.csv".format(base=os.path.splitext(self.fasta_file)[0])
            fileIO.save_counts(self.data_points, self.data_ids, self.features_file)
        else:
            logger.error("No input fasta file or features file. Exiting...")
            raise SystemExit()
        self.data_points = kmer.normalize_counts(self.data_points)
        if length_requirement is None:
            length_requirement = self.length_requirement
        if length_requirement and self.fasta_file is not None and os.path.exists(self.fasta_file):
            self.screen_by_length(length_requirement)

    def screen_by_length(self, length_requirement=None):
        """:144-157: keep contigs whose parsed sequence has at least `length_requirement` characters."""
        if length_requirement:
            self.length_requirement = length_requirement
        ids, lengths = fileIO.get_fasta_lengths(self.fasta_file)
        long_ids = [ids[i] for i in range(len(ids)) if lengths[i] >= self.length_requirement]
        self.data_points = self.data_points[np.in1d(self.data_ids, long_ids)]
        self.data_ids = np.array(long_ids)

    def equalize_reference_data(self):
        """:159-175."""
        num_ref = min(self.positive_data.shape[0], self.negative_data.shape[0])
        self.positive_data = self.positive_data[:num_ref]
        self.negative_data = self.negative_data[:num_ref]
        if self.positive_ids is not None:
            self.positive_ids = self.positive_ids[:num_ref]
        if self.negative_ids is not None:
            self.negative_ids = self.negative_ids[:num_ref]
        self.num_positive = self.num_negative = num_ref

    # ---- scoring ----------------------------------------------------------------------------------------
    def _device_scores(self):
        import torch
        from . import ops, _lib
        _lib.require_cuda()
        pts = np.ascontiguousarray(self.data_points, dtype=np.float64)
        pos = np.ascontiguousarray(self.positive_data, dtype=np.float64)
        neg = np.ascontiguousarray(self.negative_data, dtype=np.float64)
        if self.scoring_method in ("kmeans", "combo"):
            cpos, cneg = references.reference_centroids(pos, neg, self.k_clusters)
        else:
            cpos = cneg = np.zeros((0, pts.shape[1]))
        refs = torch.from_numpy(np.vstack((pos, neg))).cuda()                         # :186
        out = ops.score_cuda(torch.from_numpy(pts).cuda(), refs, pos.shape[0],
                             torch.from_numpy(np.ascontiguousarray(cpos)).cuda(),
                             torch.from_numpy(np.ascontiguousarray(cneg)).cuda(), self.k_neighbors)
        return [t.cpu().numpy() for t in out]

    def score_points(self):
        """:177-195."""
        self.num_points = self.data_points.shape[0]
        self.num_positive = self.positive_data.shape[0]
        self.num_negative = self.negative_data.shape[0]
        self.train = np.vstack((self.positive_data, self.negative_data))
        self.labels = np.append(np.ones(self.num_positive), np.zeros(self.num_negative))
        if self.scoring_method not in self.all_scoring_methods:
            raise KeyError(self.scoring_method)
        if self.scoring_method not in ("knn", "kmeans", "combo"):
            raise NotImplementedError("scoring method %r is outside the B200 hot path (knn / kmeans / combo only)"
                                      % self.scoring_method)
        knn, km, combo = self._device_scores()
        self.scores = {"knn": knn, "kmeans": km, "combo": combo}[self.scoring_method]
        return self.scores

    def knn_score_points(self):
        return self._with_method("knn")

    def kmeans_score_points(self):
        return self._with_method("kmeans")

    def combo_score_points(self):
        return self._with_method("combo")

    def _with_method(self, method):
        saved, self.scoring_method = self.scoring_method, method
        try:
            return self.score_points()
        finally:
            self.scoring_method = saved

    # ---- outputs ----------------------------------------------------------------------------------------
    def get_phamer_output_filename(self):
        return os.path.join(self.output_directory, "phamer_scores.csv")              # :439-441

    def make_summary_file(self, args=None):
        """:316-323."""
        self.phamer_output_filename = self.get_phamer_output_filename()
        fileIO.save_phamer_scores(self.data_ids, self.scores, self.phamer_output_filename, args=args)


def score_points(scoring_data, positive_training_data, negative_training_data, method=None):
    """scripts/phamer.py:451-468."""
    scorer = phamer_scorer()
    if method is not None:
        scorer.scoring_method = method
    scorer.data_points = scoring_data
    scorer.positive_data = positive_training_data
    scorer.negative_data = negative_training_data
    return scorer.score_points()
