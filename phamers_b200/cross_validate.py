"""
N-fold cross validation of the PhaMers score over the gold-standard reference features, device resident
(the reference is scripts/cross_validate.py:57-101, whose names and files this module keeps).

The reference slices the two feature matrices on the host for every fold and calls phamer.score_points twenty times.  Here
the stacked reference matrix is uploaded ONCE; a fold is two index vectors (rows held out, rows trained on) gathered on the
device, the scores of all folds are scattered into one device vector, and a single copy brings them back.  Per fold only the
k-means fits of the two training sets remain outside the scoring kernels -- they are the 2 x 1 s per fold that dominate the
reference's run -- and by default their Lloyd iterations run on the device too (references.kmeans_assign_device: same
clusterings as scikit-learn, see tests/test_gpu_kmeans.py).

    cross_validator.cross_validate()            fold labels i % N per class, drawn with the reference's two np.random.shuffle
                                                calls in the reference's order (positives, then negatives), so a seeded run
                                                scores the reference's folds
    make_metrics_file(), make_summary_file()    metrics.txt / scores.txt as scripts/cross_validate.py:155-191 writes them
    plot_*, cross_validate_all_algorithms       plots and the four non-default scoring methods: out of scope
"""
import logging
import os

import numpy as np

from . import learning, phamer, references

logger = logging.getLogger(__name__)
logger.setLevel(logging.WARNING)

_DEVICE_METHODS = ("knn", "kmeans", "combo")


class cross_validator(object):
    """Attributes as the reference's class (scripts/cross_validate.py:40-55); `scoring_function` may be replaced by any
    callable(scoring_data, positive_training, negative_training, method=...) -- the folds are then handed over on the host."""

    def __init__(self):
        self.positive_ids = self.negative_ids = None
        self.positive_data = self.negative_data = None
        self.positive_scores = self.negative_scores = None
        self.equalize_reference = False
        self.N = 20
        self.method = None
        self.scoring_function = None                 # None: the device-resident scorer below
        self.score_threshold = 0
        self.output_directory = "cross_validation"
        self.kmeans_on_device = True                 # Lloyd iterations of every fold's two k-means fits on the GPU
        self.k_clusters, self.k_neighbors = 86, 3    # scripts/phamer.py:78-79

    # -- folds -------------------------------------------------------------------------------------------
    def _equalize(self):
        """scripts/cross_validate.py:64-73: both classes cut to the size of the smaller one."""
        n_pos, n_neg = len(self.positive_data), len(self.negative_data)
        if self.equalize_reference and n_pos != n_neg:
            keep = min(n_pos, n_neg)
            self.positive_data, self.negative_data = self.positive_data[:keep], self.negative_data[:keep]
            if self.positive_ids is not None:
                self.positive_ids = self.positive_ids[:keep]
            if self.negative_ids is not None:
                self.negative_ids = self.negative_ids[:keep]
        self.num_positive, self.num_negative = len(self.positive_data), len(self.negative_data)

    def _draw_folds(self):
        """Fold label of every row: i % N shuffled, positives drawn first (scripts/cross_validate.py:75-79)."""
        labels = []
        for size in (self.num_positive, self.num_negative):
            fold_of = np.arange(size) % self.N
            np.random.shuffle(fold_of)
            labels.append(fold_of)
        return labels

    def cross_validate(self):
        """Returns (positive_scores, negative_scores): every gold-standard row scored by the model trained without its fold."""
        self._equalize()
        fold_pos, fold_neg = self._draw_folds()
        if self.scoring_function is not None or (self.method or "combo") not in _DEVICE_METHODS:
            scores = self._host_folds(fold_pos, fold_neg)
        else:
            scores = self._device_folds(fold_pos, fold_neg)
        self.positive_scores, self.negative_scores = scores[:self.num_positive], scores[self.num_positive:]
        logger.info("%d-fold cross validation complete." % self.N)
        return self.positive_scores, self.negative_scores

    def _device_folds(self, fold_pos, fold_neg):
        import torch
        from . import _lib, ops
        _lib.require_cuda()
        method = self.method or "combo"
        n_pos = self.num_positive
        host_rows = np.ascontiguousarray(np.vstack((self.positive_data, self.negative_data)), dtype=np.float64)
        rows = torch.from_numpy(host_rows).cuda()                                  # the only upload of features
        fold_of = torch.from_numpy(np.concatenate((fold_pos, fold_neg))).cuda()
        is_pos = torch.arange(rows.shape[0], device="cuda") < n_pos
        out = torch.full((rows.shape[0],), float("nan"), dtype=torch.float64, device="cuda")
        empty = torch.empty((0, rows.shape[1]), dtype=torch.float64, device="cuda")
        saved = references.kmeans_on_device
        references.kmeans_on_device = bool(self.kmeans_on_device) or saved
        try:
            for fold in range(self.N):
                logger.info("Iteration %d/%d" % (1 + fold, self.N))
                held = fold_of == fold
                query_idx = torch.nonzero(held).squeeze(1)                        # ascending: positives, then negatives
                train_idx = torch.nonzero(~held).squeeze(1)                       # likewise -> positives first (scripts/phamer.py:186)
                n_train_pos = int((is_pos & ~held).sum().item())
                cent_pos = cent_neg = empty
                if method != "knn":
                    cp, cn = references.reference_centroids(host_rows[:n_pos][fold_pos != fold], host_rows[n_pos:][fold_neg != fold],
                                                            self.k_clusters)
                    cent_pos = torch.from_numpy(np.ascontiguousarray(cp)).cuda()
                    cent_neg = torch.from_numpy(np.ascontiguousarray(cn)).cuda()
                knn, kmeans, combo = ops.score_cuda(rows.index_select(0, query_idx), rows.index_select(0, train_idx), n_train_pos,
                                                    cent_pos, cent_neg, self.k_neighbors)
                out.index_copy_(0, query_idx, {"knn": knn, "kmeans": kmeans, "combo": combo}[method])
        finally:
            references.kmeans_on_device = saved
        return out.cpu().numpy()

    def _host_folds(self, fold_pos, fold_neg):
        """A caller-supplied scoring function (or a method the device path does not cover): one call per fold with host arrays."""
        score = self.scoring_function or phamer.score_points
        out = np.zeros(self.num_positive + self.num_negative)
        for fold in range(self.N):
            in_p, in_n = fold_pos == fold, fold_neg == fold
            got = np.asarray(score(np.vstack((self.positive_data[in_p], self.negative_data[in_n])),
                                   self.positive_data[~in_p], self.negative_data[~in_n], method=self.method))
            out[:self.num_positive][in_p] = got[:in_p.sum()]
            out[self.num_positive:][in_n] = got[in_p.sum():]
        return out

    # -- reports ------------------------------------------------------------------------------------------
    def roc(self):
        """(false positive rate, true positive rate, area under the curve) of the cross-validated scores -- what
        plot_ROC (scripts/cross_validate.py:135-153) draws."""
        return learning.predictor_performance(self.positive_scores, self.negative_scores)

    def get_metric_filename(self):
        return os.path.join(self.output_directory, "metrics.txt")

    def get_summary_filename(self):
        return os.path.join(self.output_directory, "scores.txt")

    def make_metrics_file(self):
        """scripts/cross_validate.py:155-171: a title line, then the metric Series as tab-separated text."""
        table = learning.get_predictor_metrics(self.positive_scores, self.negative_scores, threshold=self.score_threshold)
        with open(self.get_metric_filename(), "w") as handle:
            handle.write("# Cross Validation Performance Metrics\n")
            table.to_csv(handle, sep="\t")

    def make_summary_file(self, id_label_map=None):
        """scripts/cross_validate.py:173-191: the positive rows, ascending by (score, id): id <tab> score [<tab> label]."""
        order = sorted(range(len(self.positive_scores)), key=lambda i: (self.positive_scores[i], self.positive_ids[i]))
        rows = ["# Cross Validation Scores"]
        for i in order:
            fields = [str(self.positive_ids[i]), str(self.positive_scores[i])]
            if id_label_map is not None:
                fields.append(str(id_label_map[self.positive_ids[i]]))
            rows.append("\t".join(fields))
        with open(self.get_summary_filename(), "w") as handle:
            handle.write("\n".join(rows))


def main(argv=None):
    """The reference's command line (scripts/cross_validate.py:240-300) without the plots."""
    import argparse
    from . import fileIO, kmer
    parser = argparse.ArgumentParser(description="N-fold cross validation of the Phamer scoring algorithm (B200 path)",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("-pf", "--positive_features_file", help="Positive features file")
    parser.add_argument("-nf", "--negative_features_file", help="Negative features file")
    parser.add_argument("-out", "--output_directory", help="Output directory")
    parser.add_argument("-N", "--N_fold", default=20, type=int, help="Number of iteration in N-fold cross validation")
    parser.add_argument("-m", "--method", default="combo", help="Scoring algorithm method")
    parser.add_argument("-equal", "--equalize_reference", action="store_true", help="Use same number of reference data from each")
    parser.add_argument("--host_kmeans", action="store_true", help="k-means fits of every fold on the host (scikit-learn), as the reference")
    args = parser.parse_args(argv)
    run = cross_validator()
    run.kmeans_on_device = not args.host_kmeans
    run.method, run.N, run.output_directory, run.equalize_reference = args.method, args.N_fold, args.output_directory, args.equalize_reference
    run.positive_ids, positive_counts = fileIO.read_feature_file(args.positive_features_file)
    run.negative_ids, negative_counts = fileIO.read_feature_file(args.negative_features_file)
    run.positive_data, run.negative_data = kmer.normalize_counts(positive_counts), kmer.normalize_counts(negative_counts)
    run.cross_validate()
    os.makedirs(run.output_directory, exist_ok=True)
    run.make_metrics_file()
    run.make_summary_file()
    print("ROC AUC = %.4f" % run.roc()[2])


if __name__ == "__main__":
    main()
