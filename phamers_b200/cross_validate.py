"""
Drop-in for the scoring part of the reference's scripts/cross_validate.py: N-fold cross validation
of the PhaMers score over the gold-standard reference features, every fold scored by the CUDA path.

    reference (scripts/cross_validate.py)            here
    cross_validator.cross_validate :57-101           same folds (np.random shuffles of i % N, positives first), one
                                                     phamer.score_points call per fold -> phm_score on the device
    make_metrics_file :155, make_summary_file :173   same files (metrics.txt, scores.txt)
    plot_* :103-153, cross_validate_all_algorithms   plotting / the four non-default methods: out of scope

The k-means centroids of each fold's training sets are host scikit-learn, exactly as in phamer.score_points
(references.py); with 20 folds that preprocessing, not the scoring, is what takes the time.
"""
import logging
import os

import numpy as np

from . import learning, phamer

logger = logging.getLogger(__name__)
logger.setLevel(logging.WARNING)


class cross_validator(object):

    def __init__(self):
        self.positive_ids = None
        self.negative_ids = None
        self.positive_data = None
        self.negative_data = None
        self.positive_scores = None
        self.negative_scores = None
        self.equalize_reference = False

        self.N = 20
        self.method = None
        self.scoring_function = phamer.score_points
        self.score_threshold = 0
        self.output_directory = "cross_validation"
        self.kmeans_on_device = False                # True: the k-means fits of every fold iterate on the GPU (references.kmeans_assign_device)

    def cross_validate(self):
        """scripts/cross_validate.py:57-101.  Returns (positive_scores, negative_scores)."""
        self.num_positive = self.positive_data.shape[0]
        self.num_negative = self.negative_data.shape[0]
        if self.equalize_reference and self.num_positive != self.num_negative:
            num_ref = min(self.num_positive, self.num_negative)
            self.positive_data = self.positive_data[:num_ref]
            self.negative_data = self.negative_data[:num_ref]
            if self.positive_ids is not None:
                self.positive_ids = self.positive_ids[:num_ref]
            if self.negative_ids is not None:
                self.negative_ids = self.negative_ids[:num_ref]
            self.num_positive = num_ref
            self.num_negative = num_ref

        positive_asmt = np.arange(self.num_positive) % self.N
        negative_asmt = np.arange(self.num_negative) % self.N
        np.random.shuffle(positive_asmt)                                 # same generator, same order of draws as the reference
        np.random.shuffle(negative_asmt)

        self.positive_scores = np.zeros(self.num_positive)
        self.negative_scores = np.zeros(self.num_negative)
        from . import references
        saved_kmeans = references.kmeans_on_device
        references.kmeans_on_device = bool(self.kmeans_on_device) or saved_kmeans
        try:
            self._score_folds(positive_asmt, negative_asmt)
        finally:
            references.kmeans_on_device = saved_kmeans
        logger.info("%d-fold cross validation complete." % self.N)
        return self.positive_scores, self.negative_scores

    def _score_folds(self, positive_asmt, negative_asmt):
        for n in range(self.N):
            logger.info("Iteration %d/%d" % (1 + n, self.N))
            where_positive = (positive_asmt == n)
            where_negative = (negative_asmt == n)
            positive_sub_div_size = np.sum(where_positive)
            scoring_data = np.vstack((self.positive_data[where_positive], self.negative_data[where_negative]))
            pos_training_data = self.positive_data[np.invert(where_positive)]
            neg_training_data = self.negative_data[np.invert(where_negative)]
            scores = self.scoring_function(scoring_data, pos_training_data, neg_training_data, method=self.method)
            self.positive_scores[where_positive] = scores[:positive_sub_div_size]
            self.negative_scores[where_negative] = scores[positive_sub_div_size:]

    def roc(self):
        """(false positive rate, true positive rate, area under the curve) of the cross-validated scores
        (learning.predictor_performance, what plot_ROC :135-153 draws)."""
        return learning.predictor_performance(self.positive_scores, self.negative_scores)

    def make_metrics_file(self):
        """scripts/cross_validate.py:155-171."""
        file_name = self.get_metric_filename()
        with open(file_name, "w") as f:
            f.write("# Cross Validation Performance Metrics\n")
        metrics_series = learning.get_predictor_metrics(self.positive_scores, self.negative_scores, threshold=self.score_threshold)
        metrics_series.to_csv(file_name, sep="\t", mode="a")

    def make_summary_file(self, id_label_map=None):
        """scripts/cross_validate.py:173-191: positive ids and scores, ascending by score."""
        text = "# Cross Validation Scores"
        pairs = sorted(zip(self.positive_scores, self.positive_ids))
        for score, id in pairs:
            if id_label_map is None:
                text += "\n{id}\t{score}".format(id=id, score=score)
            else:
                text += "\n{id}\t{score}\t{label}".format(id=id, score=score, label=id_label_map[id])
        with open(self.get_summary_filename(), "w") as f:
            f.write(text)

    def get_metric_filename(self):
        return os.path.join(self.output_directory, "metrics.txt")

    def get_summary_filename(self):
        return os.path.join(self.output_directory, "scores.txt")


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
    parser.add_argument("--device_kmeans", action="store_true", help="Lloyd iterations of every fold's k-means on the GPU")
    args = parser.parse_args(argv)
    validator = cross_validator()
    validator.kmeans_on_device = args.device_kmeans
    validator.method = args.method
    validator.N = args.N_fold
    validator.output_directory = args.output_directory
    validator.positive_ids, positive_data = fileIO.read_feature_file(args.positive_features_file)
    validator.negative_ids, negative_data = fileIO.read_feature_file(args.negative_features_file)
    validator.positive_data = kmer.normalize_counts(positive_data)
    validator.negative_data = kmer.normalize_counts(negative_data)
    validator.equalize_reference = args.equalize_reference
    validator.cross_validate()
    os.makedirs(validator.output_directory, exist_ok=True)
    validator.make_metrics_file()
    validator.make_summary_file()
    print("ROC AUC = %.4f" % validator.roc()[2])


if __name__ == "__main__":
    main()
