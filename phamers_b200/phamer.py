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
        self.data_directory = None
        self.positive_features_file = None
        self.negative_features_file = None

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
        self.data_points = self.data_points[np.isin(self.data_ids, long_ids)]
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

    # ---- files (scripts/phamer.py:405-436) -------------------------------------------------------------
    def find_data_files(self):
        """:406-415: default places of the reference feature tables inside a data directory."""
        self.positive_features_file = os.path.join(self.data_directory, "reference_features", "positive_features.csv")
        self.negative_features_file = os.path.join(self.data_directory, "reference_features", "negative_features.csv")

    def find_input_files(self):
        """:417-436: the single .fasta / .fa file of the input directory (a '*genes' file only as a last resort) and its single .csv."""
        if self.input_directory and os.path.isdir(self.input_directory):
            fasta_files = [f for f in os.listdir(self.input_directory) if f.endswith(".fasta") or f.endswith(".fa")]
            if len(fasta_files) == 1:
                self.fasta_file = os.path.join(self.input_directory, fasta_files[0])
            elif fasta_files:
                for potential_file in fasta_files:
                    if not os.path.splitext(potential_file)[0].endswith("genes"):
                        self.fasta_file = os.path.join(self.input_directory, potential_file)
                        break
                if self.fasta_file is None:
                    self.fasta_file = os.path.join(self.input_directory, fasta_files[0])
            features_files = [f for f in os.listdir(self.input_directory) if f.endswith(".csv")]
            if len(features_files) == 1:
                self.features_file = os.path.join(self.input_directory, features_files[0])

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


def decide_files(scorer, args):
    """scripts/phamer.py:470-507: explicit command-line files win over what the input / data directories hold; the output goes
    to -out, else to <input directory>/phamer_output (the reference's misspelt `input_drectory` branch is dead code)."""
    if args.kmer_length:
        scorer.kmer_length = args.kmer_length
    if args.input_directory:
        scorer.input_directory = args.input_directory
        scorer.find_input_files()
    if args.data_directory:
        scorer.data_directory = args.data_directory
        scorer.find_data_files()
    if args.output_directory:
        scorer.output_directory = args.output_directory
    else:
        base = scorer.input_directory or os.path.dirname(args.fasta_file or args.features_file or "") or "."
        scorer.output_directory = os.path.join(base, "phamer_output")
    scorer.fasta_file = args.fasta_file or scorer.fasta_file
    scorer.features_file = args.features_file or scorer.features_file
    scorer.positive_features_file = args.positive_features or scorer.positive_features_file
    scorer.negative_features_file = args.negative_features or scorer.negative_features_file


def main(argv=None):
    """The reference's command line (scripts/phamer.py:510-599) for the hot path: count the contigs of a FASTA (or read their
    feature CSV), score them against the reference features, write <out>/phamer_scores.csv.  t-SNE and plots are out of scope."""
    import argparse
    parser = argparse.ArgumentParser(description="Scores contigs based on feature similarity (B200 path)",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("-in", "--input_directory", help="Directory containing input files")
    parser.add_argument("-fasta", "--fasta_file", help="Fasta compilation file of unknown sequences")
    parser.add_argument("-features", "--features_file", help="Input feature file")
    parser.add_argument("-data", "--data_directory", help="Directory containing all relevant data files")
    parser.add_argument("-pf", "--positive_features", help="positive feature file")
    parser.add_argument("-nf", "--negative_features", help="negative feature file")
    parser.add_argument("-out", "--output_directory", help="Directory to put output files")
    parser.add_argument("-k", "--kmer_length", type=int, default=4, help="Length of k-mers analyzed")
    parser.add_argument("-l", "--length_requirement", type=int, default=5000, help="Sequence length requirement")
    parser.add_argument("-equal", "--equalize_reference", action="store_true", help="Use same number of reference data from each")
    parser.add_argument("-m", "--method", default="combo", help="Learning algorithm name")
    parser.add_argument("-do_tsne", "--do_tsne", action="store_true", help="(out of scope)")
    parser.add_argument("-plot", "--plot_tsne", action="store_true", help="(out of scope)")
    args = parser.parse_args(argv)
    if args.do_tsne or args.plot_tsne:
        raise NotImplementedError("t-SNE and plotting are outside the B200 hot path")
    scorer = phamer_scorer()
    decide_files(scorer, args)
    scorer.scoring_method = args.method
    scorer.load_reference_data(scorer.positive_features_file, scorer.negative_features_file)
    scorer.load_data(length_requirement=args.length_requirement)
    if args.equalize_reference:
        scorer.equalize_reference_data()
    if not os.path.isdir(scorer.output_directory):
        os.makedirs(scorer.output_directory)
    scorer.scores = scorer.score_points()
    scorer.make_summary_file(args=args)
    return scorer


if __name__ == "__main__":
    main()
