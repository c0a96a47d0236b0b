#!/usr/bin/env python
"""
Golden vectors for the cross-validation drop-in, generated from the UNMODIFIED reference
scripts/cross_validate.py (imported through oracle/ref_loader.py) in the build container:

    python tests/golden/make_cv_golden.py      ->  tests/golden/cross_validate_golden.npz

4-fold cross validation of the combo score over the first 500 phage and 500 bacteria rows of the shipped reference
features, numpy's global generator seeded with 777 just before the call.  Each fold runs the reference's own k-means
(scikit-learn, random_state 10) on its training sets, so the vectors are tied to the scikit-learn of this image (1.9.0).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import phamers_oracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_ROWS, N_FOLD, SEED = 500, 4, 777


def main():
    import sklearn
    cv = ref_loader.load_reference_cross_validate()
    kmer, learning, phamer = ref_loader.load_reference()
    data_dir = os.path.join(os.path.dirname(ref_loader.reference_scripts_dir()), "data", "reference_features")
    _, pos_counts = po.read_feature_file(os.path.join(data_dir, "positive_features.csv"))
    _, neg_counts = po.read_feature_file(os.path.join(data_dir, "negative_features.csv"))
    pos = kmer.normalize_counts(pos_counts[:N_ROWS])
    neg = kmer.normalize_counts(neg_counts[:N_ROWS])
    v = cv.cross_validator()
    v.scoring_function = phamer.score_points
    v.method = "combo"
    v.N = N_FOLD
    v.positive_data, v.negative_data = pos, neg
    v.positive_ids = np.array(["p%d" % i for i in range(N_ROWS)])
    v.negative_ids = np.array(["n%d" % i for i in range(N_ROWS)])
    np.random.seed(SEED)
    ps, ns = v.cross_validate()
    fpr, tpr, area = learning.predictor_performance(ps, ns)
    metrics = learning.get_predictor_metrics(ps, ns, threshold=0)
    np.savez_compressed(os.path.join(HERE, "cross_validate_golden.npz"),
                        n_rows=N_ROWS, n_fold=N_FOLD, seed=SEED, positive_scores=ps, negative_scores=ns,
                        fpr=fpr, tpr=tpr, auc=area, metric_names=np.array(list(metrics.index)),
                        metric_values=metrics.values.astype(float), sklearn_version=sklearn.__version__)
    print("positive mean %.4f negative mean %.4f  AUC %.4f  acc %.4f" % (ps.mean(), ns.mean(), area, metrics["acc"]))


if __name__ == "__main__":
    main()
