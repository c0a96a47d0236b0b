#!/usr/bin/env python
"""
Packs the PhaMers reference feature tables (data, not code) into one compressed .npz that ships
with the package, so the drop-in scorer has its phage / bacteria reference sets on machines where
the reference checkout is absent (the GPU box).

Source: /root/reference/data/reference_features/{positive,negative}_features.csv -- raw integer
4-mer counts, 256 columns in 'ATGC' bin order, column 0 = accession (positive_features.csv:1-7).
The CSV text is parsed exactly like scripts/fileIO.py:147-152 (np.loadtxt, dtype=str, then int).

Run in the build container only:  python tools/pack_reference_features.py
"""
import os
import sys
import numpy as np

SRC = os.environ.get("PHAMERS_REF_DATA", "/root/reference/data/reference_features")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "phamers_b200", "data", "reference_features.npz")


def load(path):
    data = np.loadtxt(path, delimiter=",", dtype=str)
    return np.array(list(data[:, 0])), data[:, 1:].astype(np.int64)


def main():
    pos_ids, pos = load(os.path.join(SRC, "positive_features.csv"))
    neg_ids, neg = load(os.path.join(SRC, "negative_features.csv"))
    assert pos.shape[1] == 256 and neg.shape[1] == 256
    assert pos.max() < 2 ** 32 and neg.max() < 2 ** 32
    np.savez_compressed(DST, positive_ids=pos_ids, positive_counts=pos.astype(np.uint32),
                        negative_ids=neg_ids, negative_counts=neg.astype(np.uint32),
                        kmer_length=np.int64(4), symbols=np.array("ATGC"))
    print(DST, os.path.getsize(DST), pos.shape, neg.shape)


if __name__ == "__main__":
    sys.exit(main())
