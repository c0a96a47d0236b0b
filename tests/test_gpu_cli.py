"""BASELINE configs[0] end to end through the drop-in command line: a directory holding one multi-record FASTA (60-column
lines, '_ID_' headers, N runs and lower-case runs) -> phamer_scores.csv, against the oracle."""
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import phamers_oracle as po

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _write_fasta(path, rng, n_contigs):
    lengths = np.clip(np.round(np.exp(rng.normal(np.log(10000), 1.0, size=n_contigs))), 1000, 100000).astype(np.int64)
    seqs = []
    with open(path, "w") as fh:
        for i, length in enumerate(lengths):
            gc = rng.uniform(0.25, 0.75)
            p = np.array([(1 - gc) / 2, (1 - gc) / 2, gc / 2, gc / 2])
            s = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(length), p=p)
            for alphabet in (b"N", b"atgc"):                            # 1 % masked runs of each kind
                for _ in range(max(1, int(length) // 1600)):
                    a = int(rng.integers(0, length))
                    s[a:a + 16] = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=len(s[a:a + 16]))
            text = s.tobytes().decode()
            seqs.append(text)
            fh.write(">SuperContig_%d_length_%d_ID_%d\n" % (i, length, i))
            fh.write("\n".join(text[j:j + 60] for j in range(0, len(text), 60)) + "\n")
    return lengths, seqs


def test_fasta_directory_to_score_file(tmp_path):
    from phamers_b200 import fileIO, phamer, references
    rng = np.random.default_rng(20260101)
    indir = tmp_path / "sample"
    indir.mkdir()
    lengths, seqs = _write_fasta(indir / "contigs.fasta", rng, 300)
    scorer = phamer.main(["-in", str(indir), "-out", str(tmp_path / "out"), "-equal", "-l", "5000"])
    # oracle: the same contigs counted by the C restatement, scored by the numpy / scikit-learn restatement
    blob = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
    offsets = np.concatenate(([0], np.cumsum([len(s) for s in seqs]))).astype(np.int64)
    counts = c_oracle.count(blob, offsets, 4)
    keep = lengths >= 5000
    pos, neg = references.load_reference_features(equalize=True)
    cents = references.reference_centroids(pos, neg)
    want = po.score_points(po.normalize_counts(counts[keep]), pos, neg, centroids=cents)
    by_id = fileIO.read_phamer_output(str(tmp_path / "out" / "phamer_scores.csv"))       # {contig id: score}, scripts/fileIO.py:256
    assert list(by_id) == [str(i) for i in np.nonzero(keep)[0]]
    got = np.array(list(by_id.values()))
    assert np.max(np.abs(got - want)) <= TOL and np.array_equal(np.sign(got), np.sign(want))
    assert np.array_equal(got, np.array([float(str(v)) for v in scorer.scores]))       # written with astype(str): round-trips
    # the counts were cached next to the FASTA as a feature CSV (scripts/phamer.py:131-134) and a second run reads them
    cached_ids, cached = fileIO.read_feature_file(str(indir / "contigs_features.csv"))
    assert np.array_equal(cached, counts) and [str(i) for i in cached_ids] == [str(i) for i in range(300)]
    again = phamer.main(["-in", str(indir), "-out", str(tmp_path / "out2"), "-equal", "-l", "0", "-m", "knn"])
    assert again.scores.shape == (300,) and set(np.unique(again.scores)) <= {-1.0, 1.0}
    assert np.array_equal(again.scores[keep], np.sign(want))
