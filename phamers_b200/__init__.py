"""
phamers_b200 -- B200-native implementation of the PhaMers hot path
(k-mer count -> normalise -> score against the phage / bacteria reference features).

Drop-in modules mirroring the reference's call signatures:
    phamers_b200.kmer      <- scripts/kmer.py      (count_string, count, count_file, normalize_counts, ...)
    phamers_b200.phamer    <- scripts/phamer.py    (phamer_scorer, score_points)
    phamers_b200.learning  <- scripts/learning.py  (knn, distances, closest_to, get_centroids, kmeans)
    phamers_b200.fileIO    <- scripts/fileIO.py    (feature / score CSV formats, FASTA ids)

All sequence and feature arithmetic runs in hand-written CUDA (sm_100a) behind the C-ABI declared in
include/phamers_b200.h (phamers_b200/lib/libphamers_b200.so, built by __graft_entry__.build()).  There is no CPU
fallback: importing works anywhere, but every compute entry point raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
