"""
The whole hot path as one object: contigs in, PhaMers scores out.

    ContigScorer(positive, negative)          reference features (float64, normalised) -> device, centroids cached
      .score_device(seq, offsets)             device buffers -> (counts int32[n,256], scores float64[n]) on the device
      .score_host(seq, offsets)               HOST buffers (pinned recommended) -> numpy scores; copies included
      .score_fasta(path)                      FASTA file -> (ids, scores), the phamer.py flow without the files

Stage by stage this is kmer.count_file -> kmer.normalize_counts -> phamer_scorer.score_points('combo')
(reference scripts/phamer.py:131,139,194), with the k-mer length fixed at the width of the reference features.
"""
import numpy as np
import torch

from . import _lib, ops, references


class ContigScorer(object):
    def __init__(self, positive=None, negative=None, k_clusters=86, k_neighbors=3, kmer_length=4, centroids=None,
                 equalize=True):
        _lib.require_cuda()
        if positive is None or negative is None:
            positive, negative = references.load_reference_features(equalize=equalize)
        positive = np.ascontiguousarray(positive, dtype=np.float64)
        negative = np.ascontiguousarray(negative, dtype=np.float64)
        if positive.shape[1] != 4 ** kmer_length or negative.shape[1] != 4 ** kmer_length:
            raise ValueError("reference features are %d wide, k = %d needs %d" % (positive.shape[1], kmer_length, 4 ** kmer_length))
        if centroids is None:
            centroids = references.reference_centroids(positive, negative, k_clusters)
        self.kmer_length = kmer_length
        self.k_neighbors = k_neighbors
        self.n_positive = positive.shape[0]
        self.refs = torch.from_numpy(np.vstack((positive, negative))).cuda()
        self.cent_pos = torch.from_numpy(np.ascontiguousarray(centroids[0], dtype=np.float64)).cuda()
        self.cent_neg = torch.from_numpy(np.ascontiguousarray(centroids[1], dtype=np.float64)).cuda()
        self._staging = None
        self._host_out = None

    # -- device resident --------------------------------------------------------------------------------
    def score_device(self, seq, offsets, method="combo", return_counts=True):
        # one library call: the histogram kernel also emits the scorer's query operands, the scorer forms count / row total
        # wherever it needs an exact feature; the float64 feature matrix is never written
        counts, knn, kmeans, combo = ops.count_score_cuda(seq, offsets, self.refs, self.n_positive, self.cent_pos, self.cent_neg,
                                                          self.k_neighbors)
        return counts, {"knn": knn, "kmeans": kmeans, "combo": combo}[method]

    # -- host buffers -----------------------------------------------------------------------------------
    def score_host(self, seq, offsets, method="combo"):
        """seq: uint8 host tensor / ndarray with all contigs end to end, offsets: int64[n+1].  Host->device copy of
        the bases and offsets, the three stages, and the device->host copy of the scores all happen here."""
        seq_t = seq if isinstance(seq, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(seq))
        off_t = offsets if isinstance(offsets, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64))
        total = seq_t.numel()
        padded = (total + 15) // 16 * 16 + 16
        if self._staging is None or self._staging.numel() < padded:
            self._staging = torch.empty((padded,), dtype=torch.uint8, device="cuda")
        d_seq = self._staging[:total]
        d_seq.copy_(seq_t, non_blocking=True)
        d_off = off_t.to("cuda", non_blocking=True)
        _, scores = self.score_device(d_seq, d_off, method=method, return_counts=False)
        if self._host_out is None or self._host_out.numel() < scores.numel():
            self._host_out = torch.empty((scores.numel(),), dtype=torch.float64, pin_memory=True)   # pinned once, reused
        out = self._host_out[:scores.numel()]
        out.copy_(scores, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out.numpy().copy()

    def score_fasta(self, path, length_requirement=0, method="combo"):
        from . import fileIO
        headers, d_seq, d_off = fileIO.read_fasta_arrays_cuda(path)      # records are found on the device; the bases stay there
        ids = np.array([fileIO.get_id(h) for h in headers])
        if len(ids) == 0:
            return ids, np.zeros((0,), dtype=np.float64)
        _, d_scores = self.score_device(d_seq, d_off, method=method, return_counts=False)
        scores = d_scores.cpu().numpy()
        if length_requirement:
            keep = np.diff(d_off.cpu().numpy()) >= length_requirement    # scripts/phamer.py:154
            ids, scores = ids[keep], scores[keep]
        return ids, scores
