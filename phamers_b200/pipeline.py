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
        self._copy_stream = None
        self._chunk_counts = None

    # -- device resident --------------------------------------------------------------------------------
    def score_device(self, seq, offsets, method="combo", return_counts=True):
        # one library call: the histogram kernel also emits the scorer's query operands, the scorer forms count / row total
        # wherever it needs an exact feature; the float64 feature matrix is never written
        counts, knn, kmeans, combo = ops.count_score_cuda(seq, offsets, self.refs, self.n_positive, self.cent_pos, self.cent_neg,
                                                          self.k_neighbors)
        return counts, {"knn": knn, "kmeans": kmeans, "combo": combo}[method]

    # -- host buffers -----------------------------------------------------------------------------------
    HOST_CHUNKS = 8                    # upload / compute pipeline depth of score_host
    HOST_CHUNK_MIN_BASES = 1 << 22     # ... and the least bases worth a range of its own

    def score_host(self, seq, offsets, method="combo"):
        """seq: uint8 host tensor / ndarray with all contigs end to end (pinned memory makes the copies asynchronous), offsets:
        int64[n+1].  The contigs are cut into HOST_CHUNKS ranges of about equal bases; range i + 1 is copied to the device on a
        copy stream while range i is counted and scored, and the scores come back in one copy at the end.  The host -> device
        copy (16 GB per million contigs against ~9 ms of compute) is what bounds this call: see DESIGN.md section 7."""
        seq_t = seq if isinstance(seq, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(seq))
        off_t = offsets if isinstance(offsets, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64))
        n = off_t.numel() - 1
        if n <= 0:
            return np.zeros((0,), dtype=np.float64)
        host_off = off_t.numpy()
        total = int(host_off[-1] - host_off[0])
        n_chunks = max(1, min(self.HOST_CHUNKS, n, total // self.HOST_CHUNK_MIN_BASES + 1))
        targets = host_off[0] + (total * np.arange(1, n_chunks, dtype=np.float64) / n_chunks)
        cuts = np.unique(np.concatenate(([0], np.clip(np.searchsorted(host_off, targets, side="left"), 0, n), [n])))
        padded = (total + 15) // 16 * 16 + 32 * len(cuts)
        if self._staging is None or self._staging.numel() < padded:
            self._staging = torch.empty((padded,), dtype=torch.uint8, device="cuda")
        if self._host_out is None or self._host_out.numel() < n:
            self._host_out = torch.empty((n,), dtype=torch.float64, pin_memory=True)             # pinned once, reused
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        d_off = off_t.to("cuda", non_blocking=True)
        scores = torch.empty((n,), dtype=torch.float64, device="cuda")
        widest = int(np.max(np.diff(cuts)))
        if self._chunk_counts is None or self._chunk_counts.shape[0] < widest:
            self._chunk_counts = torch.empty((widest, 256), dtype=torch.int32, device="cuda")
        self._copy_stream.wait_stream(main)                                  # the staging buffer may still be read by an earlier call
        uploads, place = [], 0
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            b0, b1 = int(host_off[lo]), int(host_off[hi])
            dst = self._staging[place:place + (b1 - b0)]
            with torch.cuda.stream(self._copy_stream):
                dst.copy_(seq_t[b0:b1], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
            uploads.append((int(lo), int(hi), b0, dst, done))
            place += (b1 - b0 + 15) // 16 * 16 + 16                          # every range starts 16-byte aligned, padding readable
        key = {"knn": 0, "kmeans": 1, "combo": 2}[method]
        for lo, hi, b0, dst, done in uploads:
            main.wait_event(done)
            part = [torch.empty((hi - lo,), dtype=torch.float64, device="cuda") if i != key else scores[lo:hi] for i in range(3)]
            ops.count_score_cuda(dst, d_off[lo:hi + 1] - b0, self.refs, self.n_positive, self.cent_pos, self.cent_neg, self.k_neighbors,
                                 out_counts=self._chunk_counts[:hi - lo], out=tuple(part))
        out = self._host_out[:n]
        out.copy_(scores, non_blocking=True)
        main.synchronize()
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
