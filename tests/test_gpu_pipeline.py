"""The host-buffer entry of the hot path (pipeline.ContigScorer.score_host: what bench.py times end to end): chunked upload on a
copy stream overlapped with counting and scoring must give exactly the device-resident path's scores."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_score_host_equals_score_device():
    from phamers_b200 import ops, pipeline
    scorer = pipeline.ContigScorer()
    seq, off = ops.synth_contigs(20260101, 12345, 6000)                   # ~96 Mbases: several upload ranges
    _, want = scorer.score_device(seq, off)
    want = want.cpu().numpy()
    host_seq = torch.empty((seq.numel(),), dtype=torch.uint8, pin_memory=True)
    host_seq.copy_(seq)
    host_off = torch.empty((off.numel(),), dtype=torch.int64, pin_memory=True)
    host_off.copy_(off)
    torch.cuda.synchronize()
    got = scorer.score_host(host_seq, host_off)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert np.array_equal(scorer.score_host(host_seq, host_off), want)    # staging buffers reused
    # pageable numpy buffers, other methods, a single contig, nothing at all
    np_seq, np_off = host_seq.numpy().copy(), host_off.numpy().copy()
    for method in ("knn", "kmeans"):
        _, dev = scorer.score_device(seq, off, method=method)
        assert np.array_equal(scorer.score_host(np_seq, np_off, method=method), dev.cpu().numpy())
    one = scorer.score_host(np_seq[:int(np_off[1])], np_off[:2])
    assert one.shape == (1,) and one[0] == want[0]
    assert scorer.score_host(np_seq[:0], np_off[:1]).shape == (0,)
    # another stream: its own workspaces, same scores
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        again = scorer.score_host(host_seq, host_off)
    assert np.array_equal(again, want)
