"""GPU parity tests for stage 1+2 (k-mer counting, normalising) through the C-ABI: bit-exact against the golden
vectors from the unmodified reference, against the oracle on seeded random inputs, and through size-independent
properties at larger sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import phamers_oracle as po

pytestmark = pytest.mark.gpu


def _seqs(g):
    blob, off = g["seq_bytes"].tobytes().decode("ascii"), g["seq_offsets"]
    return [blob[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def _device(seq_bytes, offsets):
    total = int(offsets[-1])
    buf = torch.zeros(((total + 15) // 16 * 16 + 16,), dtype=torch.uint8)
    buf[:total] = torch.from_numpy(np.array(seq_bytes[:total], dtype=np.uint8))
    return buf.cuda(), torch.from_numpy(np.asarray(offsets, dtype=np.int64)).cuda()


def _u32(t):
    return t.cpu().numpy().view(np.uint32).astype(np.int64)


@pytest.fixture(scope="module")
def counting(golden_dir):
    return np.load(os.path.join(golden_dir, "counting_golden.npz"))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_counts_match_reference_golden(counting, k):
    from phamers_b200 import kmer, ops, _lib
    seqs = _seqs(counting)
    want = counting["counts_k%d" % k]
    got = kmer.count(seqs, k)
    assert got.dtype == np.int64 and got.shape == want.shape
    assert np.array_equal(got, want)
    # every kernel variant: simple cross-check kernel, stride-1 histogram for k = 4, packed input
    d_seq, d_off = _device(counting["seq_bytes"], counting["seq_offsets"])
    naive, _ = ops.count_cuda(d_seq, d_off, k, naive=True)
    assert np.array_equal(_u32(naive), want)
    if k == 4:
        _lib.set_option("hist_stride_k4", 1)
        try:
            s1, _ = ops.count_cuda(d_seq, d_off, k)
        finally:
            _lib.set_option("hist_stride_k4", 2)
        assert np.array_equal(_u32(s1), want)
    codes, valid = ops.pack_cuda(d_seq)
    packed, _ = ops.count_packed_cuda(codes, valid, d_off, k)
    assert np.array_equal(_u32(packed), want)


def test_dispatch_conventions_and_normalize(counting):
    from phamers_b200 import kmer
    seqs = _seqs(counting)
    one = kmer.count([seqs[12]], 4)
    assert one.shape == (256,) and np.array_equal(one, counting["dispatch_list1"])
    three = kmer.count(seqs[8:11], 4)
    assert three.shape == (3, 256) and np.array_equal(three, counting["dispatch_list3"])
    norm = kmer.count(seqs[13], 4, normalize=True)
    assert norm.dtype == np.float64 and np.array_equal(norm, counting["dispatch_str_norm"])
    assert kmer.count(12345, 4) is None
    assert kmer.count_string("AAAT", 4)[1] == 1                       # scripts/kmer.py:90-91
    got = kmer.normalize_counts(counting["counts_k4"])
    assert got.dtype == np.float64
    assert np.array_equal(got, counting["normalized_k4"], equal_nan=True)   # bit-identical IEEE division, NaN rows kept
    assert np.array_equal(kmer.count(seqs[:3], 4, normalize=True), np.zeros((3, 256)))   # kmer.py:77 guard
    rna = kmer.count_string("AUGCAUGGN", 2, symbols=kmer.RNA)
    assert np.array_equal(rna, po.count_string("AUGCAUGGN", 2, symbols="AUGC"))
    with pytest.raises(NotImplementedError):
        kmer.count_string("ARND", 2, symbols=kmer.protein)


def test_count_file_matches_reference_golden(golden_dir, tmp_path):
    from phamers_b200 import kmer
    g = np.load(os.path.join(golden_dir, "fasta_golden.npz"))
    path = tmp_path / "contigs.fasta"
    with open(path, "w", newline="") as fh:
        fh.write(str(g["fasta_text"]))
    for k in (4, 5, 6):
        ids, counts = kmer.count_file(str(path), k)
        assert [str(x) for x in ids] == [str(x) for x in g["ids"]]
        assert counts.dtype == np.int64 and np.array_equal(counts, g["counts_k%d" % k])
    _, freq = kmer.count_file(str(path), 4, normalize=True)
    want = g["freq_k4"]
    assert np.array_equal(freq, want, equal_nan=True)
    assert kmer.count_file(str(tmp_path / "missing.fasta"), 4) == (None, None)


def _random_workload(rng, n, dirty):
    lengths = np.clip(np.round(np.exp(rng.normal(np.log(3000), 1.2, size=n))), 0, 60000).astype(np.int64)
    lengths[rng.integers(0, n, size=max(1, n // 50))] = rng.integers(0, 8, size=max(1, n // 50))   # shorter than k
    offsets = np.concatenate(([0], np.cumsum(lengths)))
    total = int(offsets[-1])
    seq = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=total)
    if dirty:
        for _ in range(total // 400):
            p = int(rng.integers(0, total))
            run = int(rng.integers(1, 40))
            seq[p:p + run] = rng.choice(np.frombuffer(b"NnatgcRYKMSWU-*. ", dtype=np.uint8), size=len(seq[p:p + run]))
    return seq, offsets


@pytest.mark.parametrize("k", [4, 5, 6])
@pytest.mark.parametrize("dirty", [False, True])
def test_random_contigs_match_oracle(k, dirty):
    from phamers_b200 import ops
    rng = np.random.default_rng(100 * k + dirty)
    seq, offsets = _random_workload(rng, 3000, dirty)
    want = c_oracle.count(seq, offsets, k)
    d_seq, d_off = _device(seq, offsets)
    counts, freq = ops.count_cuda(d_seq, d_off, k, freq=True)
    assert np.array_equal(_u32(counts), want)
    want_freq = c_oracle.normalize(want)
    assert np.array_equal(freq.cpu().numpy(), want_freq, equal_nan=True)
    # canonical fold = the fixed linear fold of the oracle's 4^k vector
    canon, _ = ops.count_cuda(d_seq, d_off, k, canonical=True)
    assert np.array_equal(_u32(canon), po.canonical_fold(want, k))
    # packed path
    codes, valid = ops.pack_cuda(d_seq)
    packed, _ = ops.count_packed_cuda(codes, valid, d_off, k)
    assert np.array_equal(_u32(packed), want)


def test_unaligned_offsets_and_single_long_contig():
    from phamers_b200 import ops
    rng = np.random.default_rng(5)
    # contigs starting at every alignment, including empty ones between them
    lengths = np.array([17, 0, 1, 3, 4, 5, 16, 15, 31, 33, 0, 64, 1000, 7, 511, 513, 2, 700001], dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(lengths)))
    seq = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(offsets[-1]))
    seq[offsets[-2] + 350000] = ord("N")
    d_seq, d_off = _device(seq, offsets)
    for k in (4, 5, 6):
        counts, _ = ops.count_cuda(d_seq, d_off, k)
        assert np.array_equal(_u32(counts), c_oracle.count(seq, offsets, k))


def test_pack_format():
    from phamers_b200 import ops
    text = b"ATGCNatgcATGCATGCATGCATGCATGCATGCA-T"
    d = torch.zeros((48,), dtype=torch.uint8)
    d[:len(text)] = torch.tensor(list(text), dtype=torch.uint8)
    codes, valid = ops.pack_cuda(d.cuda())
    codes = codes.cpu().numpy().view(np.uint32)
    valid = valid.cpu().numpy().view(np.uint32)
    sym = {ord("A"): 0, ord("T"): 1, ord("G"): 2, ord("C"): 3}
    padded = np.zeros(48, dtype=np.uint8)
    padded[:len(text)] = np.frombuffer(text, dtype=np.uint8)
    for i, c in enumerate(padded):
        code = (int(codes[i // 16]) >> (30 - 2 * (i % 16))) & 3
        ok = (int(valid[i // 32]) >> (31 - (i % 32))) & 1
        assert ok == (1 if c in sym else 0), i
        assert code == sym.get(int(c), 0), i


def test_properties_at_scale():
    """Size-independent checks on a device-generated workload: row sums equal L - k + 1 for clean contigs, the
    stride-2 and stride-1 histograms and the simple kernel agree, canonical mass is conserved."""
    from phamers_b200 import ops, _lib
    seq, off = ops.synth_contigs(20260101, 0, 20000)
    lengths = (off[1:] - off[:-1]).cpu().numpy()
    assert lengths.min() >= 1000 and lengths.max() <= 100000
    counts4, freq4 = ops.count_cuda(seq, off, 4, freq=True)
    c4 = _u32(counts4)
    assert np.array_equal(c4.sum(axis=1), lengths - 3)
    assert np.allclose(freq4.sum(dim=1).cpu().numpy(), 1.0, atol=1e-12)
    naive, _ = ops.count_cuda(seq, off, 4, naive=True)
    assert torch.equal(naive, counts4)
    _lib.set_option("hist_stride_k4", 1)
    try:
        s1, _ = ops.count_cuda(seq, off, 4)
    finally:
        _lib.set_option("hist_stride_k4", 2)
    assert torch.equal(s1, counts4)
    for k in (5, 6):
        ck, _ = ops.count_cuda(seq, off, k)
        assert np.array_equal(_u32(ck).sum(axis=1), lengths - k + 1)
        # marginalising the last base of the k-mers gives the (k-1)-mers except the final window
        canon, _ = ops.count_cuda(seq, off, k, canonical=True)
        assert np.array_equal(_u32(canon).sum(axis=1), lengths - k + 1)
    # a 10 000-contig prefix against the C oracle, byte for byte
    n = 10000
    end = int(off[n].item())
    host_seq = seq[:end].cpu().numpy()
    host_off = off[:n + 1].cpu().numpy()
    assert np.array_equal(c4[:n], c_oracle.count(host_seq, host_off, 4))
    # BASELINE configs[1] / [2] at full size (1 M contigs, 16 Gbases): row sums of clean contigs, every k, plain and canonical
    del counts4, freq4, naive, s1
    seq, off = ops.synth_contigs(20260101, 0, 1000000)
    lengths = (off[1:] - off[:-1])
    for k in (4, 5, 6):
        for canonical in (False, True):
            ck, _ = ops.count_cuda(seq, off, k, canonical=canonical)
            assert torch.equal(ck.sum(dim=1, dtype=torch.int64), lengths - (k - 1)), (k, canonical)
            del ck
    # determinism of the generator across shards
    seq, off = ops.synth_contigs(20260101, 0, 20000)
    seq2, off2 = ops.synth_contigs(20260101, 100, 50)
    a0, a1 = int(off[100].item()), int(off[150].item())
    assert torch.equal(seq2, seq[a0:a1])


@pytest.mark.parametrize("stride", [1, 2])
def test_every_k4_window_stride_matches_oracle(stride):
    """k = 4 histogram variants (plain 4-mers; 5-mers at every second base) on dirty, clean and awkward workloads: contigs at
    every alignment, a 700 kb contig, a 300 kb homopolymer (one bin takes every count) and a contig just past 192 kb."""
    from phamers_b200 import ops, _lib
    rng = np.random.default_rng(77)
    seq_a, off_a = _random_workload(rng, 1500, True)
    seq_b, off_b = _random_workload(rng, 1500, False)
    lengths = np.array([17, 0, 1, 3, 4, 5, 6, 16, 15, 31, 33, 0, 64, 1000, 7, 511, 513, 2, 700001, 300000, 9, 196608 + 5], dtype=np.int64)
    off_c = np.concatenate(([0], np.cumsum(lengths)))
    seq_c = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(off_c[-1]))
    seq_c[off_c[18] + 350000] = ord("N")
    seq_c[off_c[19]:off_c[20]] = ord("A")
    seq_c[off_c[19] + 123457] = ord("c")
    _lib.set_option("hist_stride_k4", stride)
    try:
        for seq, off in ((seq_a, off_a), (seq_b, off_b), (seq_c, off_c)):
            d_seq, d_off = _device(seq, off)
            want = c_oracle.count(seq, off, 4)
            counts, freq = ops.count_cuda(d_seq, d_off, 4, freq=True)
            assert np.array_equal(_u32(counts), want)
            assert np.array_equal(freq.cpu().numpy(), c_oracle.normalize(want), equal_nan=True)
            canon, _ = ops.count_cuda(d_seq, d_off, 4, canonical=True)
            assert np.array_equal(_u32(canon), po.canonical_fold(want, 4))
    finally:
        _lib.set_option("hist_stride_k4", 2)


@pytest.mark.parametrize("k,option,value,default", [(5, "hist_stride_k5", 1, 0), (5, "hist_stride_k5", 2, 0),
                                                     (5, "hist_warps_k5", 18, 18), (5, "hist_warps_k5", 8, 18),
                                                     (6, "hist_warps_k6", 4, 13), (6, "hist_warps_k6", 13, 13)])
def test_k5_k6_kernel_variants_match_oracle(k, option, value, default):
    """k = 5 as plain 5-mers or as 6-mers at every second base in 16-bit packed counters (folded every 248 steps: the 300 kb contig
    and the 100 kb homopolymer would overflow them otherwise); k = 6 with 4 or 13 warps per CTA."""
    from phamers_b200 import ops, _lib
    rng = np.random.default_rng(1000 + 10 * k + value)
    seq_a, off_a = _random_workload(rng, 1200, True)
    lengths = np.array([17, 0, 1, 5, 6, 7, 16, 15, 31, 33, 0, 64, 1000, 7, 511, 513, 2, 300001, 100000, 9], dtype=np.int64)
    off_c = np.concatenate(([0], np.cumsum(lengths)))
    seq_c = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(off_c[-1]))
    seq_c[off_c[17] + 150000] = ord("N")
    seq_c[off_c[18]:off_c[19]] = ord("G")
    _lib.set_option(option, value)
    try:
        for seq, off in ((seq_a, off_a), (seq_c, off_c)):
            d_seq, d_off = _device(seq, off)
            want = c_oracle.count(seq, off, k)
            counts, freq = ops.count_cuda(d_seq, d_off, k, freq=True)
            assert np.array_equal(_u32(counts), want)
            assert np.array_equal(freq.cpu().numpy(), c_oracle.normalize(want), equal_nan=True)
            canon, cfreq = ops.count_cuda(d_seq, d_off, k, canonical=True, freq=True)
            want_c = po.canonical_fold(want, k)
            assert np.array_equal(_u32(canon), want_c)
            assert np.array_equal(cfreq.cpu().numpy(), c_oracle.normalize(want_c), equal_nan=True)
    finally:
        _lib.set_option(option, default)


@pytest.mark.parametrize("k", [4, 5, 6])
def test_tma_staged_histogram_matches_oracle(k):
    """Option hist_tma: the sequence reaches shared memory through cp.async.bulk copies completing on mbarriers (4 stages per
    warp) instead of 128-bit loads into registers.  Same counts, plain and canonical, on dirty, awkward and long inputs."""
    from phamers_b200 import ops, _lib
    rng = np.random.default_rng(4000 + k)
    seq_a, off_a = _random_workload(rng, 1500, True)
    lengths = np.array([17, 0, 1, 5, 6, 7, 16, 15, 31, 33, 0, 64, 1000, 7, 511, 512, 513, 2, 300001, 2048, 2049, 100000, 9], dtype=np.int64)
    off_c = np.concatenate(([0], np.cumsum(lengths)))
    seq_c = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(off_c[-1]))
    seq_c[off_c[18] + 150000] = ord("N")
    _lib.set_option("hist_tma", 1)
    try:
        for seq, off in ((seq_a, off_a), (seq_c, off_c)):
            d_seq, d_off = _device(seq, off)
            want = c_oracle.count(seq, off, k)
            counts, freq = ops.count_cuda(d_seq, d_off, k, freq=True)
            assert np.array_equal(_u32(counts), want)
            assert np.array_equal(freq.cpu().numpy(), c_oracle.normalize(want), equal_nan=True)
            canon, _ = ops.count_cuda(d_seq, d_off, k, canonical=True)
            assert np.array_equal(_u32(canon), po.canonical_fold(want, k))
    finally:
        _lib.set_option("hist_tma", 0)


def test_count_property_small_sequences():
    """Random short sequences over symbols and non-symbols, through the drop-in kmer.count for k = 1..6, against the oracle's
    restatement of kmer.count_string (reference scripts/kmer.py:42-50): empty strings, strings shorter than k, blanks at every
    position relative to the windows and to the 16-byte chunks."""
    from hypothesis import given, settings, strategies as st
    from phamers_b200 import kmer

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.text(alphabet="ATGCATGCATGCNatgc-", min_size=0, max_size=90), min_size=2, max_size=12), st.integers(1, 6))
    def check(seqs, k):
        got = kmer.count(seqs, k)
        want = np.stack([po.count_string_np(s, k) for s in seqs])
        assert got.shape == want.shape and np.array_equal(got, want)

    check()


def _long_workload(rng):
    """Records of genome size next to contig-size ones: a 16 Mbp record (cut into ~250 tiles), records just below / at / above the
    listing and splitting thresholds (16 384 / 131 072 bases), tile boundaries that fall on and next to blank bytes."""
    lengths = np.array([5000, 16 * 1024 * 1024 + 77, 16383, 16384, 131071, 131072, 131073, 3, 0, 262144 + 511, 65536 * 3 + 1, 40000,
                        1000000, 17, 196608], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(lengths)))
    seq = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(off[-1]), p=[0.35, 0.33, 0.15, 0.17])
    big = int(off[1])
    for t in range(1, 40):                                              # blanks on, just before and just after tile boundaries
        edge = ((big + t * 65536) // 512) * 512
        seq[edge + int(rng.integers(-6, 7))] = ord("N")
    seq[big + 3000000:big + 3000300] = ord("n")                        # a soft-masked stretch
    seq[int(off[12]):int(off[12]) + 400000] = ord("G")                  # a homopolymer run across tiles: one bin takes every count
    return seq, off


@pytest.mark.parametrize("k", [4, 5, 6, 2])
def test_long_records_are_tiled_bit_exact(k):
    """SURVEY.md section 5 (long records: kmer.count_directory runs over whole genomes, scripts/kmer.py:143-180): a record of
    131 072 bases or more is cut into 64 kb tiles counted by different warps and merged in its output row; counts, canonical
    counts and features must still be the oracle's, bit for bit, whatever outputs are asked for."""
    from phamers_b200 import ops, _lib
    rng = np.random.default_rng(160 + k)
    seq, off = _long_workload(rng)
    d_seq, d_off = _device(seq, off)
    want = c_oracle.count(seq, off, k)
    counts, freq = ops.count_cuda(d_seq, d_off, k, freq=True)
    assert np.array_equal(_u32(counts), want)
    assert np.array_equal(freq.cpu().numpy(), c_oracle.normalize(want), equal_nan=True)
    _, only_freq = ops.count_cuda(d_seq, d_off, k, counts=False, freq=True)            # tiles merge inside the feature rows
    assert torch.equal(only_freq, freq) or np.array_equal(only_freq.cpu().numpy(), freq.cpu().numpy(), equal_nan=True)
    want_c = po.canonical_fold(want, k)
    canon, cfreq = ops.count_cuda(d_seq, d_off, k, canonical=True, freq=True)
    assert np.array_equal(_u32(canon), want_c)
    assert np.array_equal(cfreq.cpu().numpy(), c_oracle.normalize(want_c), equal_nan=True)
    _, only_cfreq = ops.count_cuda(d_seq, d_off, k, canonical=True, counts=False, freq=True)
    assert np.array_equal(only_cfreq.cpu().numpy(), cfreq.cpu().numpy(), equal_nan=True)
    # the same without the work plan (every record one work item, as in round 1) and with a workspace too small for the list
    _lib.set_option("hist_plan", 0)
    try:
        plain, _ = ops.count_cuda(d_seq, d_off, k)
    finally:
        _lib.set_option("hist_plan", 1)
    assert torch.equal(plain, counts)
    lib = _lib.load()
    small = torch.empty((int(lib.phm_kmer_count_workspace_bytes(len(off) - 1, 0, k, 0)),), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(counts)
    _lib.check(lib.phm_kmer_count(_lib.ptr(d_seq), _lib.ptr(d_off), len(off) - 1, k, 0, _lib.ptr(out), None, _lib.ptr(small),
                                  small.numel(), _lib.stream_ptr()))
    assert torch.equal(out, counts)


def test_long_record_is_counted_in_parallel():
    """One 16 Mbp record on its own: ~250 tiles on as many warps.  Device time must be far below the ~30 ms a single warp needs."""
    from phamers_b200 import ops
    rng = np.random.default_rng(16)
    n = 16 * 1024 * 1024
    seq = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=n)
    off = np.array([0, n], dtype=np.int64)
    d_seq, d_off = _device(seq, off)
    counts, _ = ops.count_cuda(d_seq, d_off, 4)
    assert np.array_equal(_u32(counts), c_oracle.count(seq, off, 4))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.count_cuda(d_seq, d_off, 4, out_counts=counts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("16 Mbp record: %.3f ms per count" % ms)
    assert ms < 1.0


def test_count_score_with_long_records(tmp_path):
    """The fused path (histogram kernel emitting the scorer's operands) on a workload with tiled records: same counts and the
    same scores as counting first and scoring the counts."""
    from phamers_b200 import ops, pipeline
    rng = np.random.default_rng(99)
    lengths = np.array([20000, 300000, 1000, 140000, 5000, 9000, 131072, 70000], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(lengths)))
    seq = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=int(off[-1]), p=[0.3, 0.3, 0.2, 0.2])
    seq[int(off[1]) + 65536 * 2] = ord("N")
    d_seq, d_off = _device(seq, off)
    scorer = pipeline.ContigScorer()
    counts, combo = scorer.score_device(d_seq, d_off)
    want = c_oracle.count(seq, off, 4)
    assert np.array_equal(_u32(counts), want)
    plain, _ = ops.count_cuda(d_seq, d_off, 4)
    _, _, combo2 = ops.score_cuda(plain, scorer.refs, scorer.n_positive, scorer.cent_pos, scorer.cent_neg, 3)
    assert torch.equal(combo, combo2)


def test_count_directory_of_genome_files(tmp_path):
    """kmer.count_directory (reference scripts/kmer.py:143-180) over a directory of genome files -- how the reference builds its
    reference feature sets: a 10-record genome (records of 0.1 - 2.2 Mbp, 70-column lines), a single-chromosome file, a gzipped
    file, a file without usable sequence (dropped) and a file the identifier does not select."""
    import gzip
    from phamers_b200 import kmer
    rng = np.random.default_rng(2024)

    def record(name, length, gc):
        p = [(1 - gc) / 2, (1 - gc) / 2, gc / 2, gc / 2]
        s = rng.choice(np.frombuffer(b"ATGC", dtype=np.uint8), size=length, p=p)
        if length > 1000:
            s[length // 2:length // 2 + 40] = ord("N")
        text = s.tobytes().decode("ascii")
        return ">%s some description\n" % name + "\n".join(text[i:i + 70] for i in range(0, length, 70)) + "\n"

    genome = "".join(record("NC_%06d.1" % i, int(L), 0.3 + 0.04 * i) for i, L in
                     enumerate([2200000, 100000, 131072, 1500000, 400000, 65536, 180000, 1000000, 131071, 700000]))
    (tmp_path / "genome_a.fna").write_text(genome)
    (tmp_path / "genome_b.fna").write_text(record("NZ_CP000001.1", 3000001, 0.62))
    with gzip.open(str(tmp_path / "genome_c.fna.gz"), "wt") as fh:
        fh.write(record("NZ_CP000002.1", 250000, 0.5) + record("NZ_CP000003.1", 3, 0.5))
    (tmp_path / "empty.fna").write_text(">nothing\nNNNN\n")
    (tmp_path / "notes.txt").write_text(">ignored\nATGCATGC\n")
    for k in (4, 6):
        ids, counts = kmer.count_directory(str(tmp_path), k)
        want_ids, want = po.count_directory(str(tmp_path), k)
        assert list(ids) == list(want_ids) and len(ids) == 3
        assert counts.dtype == np.float64 and counts.shape == want.shape
        assert np.array_equal(counts, want)


def test_normalize_counts_takes_any_numeric_rows():
    """kmer.normalize_counts (reference scripts/kmer.py:209-221) is applied by the reference to whatever it is given: integer
    counts, the float64 sums of count_directory (which can pass 2^32), features that are already normalised."""
    from phamers_b200 import kmer
    rng = np.random.default_rng(5)
    ints = rng.integers(0, 5000, size=(40, 256))
    assert np.array_equal(kmer.normalize_counts(ints), po.normalize_counts(ints))
    sums = rng.integers(0, 2 ** 40, size=(7, 256)).astype(float)          # genome-directory sums: float, beyond 32 bits
    assert np.array_equal(kmer.normalize_counts(sums), po.normalize_counts(sums))
    feats = po.normalize_counts(ints) * 3.5                                # non-integer rows
    got, want = kmer.normalize_counts(feats), po.normalize_counts(feats)
    assert np.max(np.abs(got - want)) <= 1e-15 and got.dtype == np.float64
    one_row = kmer.normalize_counts(sums[0])
    assert one_row.shape == (256,) and np.array_equal(one_row, po.normalize_counts(sums[0]))
    zero = np.zeros((2, 16))
    assert np.isnan(kmer.normalize_counts(zero)).all()
    neg = np.array([[1.0, -3.0, 4.0, 2.0]])
    assert np.array_equal(kmer.normalize_counts(neg), po.normalize_counts(neg))
