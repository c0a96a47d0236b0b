"""
The on-disk formats either side of the hot path, mirroring the reference's scripts/fileIO.py and the part of
scripts/id_parser.py the path touches:

    FASTA in        read_fasta :28, get_fasta_ids :62 (Bio.SeqIO tokenisation) + id_parser.get_id :89
    feature CSV     read_feature_file :134, save_counts :169  ('#' header lines, then id,c0,c1,...)
    score CSV       save_phamer_scores :241, read_phamer_output :256  ('# ' header, then 'id, score')

Host-side only (bytes and text); no arithmetic happens here.
"""
import gzip
import io
import os

import numpy as np


# ----------------------------------------------------------------------------------------------------------
# ids (scripts/id_parser.py)
# ----------------------------------------------------------------------------------------------------------
def _represents_float(text):
    try:
        float(text)
        return True
    except ValueError:
        return False


def _is_genbank_id(token):
    """scripts/id_parser.py:79-86."""
    return len(token) >= 2 and not _represents_float(token) and token[-2] == "."


def get_id(header):
    """scripts/id_parser.py:89-100: '_ID_' contig headers -> the token after 'ID' (minus '-circular'); headers with
    four '|' -> field 3 (phage GenBank id); anything else -> the leading GenBank-style token."""
    if "_ID_" in header:
        parts = header.strip().replace(">", "").split("_")
        return parts[1 + parts.index("ID")].replace("-circular", "")
    if header.count("|") == 4:
        return header.split("|")[3].replace(">", "")
    token = header.split(" ")[0]
    if _is_genbank_id(token):
        return token
    fields = header.split("\t")
    if len(fields) > 1 and _is_genbank_id(fields[1].replace(">", "")):
        return fields[1].replace(">", "")
    return None


# ----------------------------------------------------------------------------------------------------------
# FASTA
# ----------------------------------------------------------------------------------------------------------
_TRAILING = b" \t\n\r\x0b\x0c"


def _open_bytes(path):
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as fh:
            return fh.read()
    with open(path, "rb") as fh:
        return fh.read()


def split_fasta_bytes(raw):
    """FASTA bytes -> (record ids [first blank-separated token of each title], uint8 array of all sequence bytes end
    to end, int64 offsets[n+1]).  Tokenisation follows Bio.SeqIO's FASTA parser as the reference uses it
    (scripts/kmer.py:135): a record starts at a line beginning with '>', its sequence is every following line
    right-stripped and joined, with blanks and carriage returns removed -- so k-mers span line breaks but never
    records.  Text before the first '>' is ignored."""
    data = np.frombuffer(raw, dtype=np.uint8)
    n = data.shape[0]
    if n == 0:
        return [], np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    newline = np.flatnonzero(data == 10)
    line_start = np.concatenate(([0], newline + 1))
    line_start = line_start[line_start < n]
    header_start = line_start[data[line_start] == ord(">")]
    if header_start.shape[0] == 0:
        return [], np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)
    # end of each header line (position of its '\n', or n)
    idx = np.searchsorted(newline, header_start)
    if newline.shape[0]:
        header_end = np.where(idx < newline.shape[0], newline[np.minimum(idx, newline.shape[0] - 1)], n)
    else:
        header_end = np.full_like(header_start, n)                    # a lone header line without a line feed
    body_start = np.minimum(header_end + 1, n)
    body_end = np.concatenate((header_start[1:], [n]))

    ids = []
    for hs, he in zip(header_start, header_end):
        title = raw[hs + 1:he].decode("latin-1").rstrip()
        tokens = title.split(None, 1)
        ids.append(tokens[0] if tokens else "")

    if any(c in raw for c in (b"\t", b"\x0b", b"\x0c")):
        # rare: tab / VT / FF are stripped only at line ends -- exact line-by-line path
        pieces, lengths = [], []
        for bs, be in zip(body_start, body_end):
            body = b"".join(line.rstrip(_TRAILING) for line in raw[bs:be].split(b"\n"))
            body = body.replace(b" ", b"").replace(b"\r", b"")
            pieces.append(body)
            lengths.append(len(body))
        seq = np.frombuffer(b"".join(pieces), dtype=np.uint8)
        offsets = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
        return ids, seq, offsets

    keep = (data != 10) & (data != 13) & (data != 32)
    in_body = np.zeros(n + 1, dtype=np.int32)
    np.add.at(in_body, body_start, 1)
    np.add.at(in_body, body_end, -1)
    keep &= np.cumsum(in_body[:n]) > 0
    kept_before = np.concatenate(([0], np.cumsum(keep)))
    offsets = np.concatenate((kept_before[body_start], [kept_before[n]])).astype(np.int64)
    # bodies are disjoint and ordered, so offsets[i+1] == kept_before[body_end[i]]
    return ids, data[keep], offsets


UPLOAD_CHUNK = 256 << 20             # bytes per pinned staging buffer of the file upload (two buffers alternate)


def _upload(buf, n):
    """File bytes (an mmap, or the bytes of a decompressed .gz) -> CUDA uint8 tensor padded to 16 bytes.  The copy page cache ->
    pinned staging buffer of chunk i + 1 runs on the host while chunk i crosses PCIe (scripts/kmer.py:131-134 reads the whole
    file through Python objects instead)."""
    import torch
    d_raw = torch.empty(((n + 15) // 16 * 16 + 16,), dtype=torch.uint8, device="cuda")
    chunk = min(UPLOAD_CHUNK, max(n, 1))
    stage = [torch.empty((chunk,), dtype=torch.uint8, pin_memory=True) for _ in range(2 if n > chunk else 1)]
    done = [None] * len(stage)
    src = np.frombuffer(buf, dtype=np.uint8, count=n)
    for i, lo in enumerate(range(0, n, chunk)):
        j = i % len(stage)
        if done[j] is not None:
            done[j].synchronize()                                     # the staging buffer's previous upload has left it
        m = min(chunk, n - lo)
        stage[j].numpy()[:m] = src[lo:lo + m]
        d_raw[lo:lo + m].copy_(stage[j][:m], non_blocking=True)
        done[j] = torch.cuda.Event()
        done[j].record()
    torch.cuda.current_stream().synchronize()                         # the staging buffers are released on return
    return d_raw


def read_fasta_arrays_cuda(fasta_file):
    """Device-side tokenisation of a FASTA file (phm_fasta_index / phm_fasta_extract): the file is memory-mapped, its bytes go
    to the GPU once through pinned staging buffers and the sequence bytes never come back; only the title lines are read on the
    host, at the positions the device reports.  Returns (titles' first tokens, CUDA uint8 tensor of all sequence bytes end to end
    [padded to 16], CUDA int64 offsets[n+1]).  Files holding a tab / VT / FF (stripped by the reference at line ends only) are
    tokenised by the exact host path and uploaded.  Raises IOError when unreadable."""
    import mmap
    import torch
    from . import ops
    mapped = None
    if fasta_file.endswith(".gz"):
        raw = _open_bytes(fasta_file)
        n = len(raw)
    else:
        fh = open(fasta_file, "rb")
        n = os.fstat(fh.fileno()).st_size
        raw = mapped = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) if n else b""
        fh.close()
    try:
        if n == 0:
            return [], torch.zeros((16,), dtype=torch.uint8, device="cuda"), torch.zeros((1,), dtype=torch.int64, device="cuda")
        d_raw = _upload(raw, n)
        seq, offsets, header_pos, odd = ops.fasta_scan_cuda(d_raw[:n])
        if odd:
            ids, h_seq, h_off = split_fasta_bytes(bytes(raw[:n]))
            total = int(h_off[-1])
            pad = torch.zeros((max((total + 15) // 16 * 16, 16),), dtype=torch.uint8)
            pad[:total] = torch.from_numpy(np.ascontiguousarray(h_seq[:total]).copy()) if total else pad[:0]
            return ids, pad.cuda(), torch.from_numpy(np.ascontiguousarray(h_off)).cuda()
        ids = []
        for hs in header_pos.cpu().tolist():
            he = raw.find(b"\n", hs)
            title = bytes(raw[hs + 1:(he if he >= 0 else n)]).decode("latin-1").rstrip()
            tokens = title.split(None, 1)
            ids.append(tokens[0] if tokens else "")
        return ids, seq, offsets
    finally:
        if mapped is not None:
            mapped.close()


def read_fasta_arrays(fasta_file):
    """(titles' first tokens, sequence bytes end to end, offsets).  Raises IOError when unreadable."""
    return split_fasta_bytes(_open_bytes(fasta_file))


def read_fasta(fasta_file):
    """scripts/fileIO.py:28-42: (ndarray of ids, list of sequence strings)."""
    headers, seq, offsets = read_fasta_arrays(fasta_file)
    blob = seq.tobytes().decode("latin-1")
    sequences = [blob[offsets[i]:offsets[i + 1]] for i in range(len(headers))]
    return np.array([get_id(h) for h in headers]), sequences


def get_fasta_ids(fasta_file):
    """scripts/fileIO.py:62-77."""
    headers, _, _ = read_fasta_arrays(fasta_file)
    return np.array([get_id(h) for h in headers])


def get_fasta_lengths(fasta_file):
    """Lengths of the parsed sequences (what phamer.screen_by_length measures, scripts/phamer.py:150-154)."""
    headers, _, offsets = read_fasta_arrays(fasta_file)
    return np.array([get_id(h) for h in headers]), np.diff(offsets)


# ----------------------------------------------------------------------------------------------------------
# feature CSV
# ----------------------------------------------------------------------------------------------------------
def generate_summary(args, line_start="", header=""):
    """scripts/basic.py:22-37: argparse namespace -> header text."""
    if args is None:
        return ""
    text = str(args).replace("Namespace(", line_start).replace(")", "")
    text = text.replace(", ", "\n" + line_start).replace("=", ":\t") + "\n"
    return line_start + header + "\n" + text


def read_feature_file(feature_file, normalize=False, id=None):
    """scripts/fileIO.py:134-166: (ids ndarray[str], counts int ndarray[n, bins]); normalize -> kmer.normalize_counts."""
    ids, rows = [], []
    opener = gzip.open if feature_file.endswith(".gz") else open
    with opener(feature_file, "rt") as fh:
        for line in fh:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            first, rest = line.split(",", 1)
            ids.append(first)
            rows.append(np.array(rest.split(","), dtype=np.int64))
    features = np.stack(rows) if rows else np.zeros((0, 0), dtype=np.int64)
    ids = np.array(ids)
    if normalize:
        from . import kmer
        features = kmer.normalize_counts(features)
    if id:
        return features[ids == id]
    return ids, features


def save_counts(counts, ids, file_name, args=None, header="K-mer count file"):
    """scripts/fileIO.py:169-181: '# '-prefixed header lines, then id,c0,c1,... as integers."""
    if args is not None:
        header = generate_summary(args, header=header)
    counts = np.asarray(counts).astype(np.int64)
    if counts.ndim == 1:
        counts = counts[None, :]
    with open(file_name, "w") as fh:
        for line in header.split("\n"):
            fh.write("# " + line + "\n")
        for ident, row in zip(ids, counts):
            fh.write(str(ident) + "," + ",".join(map(str, row.tolist())) + "\n")


# ----------------------------------------------------------------------------------------------------------
# score CSV
# ----------------------------------------------------------------------------------------------------------
def save_phamer_scores(ids, scores, file_name, args=None):
    """scripts/fileIO.py:241-253: header 'PhaMers score file', rows 'id, score' with scores printed by
    numpy's float64 -> str conversion."""
    header = "PhaMers score file"
    if args is not None:
        header = generate_summary(args, header=header)
    ids = np.asarray(ids).astype(str)
    scores = np.asarray(scores, dtype=np.float64).astype(str)
    with open(file_name, "w") as fh:
        for line in header.split("\n"):
            fh.write("# " + line + "\n")
        for ident, score in zip(ids, scores):
            fh.write("%s, %s\n" % (ident, score))


def read_phamer_output(filename):
    """scripts/fileIO.py:256-272: {contig id: score}."""
    out = {}
    with open(filename, "r") as fh:
        for line in fh:
            if "#" in line or not line.strip():
                continue
            out[line.split(",")[0]] = float(line.split()[1])
    return out
