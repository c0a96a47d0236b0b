"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; the package refuses to compute without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    from phamers_b200 import _lib
    return _lib


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "phamers_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(phm_[a-z0-9_]+)\s*\(", header))
    assert {"phm_kmer_count", "phm_score", "phm_pack_fasta", "phm_normalize_counts"} <= declared
    handle = lib.load()
    for name in sorted(declared):
        assert hasattr(handle, name), "library does not export %s" % name
        assert name in lib.SIGNATURES, "ctypes binding lacks %s" % name
    assert set(lib.SIGNATURES) <= declared | {"phm_set_option"} or set(lib.SIGNATURES) <= declared


def test_version_and_bins(lib):
    handle = lib.load()
    assert handle.phm_version() == 100
    assert [handle.phm_num_bins(k, 0) for k in range(1, 7)] == [4, 16, 64, 256, 1024, 4096]
    assert [handle.phm_num_bins(k, 1) for k in (4, 5, 6)] == [136, 512, 2080]
    assert handle.phm_num_bins(7, 0) == -1


def test_argument_errors_are_reported_without_a_gpu(lib):
    handle = lib.load()
    rc = handle.phm_kmer_count(None, None, 5, 9, 0, None, None, None, 0, None)
    assert rc == -1 and b"k must be 1..6" in handle.phm_last_error()
    rc = handle.phm_set_option(b"no_such_option", 1)
    assert rc == -1 and b"unknown option" in handle.phm_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    from phamers_b200 import kmer, phamer
    with pytest.raises(lib.PhamersLibraryError):
        kmer.count_string("ATGCATGC", 4)
    with pytest.raises(lib.PhamersLibraryError):
        kmer.normalize_counts(np.ones((2, 256), dtype=int))
    with pytest.raises(lib.PhamersLibraryError):
        phamer.score_points(np.ones((2, 256)) / 256, np.ones((5, 256)) / 256, np.ones((5, 256)) / 256, method="knn")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phamers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "kmer_oracle" not in text, f
