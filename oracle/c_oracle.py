"""
TEST INFRASTRUCTURE ONLY -- ctypes wrapper over oracle/kmer_oracle.c (the plain-C restatement of
scripts/kmer.py:42-50,183-196,209-221).  Built by oracle/Makefile (also from __graft_entry__.build()).
Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkmer_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.phm_oracle_count.restype = ctypes.c_int
        _lib.phm_oracle_normalize.restype = ctypes.c_int
    return _lib


def count(seq_bytes, offsets, k):
    """int64 counts[n, 4^k] for sequences seq_bytes[offsets[i]:offsets[i+1]]."""
    seq = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    n = off.shape[0] - 1
    out = np.empty((n, 4 ** k), dtype=np.int64)
    rc = lib().phm_oracle_count(seq.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_int64(n), ctypes.c_int(k), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("phm_oracle_count failed: %d" % rc)
    return out


def normalize(counts):
    c = np.ascontiguousarray(counts, dtype=np.int64)
    out = np.empty(c.shape, dtype=np.float64)
    lib().phm_oracle_normalize(c.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(c.shape[0]),
                               ctypes.c_int64(c.shape[1]), out.ctypes.data_as(ctypes.c_void_p))
    return out
