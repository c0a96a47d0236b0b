"""Compiles phamers_b200/csrc/kmer_swar.h for the host and checks the SWAR decode (2-bit codes, validity test, blank
masks, window extraction, reverse-complement bins) exhaustively / on random bytes.  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swar_header_on_host(tmp_path):
    exe = str(tmp_path / "swar_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "swar_host_check.cpp")])
    out = subprocess.check_output([exe]).decode()
    assert "swar ok" in out
