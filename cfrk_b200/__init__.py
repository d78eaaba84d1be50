"""cfrk_b200 -- B200-native (sm_100a) implementation of CFRK's per-read k-mer counting path.

The product is libcfrk_b200.so (hand-written CUDA behind the C ABI of include/cfrk_b200.h)
and the `cfrk` command line.  This package is the thin host-side mirror used by the tests
and bench.py: ctypes bindings, nothing else.  There is no CPU fallback: importing works
anywhere, calling a compute entry point without a CUDA device raises CfrkError.
"""
from .api import (  # noqa: F401
    CfrkError, FMT_ASCII, FMT_CODES, MODE_COMPAT, MODE_EXACT, RUN_ALL_ROWS, RUN_EXACT, RUN_SPARSE,
    count_dense_device, count_dense_host, count_dense_packed_device, count_sparse_device, count_sparse_packed_device, dense_reads_per_tile, device_count, encode_2bit_device,
    global_hist_device, kmer_main, launch_count, lib, run_file, scan_fasta_device, version,
)
