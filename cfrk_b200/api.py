"""Host-side mirror of the reference's operator interface over the C ABI.

`kmer_main(data, start, length, k, device)` has the meaning of the reference's
`kmer_main(struct read*, nN, nS, k, device)` (src/kmer.cuh:6): reference-layout codes in,
dense int32 [nS, 4^k] rows out, compat semantics.  The *_device functions take raw device
pointers (ints) so that callers can keep data resident (bench.py uses torch tensors'
data_ptr(); torch is plumbing only and is never imported here).
"""
import ctypes as C
import os

import numpy as np

from . import _lib

FMT_CODES, FMT_ASCII = 0, 1
MODE_COMPAT, MODE_EXACT = 0, 1
RUN_ALL_ROWS, RUN_EXACT, RUN_SPARSE = 1, 2, 4


class CfrkError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = _lib.load().cfrk_last_error().decode(errors="replace")
        super().__init__(f"{where}: error {code}: {msg}")


def lib():
    return _lib.load()


def version():
    return lib().cfrk_version().decode()


def device_count():
    return lib().cfrk_device_count()


def launch_count():
    return int(lib().cfrk_launch_count())


def dense_reads_per_tile(k):
    return lib().cfrk_dense_reads_per_tile(k)


def _check(rc, where):
    if rc != 0:
        raise CfrkError(rc, where)


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def count_dense_host(bases, start, length, k, mode=MODE_COMPAT, fmt=FMT_CODES, device=0, out=None):
    """numpy in, numpy out through cfrk_count_dense_host (H2D + kernels + D2H inside)."""
    bases = np.ascontiguousarray(bases).view(np.uint8)
    start = np.ascontiguousarray(start, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    nS = len(start)
    if out is None:
        out = np.empty((nS, 4 ** k), dtype=np.int32)
    assert out.dtype == np.int32 and out.flags.c_contiguous and out.size == nS * 4 ** k
    rc = lib().cfrk_count_dense_host(_np_ptr(bases), fmt, _np_ptr(start), _np_ptr(length), len(bases), nS,
                                     k, mode, device, _np_ptr(out))
    _check(rc, "cfrk_count_dense_host")
    return out


def kmer_main(data, start, length, k, device=0):
    """The reference operator (src/kmer_main.cu:20-128): codes layout, compat semantics."""
    return count_dense_host(data, start, length, k, MODE_COMPAT, FMT_CODES, device)


def count_dense_device(d_bases, d_start, d_length, nN, nS, k, d_freq, mode=MODE_COMPAT, fmt=FMT_CODES,
                       read_begin=0, read_end=None, chunk_size=0, first_read_index=0, stream=0):
    """Raw device pointers (ints); asynchronous on `stream` (a cudaStream_t as int)."""
    if read_end is None:
        read_end = nS
    rc = lib().cfrk_count_dense_device(d_bases, fmt, d_start, d_length, nN, nS, read_begin, read_end, k, mode,
                                       chunk_size, first_read_index, d_freq, stream)
    _check(rc, "cfrk_count_dense_device")


def count_dense_packed_device(d_codes, d_valid, d_start, d_length, nN, nS, k, d_freq, mode=MODE_COMPAT,
                              read_begin=0, read_end=None, chunk_size=0, first_read_index=0, stream=0):
    """Dense rows from packed 2-bit reads (the output of encode_2bit_device)."""
    if read_end is None:
        read_end = nS
    _check(lib().cfrk_count_dense_packed_device(d_codes, d_valid, d_start, d_length, nN, nS, read_begin, read_end,
                                                k, mode, chunk_size, first_read_index, d_freq, stream),
           "cfrk_count_dense_packed_device")


def encode_2bit_device(d_bases, n, d_codes, d_valid, fmt=FMT_ASCII, stream=0):
    _check(lib().cfrk_encode_2bit_device(d_bases, fmt, n, d_codes, d_valid, stream), "cfrk_encode_2bit_device")


def global_hist_device(d_bases, d_start, d_length, nN, nS, k, d_hist, fmt=FMT_CODES, stream=0):
    _check(lib().cfrk_global_hist_device(d_bases, fmt, d_start, d_length, nN, nS, k, d_hist, stream),
           "cfrk_global_hist_device")


def count_sparse_device(d_bases, d_start, d_length, nN, nS, k, d_row_begin, d_row_count, d_keys, d_counts,
                        capacity, key_bytes=8, fmt=FMT_CODES, stream=0):
    """Sparse per-read rows (exact semantics); returns the total number of windows."""
    total = C.c_int64(0)
    _check(lib().cfrk_count_sparse_device(d_bases, fmt, d_start, d_length, nN, nS, k, key_bytes, d_row_begin,
                                          d_row_count, d_keys, d_counts, capacity, C.byref(total), stream),
           "cfrk_count_sparse_device")
    return total.value


def count_sparse_packed_device(d_codes, d_valid, d_start, d_length, nN, nS, k, d_row_begin, d_row_count, d_keys, d_counts,
                               capacity, key_bytes=8, stream=0):
    """Sparse per-read rows from packed 2-bit reads (the output of encode_2bit_device)."""
    total = C.c_int64(0)
    _check(lib().cfrk_count_sparse_packed_device(d_codes, d_valid, d_start, d_length, nN, nS, k, key_bytes, d_row_begin,
                                                 d_row_count, d_keys, d_counts, capacity, C.byref(total), stream),
           "cfrk_count_sparse_packed_device")
    return total.value


def scan_fasta_device(d_bytes, n, is_final, d_header, d_start, d_length, capacity, stream=0):
    """Record table of raw FASTA bytes on the GPU; returns the number of headers in the span."""
    nh = C.c_int64(0)
    _check(lib().cfrk_scan_fasta_device(d_bytes, n, int(is_final), d_header, d_start, d_length, capacity,
                                        C.byref(nh), stream), "cfrk_scan_fasta_device")
    return nh.value


def run_file(fasta, out, k, nt=12, chunk_size=8192, flags=0, device=0, devices=None):
    """cfrk <fasta> <out> <k> [nt] [chunkSize] (reference src/main.cu:232-305); devices=[0, 1, ...] spreads
    the file over several GPUs (rows still in read order)."""
    if devices is not None:
        arr = (C.c_int * len(devices))(*devices)
        _check(lib().cfrk_run_file_multi(os.fsencode(fasta), os.fsencode(out), k, nt, chunk_size, flags, arr, len(devices)),
               "cfrk_run_file_multi")
        return
    _check(lib().cfrk_run_file(os.fsencode(fasta), os.fsencode(out), k, nt, chunk_size, flags, device),
           "cfrk_run_file")
