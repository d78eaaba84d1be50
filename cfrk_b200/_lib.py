"""Loader of cfrk_b200/lib/libcfrk_b200.so.  Fails loudly: no fallback of any kind."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcfrk_b200.so")

# every symbol include/cfrk_b200.h declares (tests/test_abi.py checks header and library agree)
SYMBOLS = [
    "cfrk_version", "cfrk_last_error", "cfrk_device_count", "cfrk_launch_count", "cfrk_release", "cfrk_free_host",
    "cfrk_count_dense_host", "cfrk_set_host_threads", "cfrk_count_dense_device", "cfrk_count_dense_packed_device",
    "cfrk_dense_reads_per_tile",
    "cfrk_encode_2bit_device", "cfrk_global_hist_device", "cfrk_count_sparse_device", "cfrk_count_sparse_packed_device", "cfrk_scan_fasta_device",
    "cfrk_run_file", "cfrk_run_file_multi", "cfrk_write_rows", "cfrk_hist_allreduce_device",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a). "
            "cfrk_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.cfrk_version.restype = C.c_char_p
    L.cfrk_last_error.restype = C.c_char_p
    L.cfrk_device_count.restype = i32
    L.cfrk_launch_count.restype = C.c_uint64
    L.cfrk_release.restype = i32
    L.cfrk_set_host_threads.argtypes = [i32]
    L.cfrk_set_host_threads.restype = None
    L.cfrk_free_host.argtypes = [vp]
    L.cfrk_free_host.restype = None
    L.cfrk_dense_reads_per_tile.argtypes = [i32]
    L.cfrk_count_dense_host.argtypes = [vp, i32, vp, vp, i64, i64, i32, i32, i32, vp]
    L.cfrk_count_dense_device.argtypes = [vp, i32, vp, vp, i64, i64, i64, i64, i32, i32, i64, i64, vp, vp]
    L.cfrk_count_dense_packed_device.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i32, i32, i64, i64, vp, vp]
    L.cfrk_encode_2bit_device.argtypes = [vp, i32, i64, vp, vp, vp]
    L.cfrk_global_hist_device.argtypes = [vp, i32, vp, vp, i64, i64, i32, vp, vp]
    L.cfrk_count_sparse_device.argtypes = [vp, i32, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, i64,
                                           C.POINTER(C.c_int64), vp]
    L.cfrk_count_sparse_packed_device.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, i64,
                                                  C.POINTER(C.c_int64), vp]
    L.cfrk_scan_fasta_device.argtypes = [vp, i64, i32, vp, vp, vp, i64, C.POINTER(C.c_int64), vp]
    L.cfrk_run_file.argtypes = [C.c_char_p, C.c_char_p, i32, i32, i64, i32, i32]
    L.cfrk_hist_allreduce_device.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), i32, i32, i64, C.c_uint32, vp, vp]
    L.cfrk_write_rows.argtypes = [C.c_char_p, C.c_void_p, i64, i32, i32, i32]
    L.cfrk_run_file_multi.argtypes = [C.c_char_p, C.c_char_p, i32, i32, i64, i32, C.POINTER(C.c_int), i32]
    _lib = L
    return L
