"""ctypes view of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference path (oracle/cfrk_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module; the product package (cfrk_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

MODE_COMPAT = 0
MODE_EXACT = 1


class _Reads(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_int8)), ("length", C.POINTER(C.c_int32)),
                ("start", C.POINTER(C.c_int64)), ("nN", C.c_int64), ("nS", C.c_int64)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "cfrk_oracle.c")
        if (not os.path.exists(LIB_PATH)
                or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
            build()
        L = C.CDLL(LIB_PATH)
        p8, p32, p64 = C.c_void_p, C.c_void_p, C.c_void_p
        L.oracle_parse_fasta.argtypes = [C.c_char_p, C.POINTER(_Reads)]
        L.oracle_parse_fasta_mem.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Reads)]
        L.oracle_parse_fasta_ex.argtypes = [C.c_char_p, C.c_int, C.POINTER(_Reads)]
        L.oracle_parse_fasta_mem_ex.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_Reads)]
        L.oracle_free_reads.argtypes = [C.POINTER(_Reads)]
        for name in ("oracle_count_compat", "oracle_count_exact"):
            getattr(L, name).argtypes = [p8, p64, p32, C.c_int64, C.c_int64, C.c_int, p32]
            getattr(L, name).restype = None
        L.oracle_count_fast_mt.argtypes = [p8, p64, p32, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                           C.c_int, C.c_int, p32]
        L.oracle_count_fast_mt.restype = None
        L.oracle_global_hist.argtypes = [p8, p64, p32, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p]
        L.oracle_global_hist.restype = None
        L.oracle_count_sparse.argtypes = [p8, p64, p32, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_count_sparse.restype = C.c_int64
        L.oracle_write_cfrk.argtypes = [C.c_char_p, p32, C.c_int64, C.c_int]
        L.oracle_run_cli.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int64, C.c_int, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def parse_fasta(path=None, text=None, unwrap=False):
    """-> (data int8[nN], start int64[nS], length int32[nS]); raises ValueError on rc<0.
    unwrap: the intended reading (line terminators are not bases, last base kept)."""
    r = _Reads()
    if text is not None:
        if isinstance(text, str):
            text = text.encode()
        rc = lib().oracle_parse_fasta_mem_ex(text, len(text), int(unwrap), C.byref(r))
    else:
        rc = lib().oracle_parse_fasta_ex(os.fsencode(path), int(unwrap), C.byref(r))
    if rc != 0:
        raise ValueError(f"oracle_parse_fasta rc={rc}")
    nS, nN = r.nS, r.nN
    data = np.ctypeslib.as_array(r.data, shape=(max(nN, 1),))[:nN].copy()
    start = np.ctypeslib.as_array(r.start, shape=(max(nS, 1),))[:nS].copy()
    length = np.ctypeslib.as_array(r.length, shape=(max(nS, 1),))[:nS].copy()
    lib().oracle_free_reads(C.byref(r))
    return data, start, length


def _prep(data, start, length):
    data = np.ascontiguousarray(data).view(np.int8)
    start = np.ascontiguousarray(start, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    return data, start, length


def count_dense(data, start, length, k, mode=MODE_COMPAT):
    data, start, length = _prep(data, start, length)
    nS = len(start)
    freq = np.empty((nS, 4 ** k), dtype=np.int32)
    fn = lib().oracle_count_compat if mode == MODE_COMPAT else lib().oracle_count_exact
    if nS:
        fn(_ptr(data), _ptr(start), _ptr(length), len(data), nS, k, _ptr(freq))
    return freq


def count_dense_fast(data, start, length, k, mode=MODE_COMPAT, ascii=False, nthreads=0, out=None):
    data, start, length = _prep(data, start, length)
    nS = len(start)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    freq = out if out is not None else np.empty((nS, 4 ** k), dtype=np.int32)
    if nS:
        lib().oracle_count_fast_mt(_ptr(data), _ptr(start), _ptr(length), len(data), nS, k, mode,
                                   int(ascii), nthreads, _ptr(freq))
    return freq


def global_hist(data, start, length, k, ascii=False, nthreads=0):
    data, start, length = _prep(data, start, length)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    hist = np.zeros(4 ** k, dtype=np.uint64)
    lib().oracle_global_hist(_ptr(data), _ptr(start), _ptr(length), len(data), len(start), k,
                             int(ascii), nthreads, _ptr(hist))
    return hist


def count_sparse(data, start, length, k, ascii=False):
    data, start, length = _prep(data, start, length)
    nS = len(start)
    cap = int(np.maximum(length.astype(np.int64) - k + 1, 0).sum()) + 1
    row_ptr = np.zeros(nS + 1, dtype=np.int64)
    keys = np.zeros(cap, dtype=np.uint64)
    counts = np.zeros(cap, dtype=np.uint32)
    n = lib().oracle_count_sparse(_ptr(data), _ptr(start), _ptr(length), len(data), nS, k,
                                  int(ascii), _ptr(row_ptr), _ptr(keys), _ptr(counts), cap)
    assert n >= 0
    return row_ptr, keys[:n].copy(), counts[:n].copy()


def write_cfrk(path, freq, k):
    freq = np.ascontiguousarray(freq, dtype=np.int32)
    rc = lib().oracle_write_cfrk(os.fsencode(path), _ptr(freq), freq.shape[0], k)
    assert rc == 0


def run_cli(fasta, out, k, chunk_size=8192, mode=MODE_COMPAT, all_rows=False):
    return lib().oracle_run_cli(os.fsencode(fasta), os.fsencode(out), k, chunk_size, mode,
                                int(all_rows))


def read_cfrk(path, k):
    """Parse a dense .cfrk text file into int32[rows, 4^k] (checks the bin labels)."""
    with open(path, "rb") as f:
        txt = f.read()
    if not txt:
        return np.zeros((0, 4 ** k), dtype=np.int32)
    rows = []
    for line in txt.split(b"\n"):
        toks = line.split()
        vals = np.empty(len(toks), dtype=np.int32)
        for i, t in enumerate(toks):
            b, c = t.split(b":")
            assert int(b) == i
            vals[i] = int(c)
        rows.append(vals)
    return np.stack(rows)
