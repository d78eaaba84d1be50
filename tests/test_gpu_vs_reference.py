"""GPU: the NEW operator against THE REFERENCE'S OWN kmer_main on the same GPU, same inputs.

oracle/_ref/libcfrk_ref_gpu.so = the reference's unmodified kmer_main.cu + kmer_kernel.cu compiled
by oracle/Makefile for sm_100.  Both are called through the reference's operator contract
(struct read with pinned host buffers in, Freq out).  Skipped when the reference build is absent."""
import ctypes as C
import os

import numpy as np
import pytest

import cfrk_b200 as cf
import fixtures as fx
import oracle_binding as ob

pytestmark = pytest.mark.gpu
SO = os.path.join(ob.ORACLE_DIR, "_ref", "libcfrk_ref_gpu.so")


class RefRead(C.Structure):   # src/tipos.h:23-30
    _fields_ = [("data", C.c_void_p), ("length", C.c_void_p), ("start", C.c_void_p),
                ("Freq", C.c_void_p), ("next", C.c_void_p)]


def call_kmer_main(lib, data, start, length, k):
    nS, nN = len(start), len(data)
    rd = RefRead(data.ctypes.data, length.ctypes.data, start.ctypes.data, None, None)
    fn = getattr(lib, "_Z9kmer_mainP4readllit")
    fn.argtypes = [C.POINTER(RefRead), C.c_long, C.c_long, C.c_int, C.c_ushort]
    fn.restype = None
    fn(C.byref(rd), nN, nS, k, 0)
    assert rd.Freq
    out = np.ctypeslib.as_array(C.cast(rd.Freq, C.POINTER(C.c_int32)), shape=(nS, 4 ** k)).copy()
    if hasattr(lib, "ref_free_host"):      # the reference never frees rd->Freq (src/kmer_main.cu:115)
        lib.ref_free_host.argtypes = [C.c_void_p]
        lib.ref_free_host(rd.Freq)
    return out


@pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libcfrk_ref_gpu.so not built")
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12])
def test_same_inputs_same_rows(k):
    ref = C.CDLL(SO)
    ours = cf.lib()
    # nS * 4^k < 2^31 in the reference; k <= 12 keeps its float32 index exact (Q6)
    nS = 8192 if k <= 6 else {7: 1024, 8: 1024, 9: 1024, 10: 256, 12: 16}[k]
    # start with a read of length 1 (no visited window): the reference stores Freq[-1] for the
    # invalid windows of read 0, which we do not want to depend on
    data, start, length = fx.synthetic_codes(nS, 150, seed=100 + k, n_frac=0.002)
    length = length.copy(); length[0] = 1
    data = data.copy(); data[1] = -1
    want = call_kmer_main(ref, data, start, length, k)
    got = call_kmer_main(ours, data, start, length, k)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got, ob.count_dense_fast(data, start, length, k, ob.MODE_COMPAT))
