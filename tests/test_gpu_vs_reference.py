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


# ---- the reference's own DRIVER on this hot path -----------------------------
# oracle/_ref/cfrk_ref_linked = the reference's unmodified main.cu + fastaIO.h + kmer.cuh + tipos.h, linked against
# libcfrk_b200.so instead of kmer_main.cu + kmer_kernel.cu (oracle/Makefile ref-link).  Its reader strcat()s into
# uninitialised malloc memory (src/fastaIO.h:51-52); MALLOC_PERTURB_=255 makes glibc hand out zero-filled blocks, which
# is what the reader silently assumes -- but only on the allocator's slow path: blocks recycled through the per-thread
# cache (tcache) keep their stale bytes (popen/pclose in GetNs and the CUDA runtime free plenty), so the cache is
# switched off as well.  One visible GPU: with devCount > 1 the driver skips chunks.
LINKED = os.path.join(ob.ORACLE_DIR, "_ref", "cfrk_ref_linked")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run_linked(fasta, out, k, chunk):
    import subprocess
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0", MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
    r = subprocess.run([LINKED, str(fasta), str(out), str(k), "12", str(chunk)], env=env, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stdout.decode()[-400:] + r.stderr.decode()[-400:]
    assert r.stdout == b"", r.stdout[:300]      # the reference prints its CUDA errors on stdout: none


@pytest.mark.skipif(not os.path.exists(LINKED), reason="oracle/_ref/cfrk_ref_linked not built")
@pytest.mark.parametrize("name", ["seq1", "seq2"])
def test_reference_driver_on_this_hot_path_reproduces_test_sh(tmp_path, name):
    """test/test.sh:13-19 with the reference's own main(): k=2, nt=12, chunk 8192, diff against the committed goldens"""
    import subprocess, sys
    subprocess.check_call([sys.executable, os.path.join(GOLD, "make_standins.py"), str(tmp_path)], stdout=subprocess.DEVNULL)
    out = tmp_path / "out.cfrk"
    _run_linked(tmp_path / f"{name}.standin.fasta", out, 2, 8192)
    assert out.read_bytes() == open(os.path.join(GOLD, f"out-{name}.cfrk"), "rb").read()


@pytest.mark.skipif(not os.path.exists(LINKED), reason="oracle/_ref/cfrk_ref_linked not built")
@pytest.mark.parametrize("key", ["A_basic.k1.c8192", "A_basic.k3.c8192", "A_basic.k6.c8192", "B_withN.k2.c8192", "B_withN.k5.c8192",
                                 "D_long.k4.c8192", "E_chunk20.k3.c7", "E_chunk20.k2.c8", "E_chunk16.k2.c8", "G_short.k3.c8192",
                                 "I_gtheader.k3.c8192", "H_like_seq2.k5.c8192"])
def test_reference_driver_on_this_hot_path_matches_reference_code(tmp_path, key):
    """single-line fixtures of the manifest (the reference's reader overflows its heap on wrapped records,
    src/fastaIO.h:59-60): same bytes as the reference's own kernels produce (tests/golden/ref_shim/manifest.json)"""
    import hashlib, json
    import fixtures as fx
    m = json.load(open(os.path.join(GOLD, "ref_shim", "manifest.json")))[key]
    name = key.split(".")[0]
    text = next((t for n, t, *_ in list(fx.EDGE_SET) + list(fx.CHUNK_SET) if n == name), None) or fx.fx_like_seq(710, 151)
    fa, out = tmp_path / "in.fa", tmp_path / "out.cfrk"
    fa.write_text(text)
    _run_linked(fa, out, m["k"], m["chunk"])
    data = out.read_bytes()
    assert len(data) == m["out_bytes"] and hashlib.sha256(data).hexdigest() == m["out_sha256"]
