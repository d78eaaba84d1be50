"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import cfrk_b200 as cf
import fixtures as fx
import oracle_binding as ob

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,text,ks", fx.EDGE_SET, ids=[e[0] for e in fx.EDGE_SET])
@pytest.mark.parametrize("mode", [cf.MODE_COMPAT, cf.MODE_EXACT], ids=["compat", "exact"])
def test_edge_set_codes(name, text, ks, mode):
    data, start, length = ob.parse_fasta(text=text)
    for k in ks:
        want = ob.count_dense(data, start, length, k, mode)
        got = cf.count_dense_host(data, start, length, k, mode, cf.FMT_CODES)
        np.testing.assert_array_equal(got, want, err_msg=f"{name} k={k} mode={mode}")


@pytest.mark.parametrize("name,text,ks", fx.EDGE_SET, ids=[e[0] for e in fx.EDGE_SET])
@pytest.mark.parametrize("mode", [cf.MODE_COMPAT, cf.MODE_EXACT], ids=["compat", "exact"])
def test_edge_set_ascii(name, text, ks, mode):
    """raw file bytes + (start,length) into them == oracle on the parsed codes"""
    data, start, length = ob.parse_fasta(text=text)
    raw, rstart, rlength = fx.ascii_batch(text)
    np.testing.assert_array_equal(rlength, length)
    if mode == cf.MODE_COMPAT and (length == 0).any():
        # an empty read walks over the bytes that follow it: only the header-free layout matches
        raw, rstart, rlength = fx.ascii_compact(text)
    for k in ks:
        want = ob.count_dense(data, start, length, k, mode)
        got = cf.count_dense_host(raw, rstart, rlength, k, mode, cf.FMT_ASCII)
        np.testing.assert_array_equal(got, want, err_msg=f"{name} k={k} mode={mode}")


@pytest.mark.parametrize("k", range(1, 9))
def test_synthetic_150bp(k):
    nS = 20000 if k <= 6 else 3000
    data, start, length = fx.synthetic_codes(nS, 150, seed=42 + k, n_frac=0.001)
    for mode in (cf.MODE_COMPAT, cf.MODE_EXACT):
        want = ob.count_dense_fast(data, start, length, k, mode)
        got = cf.kmer_main(data, start, length, k) if mode == cf.MODE_COMPAT else \
            cf.count_dense_host(data, start, length, k, mode)
        np.testing.assert_array_equal(got, want)


def test_empty_and_tiny():
    z = cf.count_dense_host(np.zeros(0, np.int8), np.zeros(0, np.int64), np.zeros(0, np.int32), 3)
    assert z.shape == (0, 64)
    data = np.array([0, 1, 2, 3, -1], dtype=np.int8)
    got = cf.kmer_main(data, np.array([0]), np.array([4]), 2)
    want = ob.count_dense(data, np.array([0]), np.array([4]), 2)
    np.testing.assert_array_equal(got, want)


def test_concurrent_host_threads():
    """the reference driver calls kmer_main from several pthreads at once (src/main.cu:279-289);
    scratch buffers and streams are per host thread"""
    import threading
    data, start, length = fx.synthetic_codes(6000, 150, seed=77, n_frac=0.002)
    wants = {k: ob.count_dense_fast(data, start, length, k, ob.MODE_COMPAT) for k in (3, 5, 6, 7)}
    errs = []

    def work(k):
        try:
            for _ in range(3):
                got = cf.kmer_main(data, start, length, k)
                if not np.array_equal(got, wants[k]):
                    errs.append(f"k={k} mismatch")
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=work, args=(k,)) for k in (3, 5, 6, 7) for _ in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_pageable_and_unaligned_host_buffers():
    """host buffers need no alignment or pinning (the device copy is ours)"""
    data, start, length = fx.synthetic_codes(3000, 97, seed=5, n_frac=0.01)
    buf = np.empty(len(data) + 3, dtype=np.int8)
    view = buf[3:]           # misaligned host pointer
    view[:] = data
    for k in (2, 4, 8):
        np.testing.assert_array_equal(cf.kmer_main(view, start, length, k), ob.count_dense_fast(data, start, length, k))


@pytest.mark.parametrize("k", [1, 3, 4, 5, 6, 7, 8])
def test_long_sequences_dense(k):
    """sequences far beyond the reference's 1024-window cap: exact mode counts every window,
    compat mode stops at 1024; mixed with short reads so tiles hold both"""
    import random
    rng = random.Random(100 + k)
    lens = [200_000, 150, 1500, 0, 70_000, 3, 150, 150, 33_333]
    reads = []
    for L in lens:
        s = [rng.choice("ACGT") for _ in range(L)]
        for _ in range(L // 5000):
            s[rng.randrange(L)] = "N"
        reads.append("".join(s))
    reads += ["ACGT" * 40] * 20      # clean tail so the empty read's walk stays inside the buffer
    text = "".join(f">r{i}\n{r}\n" for i, r in enumerate(reads))
    data, start, length = ob.parse_fasta(text=text)
    for mode in (cf.MODE_EXACT, cf.MODE_COMPAT):
        want = ob.count_dense_fast(data, start, length, k, mode)
        got = cf.count_dense_host(data, start, length, k, mode)
        np.testing.assert_array_equal(got, want, err_msg=f"k={k} mode={mode}")
    if k <= 6:
        np.testing.assert_array_equal(ob.count_dense(data, start, length, k, cf.MODE_EXACT),
                                      ob.count_dense_fast(data, start, length, k, cf.MODE_EXACT))


@pytest.mark.parametrize("k", [9, 10, 12])
def test_dense_rows_up_to_k12(k):
    """the reference's operator accepts k up to 12 when nS * 4^k < 2^31"""
    nS = {9: 300, 10: 80, 12: 6}[k]
    data, start, length = fx.synthetic_codes(nS, 150, seed=k, n_frac=0.003)
    for mode in (cf.MODE_COMPAT, cf.MODE_EXACT):
        got = cf.count_dense_host(data, start, length, k, mode)
        want = ob.count_dense_fast(data, start, length, k, mode)
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("k,nS", [(5, 9000), (6, 8192), (7, 3000), (8, 700)])
@pytest.mark.parametrize("mode", [cf.MODE_COMPAT, cf.MODE_EXACT], ids=["compat", "exact"])
def test_host_operator_split_between_dma_and_host_threads(k, nS, mode):
    """cfrk_count_dense_host on batches large enough for the split: part of the rows arrive as dense rows by DMA, the
    rest as GPU-compacted (bin, count) pairs expanded by host threads (streaming stores); same rows as the oracle, whatever the share and
    the thread count, with N bases (spill across the DMA / host boundary), reads up to the 1024-window cap, a read
    without windows, and an output buffer that is not 64-byte aligned"""
    rng = np.random.default_rng(100 + k)
    lens = rng.choice([150, 151, 100, 1, 2, 37, 1030, 1500], size=nS, p=[0.6, 0.2, 0.1, 0.02, 0.02, 0.02, 0.02, 0.02]).astype(np.int32)
    start = np.concatenate([[0], np.cumsum(lens[:-1].astype(np.int64) + 1)])
    nN = int(start[-1] + lens[-1] + 1)
    data = rng.integers(0, 4, size=nN, dtype=np.int8)
    data[rng.random(nN) < 0.003] = -1
    data[start + lens] = -1
    omode = ob.MODE_COMPAT if mode == cf.MODE_COMPAT else ob.MODE_EXACT
    want = ob.count_dense_fast(data, start, lens, k, omode)
    try:
        for nt in (0, 1, 5, 16):
            cf.lib().cfrk_set_host_threads(nt)
            for rep in range(3):        # the share adapts between calls
                got = cf.count_dense_host(data, start, lens, k, mode, cf.FMT_CODES)
                np.testing.assert_array_equal(got, want, err_msg=f"nt={nt} rep={rep}")
        # unaligned rows (offset by 4 bytes): the in-cache path of the expansion
        raw = np.empty(nS * 4 ** k + 1, dtype=np.int32)
        out = raw[1:].reshape(nS, 4 ** k)
        got = cf.count_dense_host(data, start, lens, k, mode, cf.FMT_CODES, out=out)
        np.testing.assert_array_equal(got, want)
    finally:
        cf.lib().cfrk_set_host_threads(-1)


def test_release_and_reuse():
    """cfrk_release() frees the calling thread's context, the launch scratch and the cached pinned buffers; the next
    call builds them again (no unbounded growth, an explicit release); contexts of exited threads are
    taken over by new threads"""
    import threading
    data, start, length = fx.synthetic_codes(5000, 150, seed=9, n_frac=0.002)
    want = ob.count_dense_fast(data, start, length, 6, ob.MODE_COMPAT)
    L = cf.lib()
    for _ in range(3):
        np.testing.assert_array_equal(cf.kmer_main(data, start, length, 6), want)
        assert L.cfrk_release() == 0
    errs = []

    def work():
        try:
            if not np.array_equal(cf.kmer_main(data, start, length, 6), want):
                errs.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    for _ in range(6):          # thread churn: each new thread adopts the context the previous one left behind
        t = threading.Thread(target=work)
        t.start(); t.join()
    assert not errs, errs
    assert L.cfrk_release() == 0
    np.testing.assert_array_equal(cf.kmer_main(data, start, length, 6), want)
