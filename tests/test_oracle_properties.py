"""Oracle self-consistency: the fast multithreaded counter (the CPU baseline) equals the slow
restatement; exact/compat algebra; histogram and sparse views; parser edge cases.  CPU only."""
import numpy as np
import pytest

import fixtures as fx
import oracle_binding as ob


@pytest.mark.parametrize("name,text,ks", fx.EDGE_SET, ids=[e[0] for e in fx.EDGE_SET])
def test_fast_equals_slow(name, text, ks):
    data, start, length = ob.parse_fasta(text=text)
    raw, rstart, rlength = fx.ascii_batch(text)
    for k in ks:
        for mode in (ob.MODE_COMPAT, ob.MODE_EXACT):
            slow = ob.count_dense(data, start, length, k, mode)
            for nt in (1, 3, 8):
                np.testing.assert_array_equal(ob.count_dense_fast(data, start, length, k, mode, nthreads=nt), slow)
            if mode == ob.MODE_COMPAT and (length == 0).any():
                raw, rstart, rlength = fx.ascii_compact(text)   # empty reads walk into what follows
            np.testing.assert_array_equal(ob.count_dense_fast(raw, rstart, rlength, k, mode, ascii=True), slow)


def test_compat_is_exact_plus_spill_for_clean_reads():
    """N-free reads with k <= len <= 1025: compat row i = exact row i + (k-2) * e_last owed by read i+1"""
    data, start, length = fx.synthetic_codes(200, 150, seed=5)
    for k in (2, 3, 5, 8):
        c = ob.count_dense(data, start, length, k, ob.MODE_COMPAT)
        e = ob.count_dense(data, start, length, k, ob.MODE_EXACT)
        want = e.copy()
        want[:-1, -1] += k - 2
        np.testing.assert_array_equal(c, want)
        assert (e.sum(1) == 150 - k + 1).all()


def test_window_cap_1024():
    """only the first 1024 window starts of a read are visited"""
    rng = np.random.default_rng(3)
    for L in (1023, 1024, 1025, 1026, 3000):
        data = np.concatenate([rng.integers(0, 4, L, dtype=np.int8), np.array([-1], np.int8)])
        c = ob.count_dense(data, np.array([0]), np.array([L]), 2)
        assert c.sum() == min(L - 1, 1024)


def test_global_hist_and_sparse_views():
    data, start, length = ob.parse_fasta(text=fx.fx_with_n())
    for k in (2, 5, 8):
        e = ob.count_dense(data, start, length, k, ob.MODE_EXACT)
        np.testing.assert_array_equal(ob.global_hist(data, start, length, k, nthreads=3), e.sum(0).astype(np.uint64))
        rp, keys, cnt = ob.count_sparse(data, start, length, k)
        for i in range(len(start)):
            nz = np.nonzero(e[i])[0]
            np.testing.assert_array_equal(keys[rp[i]:rp[i + 1]], nz.astype(np.uint64))
            np.testing.assert_array_equal(cnt[rp[i]:rp[i + 1]], e[i][nz].astype(np.uint32))
    rp, keys, cnt = ob.count_sparse(data, start, length, 21)   # 64-bit keys
    assert (np.diff(rp) >= 0).all() and rp[-1] == len(keys) and keys.max() < 4 ** 21


def test_parser_quirks():
    # len = strlen(text) - 1: last base lost without a final newline
    d, s, l = ob.parse_fasta(text=">a\nACGT")
    assert list(l) == [3]
    # newline-as-base in multi-line records, trailing blank lines extend the last read
    d, s, l = ob.parse_fasta(text=">a\nAC\nGT\n")
    assert list(l) == [5] and list(d[:6]) == [0, 1, -1, 2, 3, -1]
    d, s, l = ob.parse_fasta(text=">a\nACGT\n\n\n")
    assert list(l) == [6]
    # CRLF: '\r' is a base
    d, s, l = ob.parse_fasta(text=">a\r\nACGT\r\n")
    assert list(l) == [5] and d[4] == -1
    # a second '>' inside a header line is still one record (grep -c counts lines)
    d, s, l = ob.parse_fasta(text=">a >b\nAC\n")
    assert list(l) == [2]
    # header with no sequence line: defined here as an empty read
    d, s, l = ob.parse_fasta(text=">a\n>b\nAC\n")
    assert list(l) == [0, 2]
    for bad in ("ACGT\n>a\nAC\n", ">a\nAC>GT\n"):
        with pytest.raises(ValueError):
            ob.parse_fasta(text=bad)
    d, s, l = ob.parse_fasta(text="")
    assert len(l) == 0


def test_tail_only_and_all_rows(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_text(fx.fx_chunk(20))
    out = tmp_path / "o"
    ob.run_cli(str(fa), str(out), 2, 8)            # 20 reads, chunk 8 -> 4 rows
    assert ob.read_cfrk(str(out), 2).shape[0] == 4
    ob.run_cli(str(fa), str(out), 2, 10)           # multiple of chunk -> empty file
    assert out.read_bytes() == b""
    ob.run_cli(str(fa), str(out), 2, 8, all_rows=True)
    assert ob.read_cfrk(str(out), 2).shape[0] == 20
