"""The .cfrk text writer (cfrk_write_rows = PrintFreq, /root/reference/src/main.cu:26-63, as a function): host only.
Both ways into the file -- rows formatted straight into the mapped output at offsets known from the rows' text sizes,
and private buffers + pwrite (CFRK_WRITER=pwrite; also what a pipe gets) -- against the oracle's writer, the
reference's golden files and a plain Python formatter."""
import os

import numpy as np
import pytest

import cfrk_b200 as cf
import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def py_text(rows, sparse):
    lines = []
    for row in rows:
        lines.append("".join(f"{b}:{v} " for b, v in enumerate(row) if not (sparse and v == 0)))
    return "\n".join(lines).encode()


def write(path, rows, k, nt, flags=0, writer=None):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    old = os.environ.pop("CFRK_WRITER", None)
    if writer:
        os.environ["CFRK_WRITER"] = writer
    try:
        rc = cf.lib().cfrk_write_rows(os.fsencode(str(path)), rows.ctypes.data, rows.shape[0], k, nt, flags)
    finally:
        os.environ.pop("CFRK_WRITER", None)
        if old is not None:
            os.environ["CFRK_WRITER"] = old
    assert rc == 0, cf.lib().cfrk_last_error()
    return open(path, "rb").read()


@pytest.mark.parametrize("writer", [None, "pwrite"], ids=["mapped", "pwrite"])
@pytest.mark.parametrize("name,nrows", [("seq1", 7898), ("seq2", 710)])
def test_golden_rows_round_trip(tmp_path, name, nrows, writer):
    """rows parsed from the reference's golden output -> text == the golden file, byte for byte"""
    gold = os.path.join(GOLD, f"out-{name}.cfrk")
    rows = ob.read_cfrk(gold, 2)
    assert rows.shape == (nrows, 16)
    for nt in (1, 3, 16):
        assert write(tmp_path / "o.cfrk", rows, 2, nt, writer=writer) == open(gold, "rb").read()


@pytest.mark.parametrize("writer", [None, "pwrite"], ids=["mapped", "pwrite"])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_rows_with_every_digit_count(tmp_path, k, writer):
    """one-digit rows (the template path), rows with 2..10-digit counts, all-zero rows (empty text when sparse),
    first / last row empty, fewer rows than threads; dense and sparse; == Python's formatting and the oracle's writer"""
    rng = np.random.default_rng(k)
    bins = 4 ** k
    n = 257 if k <= 4 else 19
    rows = rng.integers(0, 10, size=(n, bins)).astype(np.int32)
    rows[rng.random((n, bins)) < 0.5] = 0
    rows[3] = 0
    rows[5, :] = rng.integers(0, 2 ** 31 - 1, size=bins)
    rows[7, bins // 2] = 10
    rows[8, 0] = 123456
    rows[9, bins - 1] = 2 ** 31 - 1
    rows[n - 2] = 0
    for sparse in (False, True):
        for variant in ("plain", "first_empty", "last_empty"):
            r = rows.copy()
            if variant == "first_empty":
                r[0] = 0
            if variant == "last_empty":
                r[-1] = 0
            want = py_text(r, sparse)
            for nt in (1, 4, 16):
                got = write(tmp_path / "o.cfrk", r, k, nt, cf.RUN_SPARSE if sparse else 0, writer)
                assert got == want, (sparse, variant, nt)
    ob.write_cfrk(str(tmp_path / "oracle.cfrk"), rows, k)
    assert write(tmp_path / "o.cfrk", rows, k, 5, 0, writer) == open(tmp_path / "oracle.cfrk", "rb").read()
    # fewer rows than threads, one row, no row
    for m in (0, 1, 2):
        assert write(tmp_path / "o.cfrk", rows[:m], k, 16, 0, writer) == py_text(rows[:m], False)
        assert write(tmp_path / "o.cfrk", rows[:m], k, 16, cf.RUN_SPARSE, writer) == py_text(rows[:m], True)


def test_many_slices(tmp_path, monkeypatch):
    """several slices per call: k = 7 rows are 64 KiB, a 16 MiB row slot holds 256 of them -> 10 slices"""
    rng = np.random.default_rng(3)
    rows = (rng.random((2500, 4 ** 7)) < 0.01).astype(np.int32) * rng.integers(1, 30, size=(2500, 4 ** 7)).astype(np.int32)
    for sparse in (False, True):
        got = write(tmp_path / "o.cfrk", rows, 7, 8, cf.RUN_SPARSE if sparse else 0)
        assert got == py_text(rows, sparse)


def test_pipe_output(tmp_path):
    """a target that cannot be mapped or seeked (the reference driver's stdout form) gets the text sequentially"""
    import threading
    fifo = tmp_path / "fifo"
    os.mkfifo(fifo)
    rows = np.arange(5 * 16, dtype=np.int32).reshape(5, 16)
    got = {}

    def reader():
        with open(fifo, "rb") as f:
            got["text"] = f.read()
    t = threading.Thread(target=reader)
    t.start()
    rc = cf.lib().cfrk_write_rows(os.fsencode(str(fifo)), rows.ctypes.data, 5, 2, 4, 0)
    t.join()
    assert rc == 0
    assert got["text"] == py_text(rows, False)


def test_bad_arguments(tmp_path):
    L = cf.lib()
    rows = np.zeros((1, 4), dtype=np.int32)
    assert L.cfrk_write_rows(None, rows.ctypes.data, 1, 1, 1, 0) == -1
    assert L.cfrk_write_rows(os.fsencode(str(tmp_path / "x")), None, 1, 1, 1, 0) == -1
    assert L.cfrk_write_rows(os.fsencode(str(tmp_path / "x")), rows.ctypes.data, 1, 9, 1, 0) == -1
    assert L.cfrk_write_rows(os.fsencode(str(tmp_path / "nodir" / "x")), rows.ctypes.data, 1, 1, 1, 0) != 0
