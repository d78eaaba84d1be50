"""GPU: the `cfrk` command (bin/cfrk -> cfrk_run_file: streamer + kernels + writer) against the
reference's goldens, against the outputs of the reference's own code, and against the oracle."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import fixtures as fx
import oracle_binding as ob

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
CFRK = os.path.join(ROOT, "bin", "cfrk")


def run_cfrk(*args):
    r = subprocess.run([CFRK, *map(str, args)], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == b""   # the reference prints nothing on success
    return r


@pytest.fixture(scope="session")
def standins(tmp_path_factory):
    d = tmp_path_factory.mktemp("standins")
    subprocess.check_call([sys.executable, os.path.join(GOLD, "make_standins.py"), str(d)], stdout=subprocess.DEVNULL)
    return d


@pytest.mark.parametrize("name", ["seq1", "seq2"])
def test_reference_test_sh(standins, tmp_path, name):
    """reference test/test.sh:13-19: cfrk seqN.fasta out.cfrk 2 12 8192; diff out.cfrk out-seqN.cfrk"""
    out = tmp_path / "out.cfrk"
    run_cfrk(standins / f"{name}.standin.fasta", out, 2, 12, 8192)
    assert out.read_bytes() == open(os.path.join(GOLD, f"out-{name}.cfrk"), "rb").read()


def _manifest():
    with open(os.path.join(GOLD, "ref_shim", "manifest.json")) as f:
        return json.load(f)


def _fixture_text(name):
    for n, text, *_ in list(fx.EDGE_SET) + list(fx.CHUNK_SET):
        if n == name:
            return text
    return fx.fx_like_seq(710, 151)


@pytest.mark.parametrize("key", sorted(_manifest()))
def test_cli_matches_reference_code(tmp_path, key):
    m = _manifest()[key]
    fa, out = tmp_path / "in.fa", tmp_path / "out.cfrk"
    fa.write_text(_fixture_text(key.split(".")[0]))
    run_cfrk(fa, out, m["k"], 12, m["chunk"])
    data = out.read_bytes()
    assert len(data) == m["out_bytes"] and hashlib.sha256(data).hexdigest() == m["out_sha256"]


@pytest.mark.parametrize("k", [2, 5, 8])
def test_all_rows_exact_sparse_flags(tmp_path, k):
    fa = tmp_path / "in.fa"
    fa.write_text(fx.fx_with_n() + fx.fx_chunk(20))
    got, want = tmp_path / "got", tmp_path / "want"
    run_cfrk(fa, got, k, 4, 16, "--all-rows")
    assert ob.run_cli(str(fa), str(want), k, 16, ob.MODE_COMPAT, all_rows=True) == 0
    assert got.read_bytes() == want.read_bytes()
    run_cfrk(fa, got, k, 4, 16, "--all-rows", "--exact")
    assert ob.run_cli(str(fa), str(want), k, 16, ob.MODE_EXACT, all_rows=True) == 0
    assert got.read_bytes() == want.read_bytes()
    run_cfrk(fa, got, k, 4, 16, "--all-rows", "--sparse")
    dense = ob.read_cfrk(str(want.parent / "want"), k) if False else None
    ob.run_cli(str(fa), str(want), k, 16, ob.MODE_COMPAT, all_rows=True)
    rows = ob.read_cfrk(str(want), k)
    lines = got.read_bytes().split(b"\n")
    assert len(lines) == rows.shape[0]
    for r, line in zip(rows, lines):
        toks = dict((int(a), int(b)) for a, b in (t.split(b":") for t in line.split()))
        assert toks == {int(i): int(v) for i, v in enumerate(r) if v}


def test_streaming_many_buffers(tmp_path):
    """a file larger than one 64 MiB streaming window: records straddle buffers, held-back tail,
    chunk openers in the middle of a launch"""
    import numpy as np
    rng = np.random.default_rng(9)
    nS, L = 500_000, 150
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(nS, L))]
    fa = tmp_path / "big.fa"
    with open(fa, "wb") as f:
        for a in range(0, nS, 50_000):
            blk = letters[a:a + 50_000]
            hdr = np.frombuffer("".join(f">{i:09d}\n" for i in range(a, a + len(blk))).encode(), dtype=np.uint8).reshape(len(blk), 11)
            f.write(np.concatenate([hdr, blk, np.full((len(blk), 1), 10, np.uint8)], axis=1).tobytes())
    assert os.path.getsize(fa) > (64 << 20)
    got, want = tmp_path / "got", tmp_path / "want"
    run_cfrk(fa, got, 3, 8, 8192, "--all-rows")
    assert ob.run_cli(str(fa), str(want), 3, 8192, ob.MODE_COMPAT, all_rows=True) == 0
    assert hashlib.sha256(got.read_bytes()).hexdigest() == hashlib.sha256(want.read_bytes()).hexdigest()
    run_cfrk(fa, got, 4)             # default tail-only: 500000 % 8192 = 288 rows
    assert ob.run_cli(str(fa), str(want), 4, 8192) == 0
    assert got.read_bytes() == want.read_bytes() and got.read_bytes().count(b"\n") == 287


def test_swift_legacy_form(tmp_path):
    """swift/cfrk.swf:5: `cfrk <dataset> <k> <chunkSize>` with stdout captured"""
    fa = tmp_path / "in.fa"
    fa.write_text(fx.fx_basic())
    r = subprocess.run([CFRK, str(fa), "2", "4096"], capture_output=True, timeout=120)
    assert r.returncode == 0
    want = tmp_path / "want"
    ob.run_cli(str(fa), str(want), 2, 4096)
    assert r.stdout == want.read_bytes()


@pytest.mark.parametrize("k", [9, 12, 21, 31])
def test_sparse_rows_for_large_k(tmp_path, k):
    """k > 8: --sparse --exact rows = the oracle's sorted (k-mer, count) pairs"""
    fa = tmp_path / "in.fa"
    text = fx.fx_with_n() + fx.fx_long() + fx.fx_short()
    fa.write_text(text)
    out = tmp_path / "out.cfrk"
    run_cfrk(fa, out, k, 4, 8192, "--all-rows", "--sparse", "--exact")
    data, start, length = ob.parse_fasta(text=text)
    rp, keys, cnt = ob.count_sparse(data, start, length, k)
    lines = out.read_bytes().split(b"\n")
    assert len(lines) == len(start)
    for i, line in enumerate(lines):
        toks = [tuple(map(int, t.split(b":"))) for t in line.split()]
        assert toks == list(zip(keys[rp[i]:rp[i + 1]].tolist(), cnt[rp[i]:rp[i + 1]].tolist())), f"row {i}"
    r = subprocess.run([CFRK, str(fa), str(out), str(k)], capture_output=True)     # dense k > 8 is refused
    assert r.returncode == 1 and b"--sparse --exact" in r.stderr


@pytest.mark.parametrize("name", ["C_multiline", "F_crlf", "F_noeol", "F_blank", "B_withN", "R_ragged"])
@pytest.mark.parametrize("k", [3, 6, 11])
def test_exact_mode_unwraps_lines(tmp_path, name, k):
    """--exact reads FASTA the intended way (line terminators are not bases, last base kept):
    equals the oracle's unwrapped reading, dense (k <= 8) and sparse (k > 8)"""
    text = dict((n, t) for n, t, _ in fx.EDGE_SET)[name]
    fa, out = tmp_path / "in.fa", tmp_path / "out.cfrk"
    fa.write_text(text, newline="")
    data, start, length = ob.parse_fasta(text=text, unwrap=True)
    if k <= 8:
        run_cfrk(fa, out, k, 4, 8192, "--all-rows", "--exact")
        want = tmp_path / "want.cfrk"
        ob.write_cfrk(str(want), ob.count_dense(data, start, length, k, ob.MODE_EXACT), k)
        assert out.read_bytes() == want.read_bytes()
    else:
        run_cfrk(fa, out, k, 4, 8192, "--all-rows", "--exact", "--sparse")
        rp, keys, cnt = ob.count_sparse(data, start, length, k)
        lines = out.read_bytes().split(b"\n")
        assert len(lines) == len(start)
        for i, line in enumerate(lines):
            toks = [tuple(map(int, t.split(b":"))) for t in line.split()]
            assert toks == list(zip(keys[rp[i]:rp[i + 1]].tolist(), cnt[rp[i]:rp[i + 1]].tolist())), f"row {i}"


def test_wrapped_equals_single_line_in_exact_mode(tmp_path):
    import random
    rng = random.Random(4)
    seqs = ["".join(rng.choice("ACGTN" if rng.random() < 0.02 else "ACGT") for _ in range(rng.randint(1, 900))) for _ in range(200)]
    one = tmp_path / "one.fa"; wrapped = tmp_path / "wrapped.fa"
    one.write_text("".join(f">s{i}\n{s}\n" for i, s in enumerate(seqs)))
    wrapped.write_text("".join(f">s{i}\n" + "\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + "\n" for i, s in enumerate(seqs)))
    a, b = tmp_path / "a", tmp_path / "b"
    for k in (4, 7):
        run_cfrk(one, a, k, 4, 8192, "--all-rows", "--exact")
        run_cfrk(wrapped, b, k, 4, 8192, "--all-rows", "--exact")
        assert a.read_bytes() == b.read_bytes()


@pytest.mark.parametrize("k", [16, 21])
def test_genome_like_wrapped_fasta_sparse(tmp_path, k):
    """config C4 through the command line: a few long sequences, wrapped at 70 columns, N runs inside,
    --exact --sparse: the long-row path (bucket partition + warp sorts) behind the GPU-side unwrap"""
    import random
    rng = random.Random(40 + k)
    seqs = []
    for L in (200_000, 1, 150_000, 3_000, 260_000):
        s = [rng.choice("ACGT") for _ in range(L)]
        for _ in range(L // 50_000):
            p = rng.randrange(0, L - 200)
            s[p:p + 100] = "N" * 100
        seqs.append("".join(s))
    seqs[0] = seqs[0][:100_000] + seqs[0][20_000:120_000]          # a 100 kbp repeat: counts > 1
    text = "".join(f">chr{i} len={len(s)}\n" + "\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + "\n"
                   for i, s in enumerate(seqs))
    fa = tmp_path / "genome.fa"
    fa.write_text(text)
    out = tmp_path / "out.cfrk"
    run_cfrk(fa, out, k, 4, 8192, "--all-rows", "--sparse", "--exact")
    data, start, length = ob.parse_fasta(text=text, unwrap=True)
    rp, keys, cnt = ob.count_sparse(data, start, length, k)
    lines = out.read_bytes().split(b"\n")
    assert len(lines) == len(start)
    for i, line in enumerate(lines):
        toks = line.split()
        assert len(toks) == rp[i + 1] - rp[i], f"row {i}"
        got_k = np.array([int(t.split(b":")[0]) for t in toks], dtype=np.uint64)
        got_c = np.array([int(t.split(b":")[1]) for t in toks], dtype=np.uint32)
        np.testing.assert_array_equal(got_k, keys[rp[i]:rp[i + 1]], err_msg=f"row {i}")
        np.testing.assert_array_equal(got_c, cnt[rp[i]:rp[i + 1]], err_msg=f"row {i}")
