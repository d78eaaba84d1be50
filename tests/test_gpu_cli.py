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


# ---- round 2: the span pipeline (host-cut spans + lookahead, several workers, several GPUs) -------------------------
def run_cfrk_env(env, *args):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([CFRK, *map(str, args)], capture_output=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr.decode()
    return r


@pytest.mark.parametrize("window", [4096, 20000, 65536])
@pytest.mark.parametrize("k", [2, 5])
def test_many_small_spans_equal_one_span(tmp_path, window, k):
    """CFRK_WINDOW_BYTES cuts a small file into many spans handled by two workers: rows, compat spill across span
    boundaries (reads with N at the cut), empty reads that walk into the next span's records, chunk openers and the
    tail-only pass must come out as if the file were one span"""
    text = fx.fx_with_n() + fx.fx_ragged() + fx.fx_long() + fx.fx_multiline() + fx.fx_gt_in_header() + fx.fx_short() + fx.fx_basic()
    fa = tmp_path / "in.fa"
    fa.write_text(text)
    got, want = tmp_path / "got", tmp_path / "want"
    for chunk, extra in ((8192, ["--all-rows"]), (37, ["--all-rows"]), (37, []), (8192, ["--all-rows", "--exact"]),
                         (16, ["--all-rows", "--sparse"])):
        run_cfrk_env({"CFRK_WINDOW_BYTES": str(window)}, fa, got, k, 3, chunk, *extra)
        run_cfrk(fa, want, k, 3, chunk, *extra)       # one span (the file is smaller than the default window)
        assert got.read_bytes() == want.read_bytes(), (chunk, extra)
    # and the one-span run is the oracle's
    assert ob.run_cli(str(fa), str(want), k, 37, ob.MODE_COMPAT, all_rows=True) == 0
    run_cfrk_env({"CFRK_WINDOW_BYTES": str(window)}, fa, got, k, 3, 37, "--all-rows")
    assert got.read_bytes() == want.read_bytes()


def test_record_larger_than_the_window(tmp_path):
    """a record larger than the streaming window used to be CFRK_EFORMAT; the pinned buffer now grows.
    Window 64 KiB, records of 300 kB and 1 MB between short reads, dense compat rows and sparse exact rows"""
    import random
    rng = random.Random(77)
    parts = []
    for i, L in enumerate([150, 300_000, 150, 150, 1_000_000, 80, 150]):
        s = "".join(rng.choice("ACGT") for _ in range(L))
        body = "\n".join(s[j:j + 80] for j in range(0, L, 80)) if L > 1000 else s
        parts.append(f">rec{i}\n{body}\n")
    text = "".join(parts)
    fa = tmp_path / "big.fa"
    fa.write_text(text)
    got, want = tmp_path / "got", tmp_path / "want"
    run_cfrk_env({"CFRK_WINDOW_BYTES": "65536"}, fa, got, 3, 4, 8192, "--all-rows")
    assert ob.run_cli(str(fa), str(want), 3, 8192, ob.MODE_COMPAT, all_rows=True) == 0
    assert got.read_bytes() == want.read_bytes()
    run_cfrk_env({"CFRK_WINDOW_BYTES": "65536"}, fa, got, 12, 4, 8192, "--all-rows", "--sparse", "--exact")
    data, start, length = ob.parse_fasta(text=text, unwrap=True)
    rp, keys, cnt = ob.count_sparse(data, start, length, 12)
    lines = got.read_bytes().split(b"\n")
    assert len(lines) == len(start)
    for i, line in enumerate(lines):
        toks = line.split()
        assert len(toks) == rp[i + 1] - rp[i]
        assert [int(t.split(b":")[0]) for t in toks[:50]] == [int(x) for x in keys[rp[i]:rp[i + 1]][:50]]
        assert sum(int(t.split(b":")[1]) for t in toks) == int(cnt[rp[i]:rp[i + 1]].sum())


def test_gzip_input(tmp_path):
    """the reference includes <zlib.h> and never uses it (src/fastaIO.h:7); here a .gz FASTA streams through gzread"""
    import gzip
    text = fx.fx_with_n() + fx.fx_chunk(20) + fx.fx_multiline()
    fa, gz = tmp_path / "in.fa", tmp_path / "in.fa.gz"
    fa.write_text(text)
    with gzip.open(gz, "wb") as f:
        f.write(text.encode())
    a, b = tmp_path / "a", tmp_path / "b"
    for args in ((3, 4, 8192, "--all-rows"), (3, 4, 16), (12, 4, 8192, "--all-rows", "--sparse", "--exact")):
        run_cfrk(fa, a, *args)
        run_cfrk_env({"CFRK_WINDOW_BYTES": "8192"}, gz, b, *args)
        assert a.read_bytes() == b.read_bytes() and (len(a.read_bytes()) > 0)


def test_sparse_rows_never_cross_pcie_dense(tmp_path):
    """--sparse for k = 5..8 compacts the dense rows to (bin, count) pairs on the device: same text as filtering the
    dense rows on the host (compat semantics, spill included)"""
    text = fx.fx_with_n() + fx.fx_long() + fx.fx_basic()
    fa = tmp_path / "in.fa"
    fa.write_text(text)
    for k in (5, 6, 7, 8):
        got, want = tmp_path / f"got{k}", tmp_path / f"want{k}"
        run_cfrk(fa, got, k, 4, 50, "--all-rows", "--sparse")
        assert ob.run_cli(str(fa), str(want), k, 50, ob.MODE_COMPAT, all_rows=True) == 0
        rows = ob.read_cfrk(str(want), k)
        lines = got.read_bytes().split(b"\n")
        assert len(lines) == rows.shape[0]
        for r, line in zip(rows, lines):
            nz = np.nonzero(r)[0]
            assert line == b"".join(b"%d:%d " % (int(i), int(r[i])) for i in nz)


def _device_count():
    import cfrk_b200 as cf
    return cf.device_count()


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs")
def test_two_gpus_same_bytes_as_one(tmp_path):
    """cfrk_run_file_multi / --devices: spans dealt to workers on two GPUs, rows gathered in read order: the output
    is byte-identical to the one-GPU run (E_chunk20, B_withN, and a 500 k-read file)"""
    rng = np.random.default_rng(21)
    nS, L = 500_000, 150
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)[np.minimum(4, rng.integers(0, 4000, size=(nS, L)) // 999)]
    big = tmp_path / "big.fa"
    with open(big, "wb") as f:
        for a in range(0, nS, 50_000):
            blk = letters[a:a + 50_000]
            hdr = np.frombuffer("".join(f">{i:09d}\n" for i in range(a, a + len(blk))).encode(), dtype=np.uint8).reshape(len(blk), 11)
            f.write(np.concatenate([hdr, blk, np.full((len(blk), 1), 10, np.uint8)], axis=1).tobytes())
    small = tmp_path / "small.fa"
    small.write_text(fx.fx_chunk(20) + fx.fx_with_n() + fx.fx_ragged())
    one, two = tmp_path / "one", tmp_path / "two"
    for fa, env, argsets in ((small, {"CFRK_WINDOW_BYTES": "8192"}, [(2, 4, 8), (3, 4, 8, "--all-rows"), (5, 4, 7, "--all-rows", "--sparse"),
                                                                     (12, 4, 8192, "--all-rows", "--sparse", "--exact")]),
                             (big, {"CFRK_WINDOW_BYTES": str(8 << 20)}, [(3, 8, 8192, "--all-rows"), (4, 8, 8192),
                                                                          (6, 8, 8192, "--all-rows", "--sparse")])):
        for args in argsets:
            run_cfrk_env(env, fa, one, *args, "--devices=0")
            run_cfrk_env(env, fa, two, *args, "--devices=0,1")
            assert hashlib.sha256(one.read_bytes()).hexdigest() == hashlib.sha256(two.read_bytes()).hexdigest(), (str(fa), args)
    assert ob.run_cli(str(big), str(one), 3, 8192, ob.MODE_COMPAT, all_rows=True) == 0
    run_cfrk_env({"CFRK_WINDOW_BYTES": str(8 << 20)}, big, two, 3, 8, 8192, "--all-rows", "--devices=0,1")
    assert hashlib.sha256(one.read_bytes()).hexdigest() == hashlib.sha256(two.read_bytes()).hexdigest()


def _fasta_to_fastq(text, seed=5):
    """single-line FASTA records -> 4-line FASTQ with quality strings that like to begin with '@' and '+'"""
    import random
    rng = random.Random(seed)
    out = []
    lines = text.split("\n")
    i = 0
    while i + 1 < len(lines):
        if lines[i].startswith(">"):
            seq = lines[i + 1]
            q = "".join(rng.choice("@+IIIIFF#:,") for _ in range(len(seq)))
            out.append(f"@{lines[i][1:]}\n{seq}\n+{lines[i][1:] if rng.random() < 0.5 else ''}\n{q}\n")
            i += 2
        else:
            i += 1
    return "".join(out)


@pytest.mark.parametrize("window", [0, 8192])
def test_fastq_input(tmp_path, window):
    """4-line FASTQ (the form SRR datasets come in, swift/roda.sh:3): the same rows as the FASTA file of the same
    reads, in every mode, across span cuts (quality lines that begin with '@' must not be taken for headers), with
    blank lines at the end of the file, and gzip-compressed"""
    import gzip
    text = fx.fx_with_n() + fx.fx_chunk(20) + fx.fx_long() + fx.fx_short() + fx.fx_basic()
    fa, fq, fqz = tmp_path / "in.fa", tmp_path / "in.fq", tmp_path / "in.fq.gz"
    fa.write_text(text)
    fq.write_text(_fasta_to_fastq(text) + "\n\n")
    with gzip.open(fqz, "wb") as f:
        f.write(_fasta_to_fastq(text).encode())
    env = {"CFRK_WINDOW_BYTES": str(window)} if window else {}
    a, b = tmp_path / "a", tmp_path / "b"
    for args in ((2, 4, 8192, "--all-rows"), (3, 4, 7, "--all-rows"), (3, 4, 7), (5, 4, 8192, "--all-rows", "--exact"),
                 (6, 4, 16, "--all-rows", "--sparse"), (12, 4, 8192, "--all-rows", "--sparse", "--exact")):
        run_cfrk(fa, a, *args)
        run_cfrk_env(env, fq, b, *args)
        assert a.read_bytes() == b.read_bytes(), args
        run_cfrk_env(env, fqz, b, *args)
        assert a.read_bytes() == b.read_bytes(), ("gz", args)
    bad = tmp_path / "bad.fq"
    bad.write_text("@r1\nACGT\nACGT\nIIII\n")
    r = subprocess.run([CFRK, str(bad), str(b), "2"], capture_output=True, timeout=120)
    assert r.returncode == 1 and b"FASTQ" in r.stderr


def test_error_in_a_later_span_stops_the_run(tmp_path):
    """a '>' inside a sequence line (undefined in the reference, src/fastaIO.h:16) far into the file: the worker that
    scans that span fails, the other workers and the reader stop, the command exits 1 with the message -- no hang"""
    text = fx.fx_basic(n=400) + ">bad\nACGT>ACGT\n" + fx.fx_basic(n=400, seed=3)
    fa = tmp_path / "in.fa"
    fa.write_text(text)
    e = dict(os.environ, CFRK_WINDOW_BYTES="8192")
    for extra in (["--all-rows"], []):
        r = subprocess.run([CFRK, str(fa), str(tmp_path / "o"), "3", "4", "50", *extra], capture_output=True, timeout=120, env=e)
        assert r.returncode == 1 and b"inside a sequence line" in r.stderr, r.stderr
