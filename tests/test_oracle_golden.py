"""Pin the oracle: the reference's own regression goldens (test/out-seq{1,2}.cfrk, k=2,
chunkSize 8192) through the stand-in inputs, and the outputs of the reference's own code (CPU
shim) on the edge-case set.  CPU only."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

import fixtures as fx
import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
STANDIN_MD5 = {"seq1": "9d4f15b00662fa665abb42dd1ef875ea", "seq2": "ca73637b93a5e5ba2dd03869827d1423"}


@pytest.fixture(scope="session")
def standins(tmp_path_factory):
    d = tmp_path_factory.mktemp("standins")
    subprocess.check_call([sys.executable, os.path.join(GOLD, "make_standins.py"), str(d)],
                          stdout=subprocess.DEVNULL)
    return d


@pytest.mark.parametrize("name", ["seq1", "seq2"])
def test_oracle_reproduces_reference_goldens(standins, tmp_path, name):
    fa = standins / f"{name}.standin.fasta"
    assert hashlib.md5(fa.read_bytes()).hexdigest() == STANDIN_MD5[name]
    out = tmp_path / "out.cfrk"
    assert ob.run_cli(str(fa), str(out), 2, 8192) == 0   # test/test.sh:13,17
    assert out.read_bytes() == open(os.path.join(GOLD, f"out-{name}.cfrk"), "rb").read()


def test_golden_format_facts():
    """16 tokens per row, rows separated by '\\n', no trailing newline."""
    for name, nrows in (("seq1", 7898), ("seq2", 710)):
        raw = open(os.path.join(GOLD, f"out-{name}.cfrk"), "rb").read()
        assert raw.count(b"\n") == nrows - 1 and not raw.endswith(b"\n")
        assert ob.read_cfrk(os.path.join(GOLD, f"out-{name}.cfrk"), 2).shape == (nrows, 16)


def _manifest():
    with open(os.path.join(GOLD, "ref_shim", "manifest.json")) as f:
        return json.load(f)


def _fixture_text(name):
    for n, text, *_ in list(fx.EDGE_SET) + list(fx.CHUNK_SET):
        if n == name:
            return text
    if name == "H_like_seq2":
        return fx.fx_like_seq(710, 151)
    raise KeyError(name)


@pytest.mark.parametrize("key", sorted(_manifest()))
def test_oracle_matches_reference_code(tmp_path, key):
    """sha256 of the oracle's .cfrk == sha256 of what the reference's own source produced
    (tests/golden/make_ref_fixtures.py)."""
    m = _manifest()[key]
    text = _fixture_text(key.split(".")[0])
    assert hashlib.sha256(text.encode()).hexdigest() == m["fasta_sha256"], "fixture generator drifted: regenerate"
    fa, out = tmp_path / "in.fa", tmp_path / "out.cfrk"
    fa.write_text(text)
    assert ob.run_cli(str(fa), str(out), m["k"], m["chunk"]) == 0
    data = out.read_bytes()
    assert len(data) == m["out_bytes"]
    assert hashlib.sha256(data).hexdigest() == m["out_sha256"]
    small = os.path.join(GOLD, "ref_shim", key + ".cfrk")
    if os.path.exists(small):
        assert data == open(small, "rb").read()


@pytest.mark.skipif(not os.path.exists(os.path.join(ob.ORACLE_DIR, "_ref", "cfrk_ref_cpu")),
                    reason="oracle/_ref/cfrk_ref_cpu not built (needs /root/reference)")
def test_live_reference_code_on_random_input(tmp_path):
    """when the reference build is present: run it live on a fresh random FASTA"""
    import random
    rng = random.Random(20261018)
    txt = "".join(f">r{i}\n" + "".join(rng.choice("ACGTacgtNn") if rng.random() < 0.1 else rng.choice("ACGT")
                                       for _ in range(rng.randint(0, 140))) + "\n" for i in range(60))
    fa = tmp_path / "r.fa"
    fa.write_text(txt)
    for k in (1, 2, 4, 6):
        ref, mine = tmp_path / f"ref{k}", tmp_path / f"or{k}"
        subprocess.check_call([os.path.join(ob.ORACLE_DIR, "_ref", "cfrk_ref_cpu"), str(fa), str(ref), str(k), "12", "25"])
        assert ob.run_cli(str(fa), str(mine), k, 25) == 0
        assert ref.read_bytes() == mine.read_bytes()
