"""Deterministic FASTA fixtures: the edge-case set of SURVEY.md Appendix A, re-created.

Every fixture is (name, fasta_text).  `reads_from_text` gives the reference-layout batch
(codes + start + length) through the ORACLE's parser, `ascii_batch` the raw-bytes batch the
GPU streamer uses (start/length into the file text itself).
"""
import random

import numpy as np


def _seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def fx_basic(seed=1, n=50, L=100):
    rng = random.Random(seed)
    return "".join(f">r{i} test\n{_seq(rng, L)}\n" for i in range(n))


def fx_with_n(seed=2, n=50, L=120):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        s = list(_seq(rng, L, "ACGTacgt"))
        for _ in range(rng.randint(0, 4)):
            s[rng.randrange(L)] = rng.choice("NnRYKM-")
        if i == 0:
            s[0] = "N"; s[57] = "N"
        if i == 7:
            s[-1] = "N"
        if i == 9:
            s[0] = s[1] = "N"
        out.append(f">r{i} test\n{''.join(s)}\n")
    return "".join(out)


def fx_multiline(seed=3, n=30):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        s = _seq(rng, rng.randint(200, 300))
        lines = "\n".join(s[j:j + 60] for j in range(0, len(s), 60))
        out.append(f">r{i} test\n{lines}\n")
    return "".join(out)


def fx_long(seed=4):
    rng = random.Random(seed)
    lens = [1023, 1024, 1025, 1026, 1500, 3000, 5000] * 3
    return "".join(f">r{i} test\n{_seq(rng, L)}\n" for i, L in enumerate(lens[:20]))


def fx_chunk(n, seed=5, L=80):
    rng = random.Random(seed + n)
    return "".join(f">r{i} test\n{_seq(rng, L)}\n" for i in range(n))


def fx_crlf(seed=6, n=12, L=90):
    rng = random.Random(seed)
    return "".join(f">r{i} test\r\n{_seq(rng, L)}\r\n" for i in range(n))


def fx_noeol(seed=7, n=12, L=90):
    return fx_basic(seed, n, L)[:-1]


def fx_blank(seed=8, n=12, L=90):
    return fx_basic(seed, n, L) + "\n\n\n"


def fx_short(seed=9):
    rng = random.Random(seed)
    lens = [1, 2, 3, 4, 5, 1, 1, 7, 2, 9]
    return "".join(f">r{i} test\n{_seq(rng, L)}\n" for i, L in enumerate(lens))


def fx_ragged(seed=10, n=300):
    """lengths 0..400 incl. empty sequence lines, some N, mixed case"""
    rng = random.Random(seed)
    out = []
    for i in range(n):
        L = rng.choice([0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 100, 150, 151, 400, rng.randint(0, 400)])
        s = list(_seq(rng, L, "ACGTacgt"))
        if L and rng.random() < 0.3:
            s[rng.randrange(L)] = "N"
        out.append(f">r{i}\n{''.join(s)}\n")
    # an EMPTY read makes the reference walk 1024 positions into its successors; at the end of
    # the batch that walk leaves the buffer (undefined in the reference), so close with clean reads
    out += [f">tail{i}\n{_seq(rng, 150)}\n" for i in range(10)]
    return "".join(out)


def fx_gt_in_header(seed=12, n=24, L=110):
    """'>' INSIDE header lines (">s3 A>G variant"): well defined in the reference -- grep -c counts
    lines, the parser only looks at line[0] (src/fastaIO.h:16,49)"""
    rng = random.Random(seed)
    heads = ["A>G variant", "x>y>z", ">", "trailing>", "chr1:100 C>T (rs1>2)"]
    return "".join(f">s{i} {heads[i % len(heads)]}\n{_seq(rng, L + i, 'ACGTacgtN')}\n" for i in range(n))


def fx_like_seq(n, L, seed=11):
    rng = random.Random(seed)
    return "".join(f">r{i} test\n{_seq(rng, L)}\n" for i in range(n))


EDGE_SET = [
    ("A_basic", fx_basic(), (1, 2, 3, 4, 6)),
    ("B_withN", fx_with_n(), (2, 3, 5)),
    ("C_multiline", fx_multiline(), (2, 4)),
    ("D_long", fx_long(), (2, 4)),
    ("F_crlf", fx_crlf(), (1, 2, 3, 5)),
    ("F_noeol", fx_noeol(), (1, 2, 3, 5)),
    ("F_blank", fx_blank(), (1, 2, 3, 5)),
    ("G_short", fx_short(), (1, 2, 3, 5)),
    ("R_ragged", fx_ragged(), (1, 2, 3, 4, 5, 6, 7, 8)),
    ("I_gtheader", fx_gt_in_header(), (2, 3, 5)),
]
CHUNK_SET = [  # (name, text, ks, chunk sizes)
    ("E_chunk20", fx_chunk(20), (2, 3), (8, 1, 8192, 7, 3)),
    ("E_chunk16", fx_chunk(16), (2,), (8, 1, 8192, 7, 3)),
]


def ascii_batch(text):
    """(bases uint8[n], start, length) describing records INSIDE the raw file bytes, the way
    cfrk_run_file does it (runfile.cu): text of a record = everything after its header line up
    to the next header, length = len(text) - 1."""
    raw = text.encode() if isinstance(text, str) else text
    starts, lengths = [], []
    pos = 0
    headers = []
    while pos < len(raw):
        nl = raw.find(b"\n", pos)
        end = len(raw) if nl < 0 else nl + 1
        if raw[pos:pos + 1] == b">":
            headers.append((pos, end))
        pos = end
    for i, (h, s) in enumerate(headers):
        e = headers[i + 1][0] if i + 1 < len(headers) else len(raw)
        starts.append(s)
        lengths.append(max(0, e - s - 1))
    return (np.frombuffer(raw, dtype=np.uint8).copy(), np.array(starts, dtype=np.int64),
            np.array(lengths, dtype=np.int32))


def ascii_compact(text):
    """reference batch layout (one separator per read, no header lines) holding LETTERS: what
    cfrk_run_file builds for spans with empty reads, and the layout bench.py generates"""
    import oracle_binding as ob
    data, start, length = ob.parse_fasta(text=text)
    lut = np.array([ord(c) for c in "ACGT"], dtype=np.uint8)
    raw = np.where(data >= 0, lut[np.clip(data, 0, 3)], ord("N")).astype(np.uint8)
    for s_, l_ in zip(start, length):
        raw[s_ + l_] = ord("\n")
    return raw, start, length


def synthetic_codes(nS, L, seed=42, n_frac=0.0):
    """Reference-layout batch of nS reads x L bases, uniform ACGT, optional fraction of -1."""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 4, size=(nS, L + 1), dtype=np.int8)
    if n_frac > 0:
        data[rng.random((nS, L + 1)) < n_frac] = -1
    data[:, L] = -1
    start = np.arange(nS, dtype=np.int64) * (L + 1)
    length = np.full(nS, L, dtype=np.int32)
    return data.reshape(-1), start, length
