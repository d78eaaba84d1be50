"""GPU: sparse per-read rows (k up to 31) against the oracle's sort-based restatement."""
import numpy as np
import pytest
import torch

import cfrk_b200 as cf
import fixtures as fx
import oracle_binding as ob

pytestmark = pytest.mark.gpu


def run_sparse(data, start, length, k, key_bytes, fmt):
    nS = len(start)
    fill = 0xFF if fmt == cf.FMT_CODES else 0
    b = torch.full((len(data) + 16,), fill, dtype=torch.uint8, device="cuda")
    b[: len(data)] = torch.from_numpy(np.ascontiguousarray(data).view(np.uint8).copy()).cuda()
    s = torch.from_numpy(np.ascontiguousarray(start, dtype=np.int64)).cuda()
    l = torch.from_numpy(np.ascontiguousarray(length, dtype=np.int32)).cuda()
    cap = int(np.maximum(np.asarray(length, dtype=np.int64) - k + 1, 0).sum()) + 8
    rb = torch.zeros(nS + 1, dtype=torch.int64, device="cuda")
    rc = torch.full((nS,), -7, dtype=torch.int32, device="cuda")
    keys = torch.zeros(cap, dtype=torch.int32 if key_bytes == 4 else torch.int64, device="cuda")
    cnt = torch.zeros(cap, dtype=torch.int32, device="cuda")
    total = cf.count_sparse_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), nS, k, rb.data_ptr(),
                                   rc.data_ptr(), keys.data_ptr(), cnt.data_ptr(), cap, key_bytes=key_bytes, fmt=fmt)
    torch.cuda.synchronize()
    assert total == cap - 8
    kdt = np.uint32 if key_bytes == 4 else np.uint64
    return rb.cpu().numpy(), rc.cpu().numpy(), keys.cpu().numpy().view(kdt), cnt.cpu().numpy().view(np.uint32)


def check(data, start, length, k, key_bytes, fmt=cf.FMT_CODES, odata=None):
    rb, rc, keys, cnt = run_sparse(data, start, length, k, key_bytes, fmt)
    orp, okeys, ocnt = ob.count_sparse(data if odata is None else odata, start, length, k, ascii=False)
    nwin = np.maximum(np.asarray(length, dtype=np.int64) - k + 1, 0)
    np.testing.assert_array_equal(rb, np.concatenate([[0], np.cumsum(nwin)]))
    np.testing.assert_array_equal(rc, np.diff(orp))
    for i in range(len(start)):
        a, n = rb[i], rc[i]
        np.testing.assert_array_equal(keys[a:a + n].astype(np.uint64), okeys[orp[i]:orp[i + 1]], err_msg=f"row {i}")
        np.testing.assert_array_equal(cnt[a:a + n], ocnt[orp[i]:orp[i + 1]], err_msg=f"row {i}")


@pytest.mark.parametrize("k,key_bytes", [(3, 4), (9, 4), (12, 4), (16, 4), (12, 8), (17, 8), (21, 8), (31, 8)])
def test_short_reads(k, key_bytes):
    data, start, length = ob.parse_fasta(text=fx.fx_with_n() + fx.fx_ragged() + fx.fx_multiline())
    check(data, start, length, k, key_bytes)


@pytest.mark.parametrize("k,key_bytes", [(12, 4), (31, 8)])
def test_repeats_and_all_same(k, key_bytes):
    """low-complexity reads: few distinct k-mers with large counts, incl. the all-T read (max key)"""
    reads = ["T" * 150, "A" * 150, "AC" * 75, "ACGT" * 60, "T" * 40 + "N" + "T" * 60, "G" * 600, "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT",
             "C" * 2000, "ACGT" * 700, "T" * 4000 + "N" + "ACG" * 30]      # medium rows whose groups overflow -> long-row path
    text = "".join(f">r{i}\n{r}\n" for i, r in enumerate(reads))
    data, start, length = ob.parse_fasta(text=text)
    check(data, start, length, k, key_bytes)


@pytest.mark.parametrize("k,key_bytes", [(3, 4), (5, 8), (12, 4), (16, 4), (12, 8), (17, 8), (18, 8), (21, 8), (31, 8)])
def test_window_count_boundaries_and_long_reads(k, key_bytes):
    """reads around the 128/256/512-window network sizes and long reads (bucket path; uint64 keys:
    k=12,17 rows are all narrow (32-bit suffixes), k=18 mixes narrow and wide rows, k>=21 wide; k=3,5: fewer
    key bits than bucket bits, every bucket oversized -> radix-sort fallback)"""
    import random
    rng = random.Random(k)
    lens = [k - 1, k, k + 1, 127 + k, 128 + k, 255 + k, 256 + k, 257 + k, 511 + k, 512 + k - 1, 512 + k, 513 + k,
            3000, 20000, 70000]
    reads = []
    for L in lens:
        s = [rng.choice("ACGT") for _ in range(L)]
        if L > 40:
            s[L // 3] = "N"
        reads.append("".join(s))
    reads.append(("ACGTTGCA" * 4000)[:30000])      # long AND repetitive
    text = "".join(f">r{i}\n{r}\n" for i, r in enumerate(reads))
    data, start, length = ob.parse_fasta(text=text)
    check(data, start, length, k, key_bytes)


@pytest.mark.parametrize("k,key_bytes", [(7, 4), (12, 4), (16, 4), (12, 8), (21, 8), (31, 8)])
def test_two_reads_per_warp_class(k, key_bytes):
    """reads of <= 144 windows go two to a warp (sparse_half_kernel: 16 lanes x 9 keys each); window counts around
    the class border (143..146 -> the one-warp classes), odd read count (a half without a read), pairs of unequal
    length, reads without a valid window, low-complexity reads (the general run-length path) and all-distinct
    reads (its fast path) in one launch"""
    import random
    rng = random.Random(31 * k + key_bytes)
    reads = []
    for nwin in [1, 2, 8, 9, 10, 17, 18, 19, 63, 64, 100, 135, 139, 143, 144, 145, 146, 200, 256, 257, 300, 143, 144]:
        L = nwin + k - 1
        reads.append("".join(rng.choice("ACGT") for _ in range(L)))
    reads += ["A" * 150, "T" * (143 + k), "ACG" * 50, "N" * 150, "ACGT" * 30 + "N" + "TTGCA" * 6, "", "AC",
              "".join(rng.choice("ACGT") for _ in range(150)), ("AC" * 40 + "N") * 2]
    for _ in range(600):
        L = rng.choice([150, 151, 100, 144 + k - 1, 145 + k - 1, rng.randint(k, 175)])
        s_ = [rng.choice("ACGT") for _ in range(L)]
        if rng.random() < 0.3:
            s_[rng.randrange(L)] = "N"
        reads.append("".join(s_))
    if len(reads) % 2 == 0:
        reads.append("".join(rng.choice("ACGT") for _ in range(150)))     # odd count: the last warp has one read
    text = "".join(f">r{i}\n{r}\n" for i, r in enumerate(reads))
    data, start, length = ob.parse_fasta(text=text)
    check(data, start, length, k, key_bytes)


@pytest.mark.parametrize("k,key_bytes,batch", [(16, 4, 4000), (18, 8, 25000), (31, 8, 1)])
def test_long_rows_in_many_batches(k, key_bytes, batch, monkeypatch):
    """the long-row scratch is bounded by batches of rows: force tiny batches (several rows per batch,
    one row per batch, a row larger than the batch) and skewed rows (oversized buckets -> radix sort)"""
    import random
    rng = random.Random(100 + k)
    monkeypatch.setenv("CFRK_SPARSE_BATCH_KEYS", str(batch))
    reads = []
    for L in [700, 900, 5000, 150, 2500, 40000, 800, 12000, 600 + k]:
        reads.append("".join(rng.choice("ACGT") for _ in range(L)))
    reads.append("A" * 9000)                                   # one bucket holds everything
    reads.append(("ACGT" * 10 + "N") * 300)                    # few distinct k-mers, many invalid windows
    reads.append("N" * 2000)                                   # long row without a single valid window
    unit = "".join(rng.choice("ACGT") for _ in range(700))
    reads.append(unit * 12)                                    # every k-mer 11-12 times
    text = "".join(f">r{i}\n{r}\n" for i, r in enumerate(reads))
    data, start, length = ob.parse_fasta(text=text)
    check(data, start, length, k, key_bytes)


@pytest.mark.parametrize("k,key_bytes,L", [(12, 4, 150), (12, 8, 150), (9, 8, 168), (20, 8, 150), (2, 8, 150)])
def test_grouped_network_with_biased_reads(k, key_bytes, L):
    """reads of <= 160 windows with 64-bit keys take the grouped network (8 groups by the top 3 key bits); base composition
    skewed towards A fills group 0 beyond its 32 slots in part of the reads -> those take the full network:
    both paths, mixed in one launch, against the oracle"""
    rng = np.random.default_rng(k * 7 + L)
    nS = 6000
    p_a = rng.choice([0.25, 0.4, 0.55, 0.8], size=nS)
    data = np.full(nS * (L + 1), -1, dtype=np.int8)
    u = rng.random((nS, L))
    rest = (1 - p_a[:, None]) / 3
    codes = np.where(u < p_a[:, None], 0, 1 + np.minimum(2, ((u - p_a[:, None]) / rest).astype(np.int64))).astype(np.int8)
    codes[rng.random((nS, L)) < 0.002] = -1
    data.reshape(nS, L + 1)[:, :L] = codes
    start = np.arange(nS, dtype=np.int64) * (L + 1)
    length = np.full(nS, L, dtype=np.int32)
    check(data, start, length, k, key_bytes)


def test_ascii_input_and_150bp_batch():
    nS, L, k = 20000, 150, 12
    data, start, length = fx.synthetic_codes(nS, L, seed=12, n_frac=0.001)
    lut = np.array([65, 67, 71, 84], dtype=np.uint8)
    raw = np.where(data >= 0, lut[np.clip(data, 0, 3)], 78).astype(np.uint8)
    raw[start + length] = 10
    check(raw, start, length, k, 4, fmt=cf.FMT_ASCII, odata=data)


def test_capacity_error():
    data, start, length = fx.synthetic_codes(10, 150, seed=1)
    b = torch.from_numpy(np.concatenate([data.view(np.uint8), np.full(16, 255, np.uint8)])).cuda()
    s, l = torch.from_numpy(start).cuda(), torch.from_numpy(length).cuda()
    rb = torch.zeros(11, dtype=torch.int64, device="cuda")
    rc = torch.zeros(10, dtype=torch.int32, device="cuda")
    keys = torch.zeros(100, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(100, dtype=torch.int32, device="cuda")
    with pytest.raises(cf.CfrkError) as e:
        cf.count_sparse_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), 10, 12, rb.data_ptr(), rc.data_ptr(),
                               keys.data_ptr(), cnt.data_ptr(), 100)
    assert e.value.code == -1 and "capacity" in str(e.value)


def _device_batch(nS, L, seed, n_frac):
    """nS reads of L ASCII bases generated on the GPU (uniform ACGT, n_frac of them 'N'), '\\n' separated"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
    flat = torch.full((nS * (L + 1) + 16,), 10, dtype=torch.uint8, device="cuda")
    body = lut[torch.randint(0, 4, (nS, L), generator=g, device="cuda")]
    if n_frac > 0:
        body[torch.rand((nS, L), generator=g, device="cuda") < n_frac] = 78
    flat[: nS * (L + 1)].view(nS, L + 1)[:, :L] = body
    start = torch.arange(nS, dtype=torch.int64, device="cuda") * (L + 1)
    length = torch.full((nS,), L, dtype=torch.int32, device="cuda")
    return flat, start, length


@pytest.mark.parametrize("nS,L,k,key_bytes", [(2_000_000, 150, 12, 4), (30, 5_000_000, 21, 8), (12, 5_000_000, 31, 8),
                                               (2, 9_000_000, 21, 8), (1, 9_000_000, 31, 8)])
def test_scale_properties(nS, L, k, key_bytes):
    """configs C3/C4 at sizes the oracle cannot sort in a test: size-independent properties of every row
    (keys strictly increasing, counts >= 1 and summing to the row's valid windows, nothing written past the
    row) over the whole batch -- 150 M windows = two scratch batches at the default batch size -- plus the
    oracle on the first and the last row.  The 9 Mbp rows have 65536 buckets: more than the shared-memory
    counters of the counting pass hold, so they count with global atomics."""
    flat, start, length = _device_batch(nS, L, 7 + k, 0.001)
    nwin = L - k + 1
    cap = nS * nwin
    rb = torch.zeros(nS + 1, dtype=torch.int64, device="cuda")
    rc = torch.full((nS,), -7, dtype=torch.int32, device="cuda")
    kdt = torch.int32 if key_bytes == 4 else torch.int64
    keys = torch.full((cap,), -1, dtype=kdt, device="cuda")
    cnt = torch.zeros(cap, dtype=torch.int32, device="cuda")
    total = cf.count_sparse_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, rb.data_ptr(),
                                   rc.data_ptr(), keys.data_ptr(), cnt.data_ptr(), cap, key_bytes=key_bytes, fmt=cf.FMT_ASCII)
    torch.cuda.synchronize()
    assert total == cap
    assert torch.equal(rb, torch.arange(nS + 1, dtype=torch.int64, device="cuda") * nwin)
    rc64 = rc.to(torch.int64)
    assert int(rc64.min()) > 0 and int(rc64.max()) <= nwin
    # valid windows per row from the bases: a window is valid iff no 'N' among its k bases
    bad = (flat[: nS * (L + 1)].view(nS, L + 1)[:, :L] == 78).to(torch.int32)
    csum = torch.cumsum(torch.nn.functional.pad(bad, (1, 0)), dim=1)
    valid = ((csum[:, k:] - csum[:, :-k]) == 0).sum(dim=1)
    del bad, csum
    K = keys.view(nS, nwin)
    Cn = cnt.view(nS, nwin)
    col = torch.arange(nwin, device="cuda").unsqueeze(0)
    live = col < rc64.unsqueeze(1)
    assert torch.equal((Cn * live).sum(dim=1, dtype=torch.int64), valid)          # counts sum to the valid windows
    assert bool((Cn[live] >= 1).all())
    assert bool((Cn[~live] == 0).all()) and bool((K[~live] == -1).all())          # nothing behind the row's pairs
    # strictly increasing keys (unsigned order; keys use at most 62 bits, uint32 keys compared as int64)
    Kl = K.to(torch.int64) & 0xFFFFFFFF if key_bytes == 4 else K
    inc = (Kl[:, 1:] > Kl[:, :-1]) | ~live[:, 1:]
    assert bool(inc.all())
    assert int(Kl[live].max()) < 4 ** k and int(Kl[live].min()) >= 0
    # oracle on the first and the last row
    for r in (0, nS - 1):
        raw = flat[r * (L + 1): (r + 1) * (L + 1)].cpu().numpy()
        codes = np.full(L + 1, -1, dtype=np.int8)
        for ch, c in ((65, 0), (67, 1), (71, 2), (84, 3)):
            codes[:L][raw[:L] == ch] = c
        orp, okeys, ocnt = ob.count_sparse(codes, np.array([0], dtype=np.int64), np.array([L], dtype=np.int32), k, ascii=False)
        n = int(rc[r])
        assert n == orp[1]
        np.testing.assert_array_equal(K[r, :n].cpu().numpy().view(np.uint32 if key_bytes == 4 else np.uint64).astype(np.uint64), okeys)
        np.testing.assert_array_equal(Cn[r, :n].cpu().numpy().view(np.uint32), ocnt)


@pytest.mark.parametrize("k,key_bytes", [(12, 4), (21, 8), (31, 8)])
def test_packed_reads_input(k, key_bytes):
    """cfrk_count_sparse_packed_device: every size class (two reads per warp, one warp, one CTA, bucket path) reads the
    packed 2-bit words + validity masks of cfrk_encode_2bit_device instead of bytes; same rows as the oracle"""
    import random
    rng = random.Random(900 + k)
    reads = []
    for L in [150] * 300 + [k - 1, k, 100, 144 + k - 1, 145 + k - 1, 200, 300, 500, 512 + k, 700, 2000, 4000, 4500, 30000, 70000]:
        s_ = [rng.choice("ACGT") for _ in range(L)]
        if L > 20 and rng.random() < 0.5:
            s_[rng.randrange(L)] = "N"
        reads.append("".join(s_))
    rng.shuffle(reads)
    text = "".join(f">r{i}\n{r}\n" for i, r in enumerate(reads))
    data, start, length = ob.parse_fasta(text=text)
    nS, nN = len(start), len(data)
    b = torch.from_numpy(np.concatenate([data.view(np.uint8), np.full(16, 255, np.uint8)])).cuda()
    nb = (nN + 15) // 16 + 8
    codes = torch.zeros(nb, dtype=torch.int32, device="cuda")
    valid = torch.zeros(nb, dtype=torch.int16, device="cuda")
    cf.encode_2bit_device(b.data_ptr(), nN, codes.data_ptr(), valid.data_ptr(), fmt=cf.FMT_CODES)
    d_s, d_l = torch.from_numpy(start).cuda(), torch.from_numpy(length).cuda()
    cap = int(np.maximum(length.astype(np.int64) - k + 1, 0).sum()) + 1
    rb = torch.zeros(nS + 1, dtype=torch.int64, device="cuda")
    rc = torch.zeros(nS, dtype=torch.int32, device="cuda")
    keys = torch.zeros(cap, dtype=torch.int32 if key_bytes == 4 else torch.int64, device="cuda")
    cnt = torch.zeros(cap, dtype=torch.int32, device="cuda")
    cf.count_sparse_packed_device(codes.data_ptr(), valid.data_ptr(), d_s.data_ptr(), d_l.data_ptr(), nN, nS, k, rb.data_ptr(),
                                  rc.data_ptr(), keys.data_ptr(), cnt.data_ptr(), cap, key_bytes=key_bytes)
    torch.cuda.synchronize()
    orp, okeys, ocnt = ob.count_sparse(data, start, length, k)
    hrb, hrc = rb.cpu().numpy(), rc.cpu().numpy()
    hk = keys.cpu().numpy().view(np.uint32 if key_bytes == 4 else np.uint64).astype(np.uint64)
    hc = cnt.cpu().numpy().view(np.uint32)
    for i in range(nS):
        a, n = int(hrb[i]), int(hrc[i])
        assert n == orp[i + 1] - orp[i], f"row {i} len {length[i]}"
        np.testing.assert_array_equal(hk[a:a + n], okeys[orp[i]:orp[i + 1]], err_msg=f"row {i}")
        np.testing.assert_array_equal(hc[a:a + n], ocnt[orp[i]:orp[i + 1]], err_msg=f"row {i}")
