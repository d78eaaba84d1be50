"""GPU: the stand-alone stages of the pipeline (2-bit encode, whole-dataset histogram) and the
device-resident operator with read ranges / chunk openers, against the oracle."""
import numpy as np
import pytest
import torch

import cfrk_b200 as cf
import fixtures as fx
import oracle_binding as ob

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def padded_bases(raw, fill):
    t = torch.full((len(raw) + 16,), fill, dtype=torch.uint8, device="cuda")
    t[: len(raw)] = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint8)).cuda()
    return t


@pytest.mark.parametrize("fmt", ["ascii", "codes"])
def test_encode_2bit(fmt):
    text = fx.fx_with_n() + fx.fx_multiline()
    if fmt == "ascii":
        raw = np.frombuffer(text.encode(), dtype=np.uint8)
        lut = np.full(256, -1, dtype=np.int8)          # src/fastaIO.h:123-139
        for ch, v in zip("ACGTacgt", [0, 1, 2, 3] * 2):
            lut[ord(ch)] = v
        want_code = lut[raw]
        bases = padded_bases(raw, 0)
        f = cf.FMT_ASCII
    else:
        data, _, _ = ob.parse_fasta(text=text)
        raw, want_code = data.view(np.uint8), data
        bases = padded_bases(raw, 0xFF)
        f = cf.FMT_CODES
    n = len(raw)
    nb = (n + 15) // 16
    codes = torch.zeros(nb, dtype=torch.int32, device="cuda")
    valid = torch.zeros(nb, dtype=torch.int16, device="cuda")
    cf.encode_2bit_device(bases.data_ptr(), n, codes.data_ptr(), valid.data_ptr(), fmt=f)
    torch.cuda.synchronize()
    c = codes.cpu().numpy().view(np.uint32)
    v = valid.cpu().numpy().view(np.uint16)
    pos = np.arange(n)
    got_valid = (v[pos // 16] >> (15 - pos % 16)) & 1
    got_code = (c[pos // 16] >> (2 * (15 - pos % 16))) & 3
    np.testing.assert_array_equal(got_valid, (want_code >= 0).astype(np.uint16))
    np.testing.assert_array_equal(got_code[want_code >= 0], want_code[want_code >= 0].astype(np.uint32))
    assert (v[-1] & ((1 << (16 * nb - n)) - 1)) == 0   # nothing valid past the end of the buffer


@pytest.mark.parametrize("k", [1, 2, 4, 6, 8, 10, 12, 13])
def test_global_hist(k):
    data, start, length = ob.parse_fasta(text=fx.fx_with_n() + fx.fx_ragged() + fx.fx_long())
    want = ob.global_hist(data, start, length, k)
    hist = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
    b = padded_bases(data, 0xFF)
    s, l = dev(start), dev(length)
    cf.global_hist_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), len(start), k, hist.data_ptr())
    cf.global_hist_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), len(start), k, hist.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(hist.cpu().numpy().astype(np.uint64), 2 * want)   # accumulates


@pytest.mark.parametrize("k", [9, 10, 12, 13])
@pytest.mark.parametrize("layout", ["codes", "fasta_bytes", "skewed"])
def test_global_hist_split_path(k, layout, monkeypatch):
    """the partitioned histogram (hist_split.cu; default only for large batches) forced on small inputs:
    reference codes; raw FASTA bytes, where header lines full of ACGT letters lie between the reads and must
    not be counted; and a skewed batch whose k-mers overflow the fixed partition capacity (those suffixes
    are counted directly)"""
    monkeypatch.setenv("CFRK_HIST_SPLIT", "1")
    if layout == "fasta_bytes":
        import random
        rng = random.Random(k)
        recs = []
        for i in range(400):
            L = rng.choice([0, 1, k - 1, k, k + 1, 40, 150, 151, 700, 5000])
            seq = "".join(rng.choice("ACGT") for _ in range(L))
            if L > 30:
                seq = seq[:L // 2] + "N" + seq[L // 2 + 1:]
            recs.append(f">ACGTACGTACGTACGT_read{i}_GATTACA\n{seq}\n")
        data, start, length = fx.ascii_batch("".join(recs))     # the raw file bytes, headers included
        want = ob.global_hist(data, start, length, k, ascii=True)
        b, fmt, n = padded_bases(data, 0), cf.FMT_ASCII, len(data)
    else:
        if layout == "skewed":
            reads = ["A" * 30000, "ACGT" * 5000, "T" * 9000 + "N" + "T" * 9000] + ["AC" * 300] * 50
            data, start, length = ob.parse_fasta(text="".join(f">r{i}\n{r}\n" for i, r in enumerate(reads)))
        else:
            data, start, length = ob.parse_fasta(text=fx.fx_with_n() + fx.fx_ragged() + fx.fx_long())
        want = ob.global_hist(data, start, length, k)
        b, fmt, n = padded_bases(data, 0xFF), cf.FMT_CODES, len(data)
    hist = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
    s, l = dev(start), dev(length)
    for _ in range(2):
        cf.global_hist_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), n, len(start), k, hist.data_ptr(), fmt=fmt)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(hist.cpu().numpy().astype(np.uint64), 2 * want)   # accumulates


@pytest.mark.parametrize("k", [2, 4, 5, 6, 8])
def test_device_read_ranges_and_chunk_openers(k):
    """one launch over many reference chunks == one kmer_main call per chunk (src/main.cu:222,294,300);
    read ranges (multi-GPU shards, row rings) stitch together exactly"""
    nS = 3000 if k <= 6 else 600
    data, start, length = fx.synthetic_codes(nS, 97, seed=7 + k, n_frac=0.01)
    chunk = 256
    want = np.concatenate([
        ob.count_dense(data[start[a]:], start[a:a + chunk] - start[a], length[a:a + chunk], k)
        for a in range(0, nS, chunk)])
    b, s, l = padded_bases(data, 0xFF), dev(start), dev(length)
    out = torch.full((nS, 4 ** k), -1, dtype=torch.int32, device="cuda")
    rpt = cf.dense_reads_per_tile(k)
    cuts = [0, 2 * rpt, 5 * rpt + 3, nS - 1, nS]      # ranges need no alignment
    for a, e in zip(cuts, cuts[1:]):
        cf.count_dense_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), nS, k,
                              out[a:].data_ptr(), read_begin=a, read_end=e, chunk_size=chunk)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    # first_read_index shifts the chunk phase
    want2 = np.concatenate([
        ob.count_dense(data[start[a]:], start[a:e] - start[a], length[a:e], k)
        for a, e in [(0, 56)] + [(x, min(nS, x + chunk)) for x in range(56, nS, chunk)]])
    cf.count_dense_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), nS, k, out.data_ptr(),
                          chunk_size=chunk, first_read_index=200)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want2)


def test_full_size_properties():
    """BASELINE size (10 M x 150 bp is bench.py's job); here 2 M reads: conservation laws that do not
    need the oracle: row sums, and a checksum against the CPU counter on a strided sample"""
    nS, L, k = 2_000_000, 150, 4
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    codes = torch.randint(0, 4, (nS, L + 1), dtype=torch.uint8, device="cuda", generator=g)
    codes[:, L] = 0xFF
    flat = torch.cat([codes.view(-1), torch.full((16,), 0xFF, dtype=torch.uint8, device="cuda")])
    start = torch.arange(nS, dtype=torch.int64, device="cuda") * (L + 1)
    length = torch.full((nS,), L, dtype=torch.int32, device="cuda")
    out = torch.empty((nS, 4 ** k), dtype=torch.int32, device="cuda")
    cf.count_dense_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, out.data_ptr())
    torch.cuda.synchronize()
    sums = out.sum(1)
    # compat: L-k+1 valid windows + (k-2) spilled in from the next read (none for the last read)
    assert int(sums[:-1].min()) == int(sums[:-1].max()) == (L - k + 1) + (k - 2)
    assert int(sums[-1]) == L - k + 1
    sel = np.arange(0, nS, 9973)
    h = codes.cpu().numpy().view(np.int8)
    for i in sel[:50]:
        hi = min(nS, i + 2)
        want = ob.count_dense(h[i:hi].reshape(-1), np.arange(hi - i) * (L + 1), np.full(hi - i, L), k)[0]
        np.testing.assert_array_equal(out[i].cpu().numpy(), want)
    # exact mode: global histogram == column sums of the rows
    cf.count_dense_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, out.data_ptr(),
                          mode=cf.MODE_EXACT)
    hist = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
    cf.global_hist_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, hist.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out.sum(0, dtype=torch.int64), hist.to(torch.int64))


@pytest.mark.parametrize("name", ["A_basic", "C_multiline", "F_crlf", "F_noeol", "F_blank", "G_short", "R_ragged", "I_gtheader"])
@pytest.mark.parametrize("final", [True, False])
def test_scan_fasta_device(name, final):
    """record table built on the GPU == the table the reference parser implies (fixtures.ascii_batch)"""
    text = dict((n, t) for n, t, _ in fx.EDGE_SET)[name]
    raw, start, length = fx.ascii_batch(text)
    n = len(raw)
    buf = padded_bases(raw, 0)
    cap = len(start) + 4
    hdr = torch.zeros(cap, dtype=torch.int64, device="cuda")
    st = torch.zeros(cap, dtype=torch.int64, device="cuda")
    ln = torch.zeros(cap, dtype=torch.int32, device="cuda")
    nh = cf.scan_fasta_device(buf.data_ptr(), n, final, hdr.data_ptr(), st.data_ptr(), ln.data_ptr(), cap)
    torch.cuda.synchronize()
    assert nh == len(start)
    want_hdr = [i for i in range(n) if raw[i] == ord(">") and (i == 0 or raw[i - 1] == 10)]
    np.testing.assert_array_equal(hdr.cpu().numpy()[:nh], want_hdr)
    m = nh if final else nh - 1
    np.testing.assert_array_equal(st.cpu().numpy()[:m], start[:m])
    np.testing.assert_array_equal(ln.cpu().numpy()[:m], length[:m])


def test_scan_fasta_device_errors():
    for bad, frag in ((b"ACGT\n>a\nAC\n", "before the first"), (b">a\nAC>GT\n", "inside a sequence line")):
        raw = np.frombuffer(bad, dtype=np.uint8)
        buf = padded_bases(raw, 0)
        z = torch.zeros(8, dtype=torch.int64, device="cuda")
        l = torch.zeros(8, dtype=torch.int32, device="cuda")
        with pytest.raises(cf.CfrkError) as e:
            cf.scan_fasta_device(buf.data_ptr(), len(raw), True, z.data_ptr(), z.clone().data_ptr(), l.data_ptr(), 8)
        assert e.value.code == -5 and frag in str(e.value)
    # capacity too small
    raw = np.frombuffer(b">a\nA\n>b\nC\n>c\nG\n", dtype=np.uint8)
    buf = padded_bases(raw, 0)
    z = torch.zeros(2, dtype=torch.int64, device="cuda")
    with pytest.raises(cf.CfrkError) as e:
        cf.scan_fasta_device(buf.data_ptr(), len(raw), True, z.data_ptr(), z.clone().data_ptr(),
                             torch.zeros(2, dtype=torch.int32, device="cuda").data_ptr(), 2)
    assert e.value.code == -1


@pytest.mark.parametrize("k,nS", [(6, 400_000), (7, 150_000), (8, 40_000)])
def test_big_rows_under_load(k, nS):
    """~10 GB of rows per launch: every TMA zero store must land before the reductions of its tile
    (a lost or clobbered count shows up in the row sums), and a strided sample equals the oracle"""
    L = 150
    g = torch.Generator(device="cuda"); g.manual_seed(k)
    codes = torch.randint(0, 4, (nS, L + 1), dtype=torch.uint8, device="cuda", generator=g)
    codes[:, L] = 0xFF
    flat = torch.cat([codes.view(-1), torch.full((16,), 0xFF, dtype=torch.uint8, device="cuda")])
    start = torch.arange(nS, dtype=torch.int64, device="cuda") * (L + 1)
    length = torch.full((nS,), L, dtype=torch.int32, device="cuda")
    out = torch.full((nS, 4 ** k), 7, dtype=torch.int32, device="cuda")     # poisoned
    for rep in range(2):
        cf.count_dense_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, out.data_ptr())
    torch.cuda.synchronize()
    sums = torch.empty(nS, dtype=torch.int64, device="cuda")
    for a in range(0, nS, 8192):
        sums[a:a + 8192] = out[a:a + 8192].sum(1, dtype=torch.int64)
    assert int(sums[:-1].min()) == int(sums[:-1].max()) == (L - k + 1) + (k - 2)
    assert int(sums[-1]) == L - k + 1
    assert int(out.min()) == 0
    h = codes.cpu().numpy().view(np.int8)
    for i in list(range(0, nS, nS // 23))[:23]:
        hi = min(nS, i + 2)
        want = ob.count_dense(h[i:hi].reshape(-1), np.arange(hi - i) * (L + 1), np.full(hi - i, L), k)[0]
        np.testing.assert_array_equal(out[i].cpu().numpy(), want)


@pytest.mark.parametrize("k", [1, 3, 4, 5, 6, 7, 8])
def test_packed_reads_layout(k):
    """encode once (2-bit codes + validity), count from the packed layout: same rows as from the bytes"""
    text = fx.fx_with_n() + fx.fx_ragged() + fx.fx_multiline()
    raw, start, length = fx.ascii_compact(text)
    data, _, _ = ob.parse_fasta(text=text)
    n = len(raw)
    b = padded_bases(raw, 0)
    nb = (n + 15) // 16
    codes = torch.zeros(nb + 1, dtype=torch.int32, device="cuda")
    valid = torch.zeros(nb + 1, dtype=torch.int16, device="cuda")
    cf.encode_2bit_device(b.data_ptr(), n, codes.data_ptr(), valid.data_ptr(), fmt=cf.FMT_ASCII)
    s, l = dev(start), dev(length)
    for mode in (cf.MODE_COMPAT, cf.MODE_EXACT):
        out = torch.full((len(start), 4 ** k), -3, dtype=torch.int32, device="cuda")
        cf.count_dense_packed_device(codes.data_ptr(), valid.data_ptr(), s.data_ptr(), l.data_ptr(), n, len(start), k,
                                     out.data_ptr(), mode=mode)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), ob.count_dense(data, start, length, k, mode))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("mode", [cf.MODE_COMPAT, cf.MODE_EXACT], ids=["compat", "exact"])
def test_no_write_outside_the_rows(k, mode):
    """guard rows before and after the output stay untouched (the reference stores Freq[-1];
    compute-sanitizer is closed on this pool, so out-of-bounds writes are checked this way)"""
    data, start, length = ob.parse_fasta(text=fx.fx_with_n() + fx.fx_ragged() + fx.fx_short())
    nS = len(start)
    bins = 4 ** k
    guard = max(1, (4096 + bins - 1) // bins)          # at least 16 KiB of guard on each side
    buf = torch.full(((nS + 2 * guard) * bins,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    out = buf[guard * bins:(guard + nS) * bins]
    b, s, l = padded_bases(data, 0xFF), dev(start), dev(length)
    cuts = [0, nS // 3, nS // 3 + 1, nS]
    for a, e in zip(cuts, cuts[1:]):
        cf.count_dense_device(b.data_ptr(), s.data_ptr(), l.data_ptr(), len(data), nS, k, out[a * bins:].data_ptr(),
                              mode=mode, read_begin=a, read_end=e)
    torch.cuda.synchronize()
    assert bool((buf[: guard * bins] == 0x5A5A5A5A).all()) and bool((buf[(guard + nS) * bins:] == 0x5A5A5A5A).all())
    np.testing.assert_array_equal(out.view(nS, bins).cpu().numpy(), ob.count_dense_fast(data, start, length, k, mode))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("k", [2, 4, 6])
def test_read_range_shards_on_two_gpus(k):
    """SURVEY 8e with the kernels (the gloo test does it with the oracle): the batch is resident on both GPUs, GPU g
    counts its read range [b_g, b_g+1) of the FULL batch (so the compat halo -- the first read of the next shard --
    is there), the rows gathered in order equal the one-GPU rows and the oracle's; the per-GPU histograms add up"""
    import threading
    from cfrk_b200.sharding import shard_bounds
    data, start, length = ob.parse_fasta(text=fx.fx_with_n() + fx.fx_ragged() + fx.fx_basic(n=300) + fx.fx_long())
    nS, nN = len(start), len(data)
    want = ob.count_dense_fast(data, start, length, k, ob.MODE_COMPAT)
    b = shard_bounds(length, 2, align=cf.dense_reads_per_tile(k))
    rows, hists, errs = [None, None], [None, None], []

    def work(g):
        try:
            with torch.cuda.device(g):
                dev_ = torch.device("cuda", g)
                bases = torch.full((nN + 16,), 0xFF, dtype=torch.uint8, device=dev_)
                bases[:nN] = torch.from_numpy(data.view(np.uint8).copy()).to(dev_)
                d_s, d_l = torch.from_numpy(start).to(dev_), torch.from_numpy(length).to(dev_)
                r0, r1 = b[g], b[g + 1]
                out = torch.empty(((r1 - r0), 4 ** k), dtype=torch.int32, device=dev_)
                cf.count_dense_device(bases.data_ptr(), d_s.data_ptr(), d_l.data_ptr(), nN, nS, k, out.data_ptr(),
                                      read_begin=r0, read_end=r1, stream=torch.cuda.current_stream().cuda_stream)
                h = torch.zeros(4 ** k, dtype=torch.int32, device=dev_)
                cf.global_hist_device(bases.data_ptr(), d_s[r0:r1].data_ptr(), d_l[r0:r1].data_ptr(), nN, r1 - r0, k, h.data_ptr(),
                                      stream=torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                rows[g], hists[g] = out.cpu().numpy(), h.cpu().numpy().astype(np.int64)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=work, args=(g,)) for g in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    np.testing.assert_array_equal(np.concatenate(rows), want)
    np.testing.assert_array_equal(hists[0] + hists[1], ob.global_hist(data, start, length, k).astype(np.int64))
