#!/usr/bin/env python
"""Parity + timing of the dense kernels for small k under the environment switches of the build
(CFRK_DENSE_LANE, CFRK_LANE_SPLIT_K3, CFRK_LANE_SPLIT_K4): the switches are read once per process, so
tests/test_gpu_lane.py and the A/B runs of profiles/r2_notes.md call this script in a subprocess.

    python tests/manual/lane_variants.py [--ks 1,2,3,4] [--time-reads 2000000]
Prints one JSON line per k: parity_ok over the edge fixtures + a 150-bp batch with N bases (compat and
exact, codes / ascii / packed), and Gbases/s on --time-reads reads (0 = no timing)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cfrk_b200 as cf  # noqa: E402
import fixtures as fx  # noqa: E402
import oracle_binding as ob  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_dense(bases_t, start, length, nN, k, mode, fmt, lo=0, hi=None, chunk=0, first=0):
    nS = len(start)
    hi = nS if hi is None else hi
    out = torch.full(((hi - lo) * 4 ** k + 64,), -1, dtype=torch.int32, device="cuda")
    d_start, d_length = dev(start), dev(length)     # named: a temporary would be freed (and reused) before the launch
    cf.count_dense_device(bases_t.data_ptr(), d_start.data_ptr(), d_length.data_ptr(), nN, nS, k, out.data_ptr(),
                          mode=mode, fmt=fmt, read_begin=lo, read_end=hi, chunk_size=chunk, first_read_index=first)
    torch.cuda.synchronize()
    assert bool((out[(hi - lo) * 4 ** k:] == -1).all()), "wrote past the rows"
    return out[: (hi - lo) * 4 ** k].view(hi - lo, 4 ** k).cpu().numpy()


def pad(raw, fill):
    t = torch.full((len(raw) + 16,), fill, dtype=torch.uint8, device="cuda")
    t[: len(raw)] = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint8).copy()).cuda()
    return t


def parity(k):
    ok = True
    texts = [t for _, t, _ in fx.EDGE_SET]
    big = "".join(t if t.endswith("\n") else t + "\n" for t in texts)   # (a '>' may not land inside a line)
    for mode, omode in ((cf.MODE_COMPAT, ob.MODE_COMPAT), (cf.MODE_EXACT, ob.MODE_EXACT)):
        # codes layout (reference batch), whole edge set in one batch
        data, start, length = ob.parse_fasta(text=big)
        want = ob.count_dense_fast(data, start, length, k, omode)
        got = run_dense(pad(data, 0xFF), start, length, len(data), k, mode, cf.FMT_CODES)
        ok &= bool(np.array_equal(got, want))
        # ranges + chunk openers
        n = len(start)
        for lo, hi in ((0, n // 3), (n // 3, n // 3 + 37), (n // 3 + 37, n)):
            g2 = run_dense(pad(data, 0xFF), start, length, len(data), k, mode, cf.FMT_CODES, lo, hi)
            ok &= bool(np.array_equal(g2, want[lo:hi]))
        # ascii letters in the compact layout
        raw, s2, l2 = fx.ascii_compact(big)
        got = run_dense(pad(raw, 0), s2, l2, len(raw), k, mode, cf.FMT_ASCII)
        ok &= bool(np.array_equal(got, want))
        # synthetic 150-bp batch with N, chunk openers every 100 reads
        d3, s3, l3 = fx.synthetic_codes(20011, 150, seed=3 + k, n_frac=0.004)
        w3 = np.concatenate([ob.count_dense_fast(d3[a * 151:], s3[a:a + 100] - a * 151, l3[a:a + 100], k, omode)
                             for a in range(0, 20011, 100)]) if mode == cf.MODE_COMPAT else ob.count_dense_fast(d3, s3, l3, k, omode)
        g3 = run_dense(pad(d3, 0xFF), s3, l3, len(d3), k, mode, cf.FMT_CODES, chunk=100)
        ok &= bool(np.array_equal(g3, w3))
        # packed reads
        nb = (len(d3) + 15) // 16 + 8
        codes = torch.zeros(nb, dtype=torch.int32, device="cuda")
        valid = torch.zeros(nb, dtype=torch.int16, device="cuda")
        b3 = pad(d3, 0xFF)
        cf.encode_2bit_device(b3.data_ptr(), len(d3), codes.data_ptr(), valid.data_ptr(), fmt=cf.FMT_CODES)
        out = torch.empty((20011, 4 ** k), dtype=torch.int32, device="cuda")
        d_s3, d_l3 = dev(s3), dev(l3)
        cf.count_dense_packed_device(codes.data_ptr(), valid.data_ptr(), d_s3.data_ptr(), d_l3.data_ptr(), len(d3), 20011, k,
                                     out.data_ptr(), mode=mode, chunk_size=100)
        torch.cuda.synchronize()
        ok &= bool(np.array_equal(out.cpu().numpy(), w3))
    return ok


def timing(k, nS):
    from bench import make_reads_device, alg_bytes_per_read
    L = 150
    res = {}
    for name, nfrac in (("clean", 0.0), ("n0.001", 0.001)):
        flat, start, length = make_reads_device(torch, nS, L, 42, nfrac, "ascii", torch.device("cuda"))
        out = torch.empty(nS * 4 ** k, dtype=torch.int32, device="cuda")
        def go():
            cf.count_dense_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, out.data_ptr(),
                                  fmt=cf.FMT_ASCII, stream=torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            go()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            go()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[name] = {"gbases_s": round(nS * L / ms / 1e6, 1), "alg_gb_s": round(nS * alg_bytes_per_read(L, k) / ms / 1e6, 1)}
        del flat, start, length, out
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", default="1,2,3,4")
    ap.add_argument("--time-reads", type=int, default=0)
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    env = {k: v for k, v in os.environ.items() if k.startswith("CFRK_")}
    allok = True
    for k in [int(x) for x in a.ks.split(",")]:
        line = {"k": k, "env": env}
        if not a.no_parity:
            line["parity_ok"] = parity(k)
            allok &= line["parity_ok"]
        if a.time_reads:
            line["timing"] = timing(k, a.time_reads)
        print(json.dumps(line), flush=True)
    sys.exit(0 if allok else 1)
