#!/usr/bin/env python
"""Configs C3/C4 (BASELINE.json configs[2], [3]): sparse per-read counts.
    python tools/bench_sparse.py --reads 10000000 --read-len 150 --k 12            (C3 shape, per GPU)
    python tools/bench_sparse.py --reads 100 --read-len 5000000 --k 21 --key-bytes 8   (C4 shape)
Prints one JSON line: Gbases/s and algorithmic GB/s (len + 8 + (key_bytes+4) * distinct per read)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cfrk_b200 as cf  # noqa: E402
from bench import make_reads_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--read-len", type=int, default=150)
ap.add_argument("--k", type=int, default=12)
ap.add_argument("--key-bytes", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
nS, L, k = a.reads, a.read_len, a.k
flat, start, length = make_reads_device(torch, nS, L, 44, 0.001, "ascii", dev)
cap = nS * (L - k + 1)
rb = torch.zeros(nS + 1, dtype=torch.int64, device=dev)
rc = torch.zeros(nS, dtype=torch.int32, device=dev)
keys = torch.zeros(cap, dtype=torch.int32 if a.key_bytes == 4 else torch.int64, device=dev)
cnt = torch.zeros(cap, dtype=torch.int32, device=dev)
def step():
    return cf.count_sparse_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, k, rb.data_ptr(),
                                  rc.data_ptr(), keys.data_ptr(), cnt.data_ptr(), cap, key_bytes=a.key_bytes, fmt=cf.FMT_ASCII,
                                  stream=torch.cuda.current_stream().cuda_stream)
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
distinct = int(rc.sum(dtype=torch.int64))
alg = nS * (L + 8) + distinct * (a.key_bytes + 4)
print(json.dumps({"metric": "Gbases/sec, sparse per-read k-mer counts", "value": round(nS * L / ms / 1e6, 2), "k": k,
                  "reads": nS, "read_len": L, "key_bytes": a.key_bytes, "ms": round(ms, 3), "distinct_pairs": distinct,
                  "alg_gb_s": round(alg / ms / 1e6, 1), "frac_of_6448": round(alg / ms / 1e6 / 6448.1, 4)}))
