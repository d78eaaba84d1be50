#!/usr/bin/env python
"""Configs C3/C4 (BASELINE.json configs[2], [3]): sparse per-read counts.
    python tools/bench_sparse.py --reads 10000000 --read-len 150 --k 12            (C3 shape, per GPU)
    python tools/bench_sparse.py --reads 100 --read-len 5000000 --k 21 --key-bytes 8   (C4 shape)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sparse.py --reads 12500000 ...     (C3 sharded by read range: every rank counts its own
                                                        --reads reads, no data-path collective; weak scaling)
Prints one JSON line: Gbases/s (whole job, max time over ranks) and algorithmic GB/s
(len + 8 + (key_bytes+4) * distinct per read)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cfrk_b200 as cf  # noqa: E402
from bench import make_reads_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
ap.add_argument("--read-len", type=int, default=150)
ap.add_argument("--k", type=int, default=12)
ap.add_argument("--key-bytes", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
nS, L, k = a.reads, a.read_len, a.k
flat, start, length = make_reads_device(torch, nS, L, 44 + rank, 0.001, "ascii", dev)
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
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
distinct = int(rc.sum(dtype=torch.int64))
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # reporting only: the max over ranks
    ms = float(t.item())
    d = torch.tensor([distinct], dtype=torch.int64, device=dev)
    dist.all_reduce(d, op=dist.ReduceOp.SUM)
    distinct = int(d.item())
alg = world * nS * (L + 8) + distinct * (a.key_bytes + 4)
if rank == 0:
    print(json.dumps({"metric": "Gbases/sec, sparse per-read k-mer counts", "value": round(world * nS * L / ms / 1e6, 2), "k": k,
                      "n_gpus": world, "scaling": "weak", "reads_per_gpu": nS, "read_len": L, "key_bytes": a.key_bytes,
                      "ms": round(ms, 3), "distinct_pairs": distinct, "alg_gb_s": round(alg / ms / 1e6, 1),
                      "frac_of_6448_per_gpu": round(alg / ms / 1e6 / 6448.1 / world, 4)}))
if world > 1:
    dist.destroy_process_group()
