#!/usr/bin/env python
"""Config C5 (BASELINE.json configs[4]): whole-dataset k-mer histogram, k=12, reads sharded by range
over N GPUs, per-GPU global_hist_kernel then ONE NCCL all-reduce of the 4^k uint32 histogram.

    python tools/bench_hist.py [--k 12] [--reads 6666667]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_hist.py --gpus N

Strong scaling: the 1 Gbase data set is split across the ranks.  Prints one JSON line (rank 0).
The parity check (tests/test_gpu_stages.py::test_global_hist) is separate; here the all-reduced
histogram is checked against conservation: sum == number of valid windows.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cfrk_b200 as cf  # noqa: E402
from bench import make_reads_device  # noqa: E402
from cfrk_b200.sharding import allreduce_histogram  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--k", type=int, default=12)
    ap.add_argument("--reads", type=int, default=6_666_667)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nS = a.reads // world + (1 if rank < a.reads % world else 0)     # this rank's read range
    L = a.read_len
    flat, start, length = make_reads_device(torch, nS, L, 46 + rank, 0.0, "ascii", dev)
    hist = torch.zeros(4 ** a.k, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        hist.zero_()
        cf.global_hist_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, a.k,
                              hist.data_ptr(), fmt=cf.FMT_ASCII, stream=stream)
        allreduce_histogram(hist)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * a.steps + 1)]
    t_count = t_red = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        hist.zero_()
        ev[3 * s].record()
        cf.global_hist_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nS * (L + 1), nS, a.k,
                              hist.data_ptr(), fmt=cf.FMT_ASCII, stream=stream)
        ev[3 * s + 1].record()
        allreduce_histogram(hist)
        ev[3 * s + 2].record()
    e1.record()
    torch.cuda.synchronize()
    for s in range(a.steps):
        t_count += ev[3 * s].elapsed_time(ev[3 * s + 1])
        t_red += ev[3 * s + 1].elapsed_time(ev[3 * s + 2])
    total = torch.tensor([e0.elapsed_time(e1), t_count, t_red], device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    total_ms, count_ms, red_ms = (float(x) / a.steps for x in total)
    windows = a.reads * (L - a.k + 1)
    ok = int(hist.sum(dtype=torch.int64)) == windows
    if rank == 0:
        print(json.dumps({
            "metric": "Gbases/sec, whole-dataset k-mer histogram", "value": round(a.reads * L / total_ms / 1e6, 2),
            "unit": "Gbases/s", "n_gpus": world, "k": a.k, "reads": a.reads, "read_len": L, "scaling": "strong",
            "ms_per_step": round(total_ms, 3), "count_ms": round(count_ms, 3), "allreduce_ms": round(red_ms, 3),
            "atomics_per_s_per_gpu": round(windows / world / count_ms / 1e6, 2),
            "hist_bytes": 4 ** a.k * 4, "conservation_ok": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
