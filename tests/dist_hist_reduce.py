"""Run under torchrun (tests/test_gpu_multi.py, tools/gpu/*.sh): the in-place histogram sum over NVLink peer memory
(cfrk_hist_allreduce_device) == NCCL all-reduce of the same tables == the oracle's histogram of all reads."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import cfrk_b200 as cf                      # noqa: E402
from cfrk_b200.sharding import HistReducer  # noqa: E402
import fixtures as fx                       # noqa: E402
import oracle_binding as ob                 # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out = {"world": world}
    modes = ["p2p"]
    for k in (3, 8, 12):
        data, start, length = fx.synthetic_codes(4000 + 100 * rank, 150, seed=10 * k + rank, n_frac=0.002)
        nN = len(data)
        bases = torch.full((nN + 16,), 0xFF, dtype=torch.uint8, device=dev)
        bases[:nN] = torch.from_numpy(data.view(np.uint8).copy()).to(dev)
        d_s, d_l = torch.from_numpy(start).to(dev), torch.from_numpy(length).to(dev)
        want = torch.from_numpy(ob.global_hist(data, start, length, k).astype(np.int64)).to(dev)
        dist.all_reduce(want)                                       # oracle histograms of all ranks, summed by NCCL
        red = HistReducer(4 ** k, dev)
        if k == 3 and red.multicast:
            modes.append("nvls")
        out["multicast"] = bool(red.multicast)
        for mode in modes:
            red.mode = mode
            for rep in range(3):                                    # epochs: the flag words are reused
                red.table.zero_()
                cf.global_hist_device(bases.data_ptr(), d_s.data_ptr(), d_l.data_ptr(), nN, len(start), k, red.table.data_ptr(),
                                      fmt=cf.FMT_CODES, stream=stream)
                local = red.table.clone()
                red.allreduce(stream)
                torch.cuda.synchronize()
                assert red.status() == 0, f"a peer did not show up: {red.status()}"
                dist.all_reduce(local)
                assert torch.equal(red.table, local), f"k={k} {mode} rep={rep}: != NCCL all-reduce"
                assert torch.equal(red.table.to(torch.int64), want), f"k={k} {mode} rep={rep}: != oracle"
        if k == 12:
            for mode in modes + ["nccl"]:
                red.mode = mode if mode != "nccl" else "p2p"
                t = red.table
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for it in range(25):
                    if it == 5:
                        dist.barrier()
                        torch.cuda.synchronize()
                        e0.record()
                    if mode == "nccl":
                        dist.all_reduce(t)
                    else:
                        red.allreduce(stream)
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                out[f"allreduce_ms_{mode}_64MiB"] = round(float(ms), 4)
        del red
    dist.barrier()
    if rank == 0:
        print("HIST_REDUCE_OK " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
