"""Read-range sharding across the GPUs of one box.

Reads are independent units: rank g gets the contiguous range [bounds[g], bounds[g+1]) balanced by
bases, cut at multiples of the row-tile size so that every shard's rows line up.  Per-read outputs
need no collective; only the optional whole-dataset histogram is summed across ranks (NCCL on
GPUs, gloo in the CPU tests).  In compat mode rank g also reads the first read of rank g+1 (its
spill lands in g's last row): pass the full batch and a [read_begin, read_end) range to
cfrk_count_dense_device and the kernel does that by itself.
"""
import numpy as np


def shard_bounds(lengths, world, align=1):
    """bounds[0..world] with bounds[g] % align == 0 (except the last), balanced by sum(length+1)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    nS = len(lengths)
    cum = np.concatenate([[0], np.cumsum(lengths + 1)])
    total = cum[-1]
    bounds = [0]
    for g in range(1, world):
        r = int(np.searchsorted(cum, total * g / world, side="left"))
        r = min(nS, max(bounds[-1], (r + align // 2) // align * align))
        bounds.append(r)
    bounds.append(nS)
    return bounds


def allreduce_histogram(hist_tensor, group=None):
    """Sum the per-rank whole-dataset histograms (torch tensor, uint32 counts held as int32/int64)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist_tensor, op=dist.ReduceOp.SUM, group=group)
    return hist_tensor
