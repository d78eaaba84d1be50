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


class HistReducer:
    """The whole-dataset histogram table of this rank in memory its peers can address, and the in-place sum over
    the ranks of one box by cfrk_hist_allreduce_device (one kernel per GPU over NVLink peer memory; hist_reduce.cu).

    torch.distributed is the plumbing only: symmetric allocation + address exchange (torch symmetric memory).
    `table` is the int32[4^k] tensor to count into (cfrk_global_hist_device adds into it); `allreduce()` makes every
    rank's table the sum.  mode: "p2p" (16-byte loads / stores between the GPUs), "nvls" (multimem instructions on the
    multicast address: the switch adds) or "auto" (nvls on 8 ranks when there is a multicast mapping -- measured
    0.177 vs 0.213 ms for 64 MiB -- else p2p: 0.126 vs 0.195 ms on 2 ranks)."""

    def __init__(self, n_bins, device, group=None, mode="auto"):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._torch, self._C, self._lib = torch, C, _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("HistReducer: at most 8 ranks (one box)")
        flag_words = 65536 // 4                      # CFRK_HIST_REDUCE_FLAG_BYTES
        self.n_bins = int(n_bins)
        # one symmetric buffer: the table, then the flag words (zero = no epoch yet)
        self._buf = symm.empty(self.n_bins + flag_words, dtype=torch.int32, device=device)
        self._buf.zero_()
        self._hdl = symm.rendezvous(self._buf, group.group_name)
        self.table = self._buf[: self.n_bins]
        self._flags = self._buf[self.n_bins:]
        ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        self._tables = (C.c_void_p * self.world)(*ptrs)
        self._flagp = (C.c_void_p * self.world)(*[p + 4 * self.n_bins for p in ptrs])
        self.multicast = int(self._hdl.multicast_ptr) if getattr(self._hdl, "has_multicast_support", False) and self._hdl.multicast_ptr else 0
        if mode == "nvls" and not self.multicast:
            raise RuntimeError("HistReducer: no multicast mapping on this system")
        self.mode = "nvls" if mode == "nvls" or (mode == "auto" and self.multicast and self.world >= 8) else "p2p"
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                          # every rank has zeroed its flag words before the first kernel

    def allreduce(self, stream=None):
        torch = self._torch
        self.epoch += 1
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        rc = self._lib.cfrk_hist_allreduce_device(self._tables, self._flagp, self.rank, self.world, self.n_bins, self.epoch,
                                                  self.multicast if self.mode == "nvls" else None, st)
        if rc != 0:
            raise RuntimeError("cfrk_hist_allreduce_device: " + self._lib.cfrk_last_error().decode())
        return self.table

    def status(self):
        """0, or 1 + the rank that did not show up at a meeting (after a synchronize)"""
        return int(self._flags[8192].item())         # CFRK_HIST_REDUCE_STATUS_WORD
