"""world_size-2 gloo run of the host-side multi-GPU logic: read-range shards are a
partition, per-rank rows concatenate to the single-rank result (compat spill across the shard
boundary included), and the all-reduced per-rank histograms equal the global histogram.
The per-rank compute here is the ORACLE (CPU); on the GPU box the same logic drives the kernels
(bench.py --gpus N)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fixtures as fx

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, text, k, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle_binding as ob
    from cfrk_b200.sharding import allreduce_histogram, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data, start, length = ob.parse_fasta(text=text)
    b = shard_bounds(length, world, align=4)
    r0, r1 = b[rank], b[rank + 1]
    # rows of my shard: compat needs the next shard's first read as halo -> count [r0, r1+1) and drop the extra row
    hi = min(len(start), r1 + 1)
    rows = ob.count_dense(data, start[r0:hi], length[r0:hi], k, ob.MODE_COMPAT)[: r1 - r0]
    # chunk-local read 0 of the whole batch loses its spill; a shard's first read does not
    # (its spill belongs to the previous shard's last row, produced there through the halo)
    hist = torch.from_numpy(ob.global_hist(data, start[r0:r1], length[r0:r1], k).astype(np.int64))
    allreduce_histogram(hist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, rows))
    if rank == 0:
        ret["bounds"] = b
        ret["rows"] = np.concatenate([g[2] for g in gathered])
        ret["hist"] = hist.numpy()
    dist.destroy_process_group()


def test_two_rank_sharding_and_histogram_reduce():
    import oracle_binding as ob
    text = fx.fx_with_n() + fx.fx_ragged(seed=3, n=80)
    k = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29641, text, k, ret), nprocs=2, join=True)
    data, start, length = ob.parse_fasta(text=text)
    b = ret["bounds"]
    assert b[0] == 0 and b[-1] == len(start) and b[1] % 4 == 0 and 0 < b[1] < len(start)
    np.testing.assert_array_equal(ret["rows"], ob.count_dense(data, start, length, k, ob.MODE_COMPAT))
    np.testing.assert_array_equal(ret["hist"], ob.global_hist(data, start, length, k).astype(np.int64))


def test_shard_bounds_properties():
    from cfrk_b200.sharding import shard_bounds
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for align in (1, 16, 256):
            lengths = rng.integers(0, 400, size=10_000)
            b = shard_bounds(lengths, world, align)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(lengths)
            assert all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % align == 0 for x in b[:-1])
            if world > 1 and align == 1:
                w = [(lengths[x:y] + 1).sum() for x, y in zip(b, b[1:])]
                assert max(w) - min(w) <= 2 * 401
