#!/usr/bin/env python
"""bench.py -- Gbases/s of the CFRK per-read k-mer counting path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--k 4,5,6,7,8]

Workload (BASELINE.json configs[1], "C2"): 10 M synthetic reads x 150 bp, uniform ACGT, dense
per-read int32 counts for k = 4..8.  One STEP = one pass of the hot path over the whole read
set for every k in the sweep.  At N > 1 every rank owns its own 10 M reads (read-range
sharding, no data-path collective): weak scaling.

  value         device-resident: ASCII bases + offsets already in HBM, rows written to an HBM
                ring much larger than L2 (so they really go to DRAM); CUDA events on the
                launching stream; max over ranks.
  e2e           the same sweep through the reference-facing operator (cfrk_count_dense_host =
                kmer_main's contract: HOST buffers in, HOST rows out, H2D + kernels + D2H inside
                the timed region) on one reference chunk (8192 reads) per k.
  roofline      dominant kernel (largest share of the step): algorithmic bytes per launch
                (reads x (L + 8 + 4^k*4)) / its average launch time, against the
                measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the oracle's multithreaded CPU counter on the host cores (bounded sample).

  configs       (outside the headline's timed region, each with its own CUDA events and an ON-GPU
                CHECK in the same run): dense k = 1..3; C2 variant B (0.1 % N bases, seed 43: the
                data-dependent spill path); C3 = 100 M x 150 bp sparse per-read rows, k = 12, the
                read set SPLIT over the ranks (strong scaling); C4 = 1000 x 5 Mbp sparse rows,
                k = 16 / 21 / 31, split over the ranks; C5 = 1 Gbase whole-dataset histogram,
                k = 12, split over the ranks + NCCL all-reduce (count_ms / allreduce_ms).
  checks        dense, every k of the run, over ALL 10 M rows: column sums of the rows ==
                whole-dataset histogram of the visited windows + the spill total (computed from
                the lengths), and a read prefix bit-exact against the CPU oracle.

--impl reference runs the reference's own kmer_main (unmodified kmer_main.cu + kmer_kernel.cu
built by oracle/Makefile into oracle/_ref/libcfrk_ref_gpu.so) through the same host-buffer
contract on the same chunk; the reference is GPU-only, so this is its GPU path on the same
B200 (DESIGN.md "reference arm").  If that library is missing it times the CPU oracle.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbases/sec (k-mers counted/sec), per-read dense k-mer counting"
UNIT = "Gbases/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", default="4,5,6,7,8")
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--n-frac", type=float, default=0.0, help="fraction of bases replaced by N")
    ap.add_argument("--ring-gib", type=float, default=16.0)
    ap.add_argument("--e2e-reads", type=int, default=8192)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--mode", default="compat", choices=["compat", "exact"])
    ap.add_argument("--fmt", default="packed", choices=["ascii", "codes", "packed"],
                    help="packed: every step encodes the ASCII bases once (2-bit + validity) and counts every k from that")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip k=1..3, C2-B, C3, C4, C5")
    ap.add_argument("--no-checks", action="store_true", help="skip the on-GPU checks of the dense rows")
    ap.add_argument("--configs", default="small_k,c2b,c3,c4,c5", help="which extra configs to measure")
    ap.add_argument("--c3-reads", type=int, default=100_000_000, help="C3: total reads over all ranks")
    ap.add_argument("--c4-seqs", type=int, default=1000, help="C4: total sequences over all ranks")
    ap.add_argument("--c4-len", type=int, default=5_000_000)
    ap.add_argument("--c5-reads", type=int, default=6_666_667, help="C5: total reads over all ranks")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def alg_bytes_per_read(L, k):
    return L + 8 + 4 ** k * 4   # ASCII bases + offset + dense int32 row


# ------------------------------------------------------------------------------------------
def make_reads_device(torch, nS, L, seed, n_frac, fmt, device):
    """[nS, L+1] bytes: L bases + one separator, generated on the GPU in slabs."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    flat = torch.empty(nS * (L + 1) + 16, dtype=torch.uint8, device=device)
    flat[-16:] = 0xFF if fmt == "codes" else 0
    view = flat[: nS * (L + 1)].view(nS, L + 1)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    slab = 1 << 20
    for a in range(0, nS, slab):
        b = min(nS, a + slab)
        c = torch.randint(0, 4, (b - a, L + 1), dtype=torch.uint8, device=device, generator=g)
        if fmt == "ascii":
            c = lut[c.long()]
        if n_frac > 0:
            m = torch.rand((b - a, L + 1), device=device, generator=g) < n_frac
            c[m] = 78 if fmt == "ascii" else 0xFF
        c[:, L] = 10 if fmt == "ascii" else 0xFF
        view[a:b] = c
    start = torch.arange(nS, dtype=torch.int64, device=device) * (L + 1)
    length = torch.full((nS,), L, dtype=torch.int32, device=device)
    return flat, start, length


def make_reads_host(nS, L, seed, fmt):
    """One reference chunk in the reference layout.  Read 0 is given length 1 (no visited window):
    for k > 2 every read has windows that straddle its terminator, the reference adds those of
    read 0 at Freq[-1] (src/kmer_kernel.cu:84-87), and with a multi-GB Freq that store is an
    illegal memory access on a B200 (seen in round 1).  Both arms get the same chunk."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 4, size=(nS, L + 1), dtype=np.uint8)
    if fmt == "ascii":
        c = np.array([65, 67, 71, 84], dtype=np.uint8)[c]
        c[:, L] = 10
        c[0, 1] = 10
    else:
        c[:, L] = 0xFF
        c[0, 1] = 0xFF
    start = np.arange(nS, dtype=np.int64) * (L + 1)
    length = np.full(nS, L, dtype=np.int32)
    length[0] = 1
    return c.reshape(-1), start, length


def cpu_baseline(args, ks, L, budget_s):
    """oracle's multithreaded counter (all host cores) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob   # test infrastructure: allowed here as the CPU baseline only
    cores = os.cpu_count() or 1
    nS = args.e2e_reads
    data, start, length = make_reads_host(nS, L, 7, "codes")
    mode = ob.MODE_COMPAT if args.mode == "compat" else ob.MODE_EXACT
    outs = {k: np.empty((nS, 4 ** k), dtype=np.int32) for k in ks}
    for k in ks:   # warm (page-faults the output buffers)
        ob.count_dense_fast(data.view(np.int8), start, length, k, mode, nthreads=cores, out=outs[k])
    reps, t0 = 0, time.perf_counter()
    while True:
        for k in ks:
            ob.count_dense_fast(data.view(np.int8), start, length, k, mode, nthreads=cores, out=outs[k])
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or reps >= 2000:
            break
    bases = reps * len(ks) * nS * L
    return {"value": bases / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} x sweep k={ks} over {nS} reads x {L} bp (oracle_count_fast_mt, {cores} threads, "
                      f"{dt:.1f} s)"}


# ------------------------------------------------------------------------------------------
# On-GPU checks and the extra configs (every north-star config measured AND verified in
# the driver's run).  `ob` (the oracle binding) is used here as the CHECKER only.
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    return ob


def timed(torch, fn, warmup, reps):
    """mean ms of fn() over `reps` calls after `warmup`, CUDA events on the current stream"""
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dense_check(torch, cf, flat, start, length, nN, nS, k, mode, ascii_fmt, ring, plan, launch, stream, prefix_reads):
    """Over ALL rows of one k: column sums of the rows == histogram of the visited windows + the
    spill total; and rows [0, P) bit-exact against the CPU oracle.
    compat (src/kmer_kernel.cu:83-88): read i visits starts t < vis_i = min(len_i - 1, 1024); a visited
    window is counted in its own row if its k bases are valid and inside the read, else it adds 1 to
    the LAST bin of row i-1 (lost for i = 0).  So with tend_i = min(len_i, vis_i + k - 1):
        colsum = hist(reads cut to tend_i)  +  e_last * sum_{i >= 1} (vis_i - valid_i)."""
    bins = 4 ** k
    len64 = length.to(torch.int64)
    compat = mode == cf.MODE_COMPAT
    if compat:
        vis = torch.clamp(len64 - 1, min=0, max=1024)
        tend = torch.where(vis > 0, torch.minimum(len64, vis + (k - 1)), torch.zeros_like(vis)).to(torch.int32)
    else:
        vis = None
        tend = length
    fmt = cf.FMT_ASCII if ascii_fmt else cf.FMT_CODES
    hist = torch.zeros(bins, dtype=torch.int32, device=flat.device)
    cf.global_hist_device(flat.data_ptr(), start.data_ptr(), tend.data_ptr(), nN, nS, k, hist.data_ptr(), fmt=fmt, stream=stream)
    col = torch.zeros(bins, dtype=torch.int64, device=flat.device)
    for a, b in plan:
        launch(k, a, b)
        col += ring[: (b - a) * bins].view(b - a, bins).sum(dim=0, dtype=torch.int64)
    expect = hist.to(torch.int64)
    spill = 0
    if compat:
        h0 = torch.zeros(bins, dtype=torch.int32, device=flat.device)
        cf.global_hist_device(flat.data_ptr(), start.data_ptr(), tend.data_ptr(), nN, 1, k, h0.data_ptr(), fmt=fmt, stream=stream)
        spill = int(vis[1:].sum()) - (int(expect.sum()) - int(h0.sum(dtype=torch.int64)))
        expect[bins - 1] += spill
    colsum_ok = bool(torch.equal(col, expect))
    # prefix against the oracle (one kmer_main call: read 0 drops its spill, row P-1 takes read P's)
    ob = _oracle()
    P = int(min(prefix_reads, nS - 1, max(1, (1 << 30) // (bins * 4))))
    launch(k, 0, P)
    got = ring[: P * bins].view(P, bins).cpu().numpy()
    nbytes = int(start[P + 1].item()) if P + 1 < nS else nN
    hdata = flat[:nbytes].cpu().numpy()
    want = ob.count_dense_fast(hdata.view("int8"), start[: P + 1].cpu().numpy(), length[: P + 1].cpu().numpy(), k,
                               ob.MODE_COMPAT if compat else ob.MODE_EXACT, ascii=ascii_fmt)
    prefix_ok = bool((got == want[:P]).all())
    return {"k": k, "rows": int(nS), "colsum_eq_hist_plus_spill": colsum_ok, "spill_total": int(spill),
            "windows_counted": int(col.sum()), "oracle_prefix_reads": P, "oracle_prefix_ok": prefix_ok}


def valid_windows_per_read(torch, flat, nS, L, k, lo, hi):
    """reads [lo, hi) of a uniform-length ASCII batch: number of windows without a non-ACGT byte"""
    v = flat[lo * (L + 1): hi * (L + 1)].view(hi - lo, L + 1)[:, :L]
    bad = ~((v == 65) | (v == 67) | (v == 71) | (v == 84))
    csum = torch.cumsum(torch.nn.functional.pad(bad.to(torch.int32), (1, 0)), dim=1)
    return ((csum[:, k:] - csum[:, :-k]) == 0).sum(dim=1)


def sparse_rows_check(torch, flat, nS, L, k, key_bytes, rc, keys, cnt, lo, hi):
    """rows [lo, hi) of a sparse result over uniform-length reads: counts sum to the valid windows of
    the read, counts >= 1, keys strictly increasing and < 4^k"""
    nwin = L - k + 1
    valid = valid_windows_per_read(torch, flat, nS, L, k, lo, hi)
    rc64 = rc[lo:hi].to(torch.int64)
    K = keys[lo * nwin: hi * nwin].view(hi - lo, nwin)
    Cn = cnt[lo * nwin: hi * nwin].view(hi - lo, nwin)
    live = torch.arange(nwin, device=flat.device).unsqueeze(0) < rc64.unsqueeze(1)
    ok = bool(torch.equal((Cn * live).sum(dim=1, dtype=torch.int64), valid))
    ok = ok and bool((Cn[live] >= 1).all())
    Kl = (K.to(torch.int64) & 0xFFFFFFFF) if key_bytes == 4 else K
    ok = ok and bool(((Kl[:, 1:] > Kl[:, :-1]) | ~live[:, 1:]).all())
    if bool(live.any()):
        ok = ok and int(Kl[live].max()) < 4 ** k and int(Kl[live].min()) >= 0
    return ok, int(valid.sum()), int(rc64.sum())


def sparse_oracle_check(torch, flat, start, length, k, key_bytes, rb, rc, keys, cnt, nrows):
    """rows [0, nrows) against the CPU oracle's sorted distinct k-mers and counts"""
    ob = _oracle()
    end = int(start[nrows - 1].item()) + int(length[nrows - 1].item()) + 1
    h = flat[:end].cpu().numpy().view("int8")
    orp, okeys, ocnt = ob.count_sparse(h, start[:nrows].cpu().numpy(), length[:nrows].cpu().numpy(), k, ascii=True)
    hrb, hrc = rb[: nrows + 1].cpu().numpy(), rc[:nrows].cpu().numpy()
    upto = int(hrb[nrows])
    hk = keys[:upto].cpu().numpy().view("uint32" if key_bytes == 4 else "uint64").astype("uint64")
    hc = cnt[:upto].cpu().numpy().view("uint32")
    for r in range(nrows):
        n = int(hrc[r])
        if n != orp[r + 1] - orp[r]:
            return False
        a = int(hrb[r])
        if not (np.array_equal(hk[a: a + n], okeys[orp[r]: orp[r + 1]]) and np.array_equal(hc[a: a + n], ocnt[orp[r]: orp[r + 1]])):
            return False
    return True


def make_genome_like(torch, nseq, L, seed, device):
    """C4: each sequence = 0.8 L uniform bases + 10 copies of a 0.02 L unit (so counts > 1
    exist), '\n' separated, ASCII"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    unit = L // 50
    body = L - 10 * unit
    flat = torch.empty(nseq * (L + 1) + 16, dtype=torch.uint8, device=device)
    flat[-16:] = 0
    view = flat[: nseq * (L + 1)].view(nseq, L + 1)
    for i in range(nseq):
        view[i, :body] = lut[torch.randint(0, 4, (body,), dtype=torch.uint8, device=device, generator=g).long()]
        u = lut[torch.randint(0, 4, (unit,), dtype=torch.uint8, device=device, generator=g).long()]
        view[i, body:L] = u.repeat(10)
    view[:, L] = 10
    start = torch.arange(nseq, dtype=torch.int64, device=device) * (L + 1)
    length = torch.full((nseq,), L, dtype=torch.int32, device=device)
    return flat, start, length


def split_even(total, world, rank):
    return total // world + (1 if rank < total % world else 0)


def run_sparse_config(torch, dist, cf, dev, rank, world, peak, name, flat, start, length, nS_local, L, k, key_bytes,
                      batch_rows, total_units_all_ranks, warmup, reps, oracle_rows, check_rows):
    """One sparse per-read config over this rank's rows in batches of `batch_rows` (the outputs of a
    batch are reused by the next: rows do not stay).  Strong scaling: the read set is split over
    the ranks, value = all bases / max-over-ranks time."""
    stream = torch.cuda.current_stream().cuda_stream
    nwin = L - k + 1
    cap = batch_rows * nwin
    rb = torch.zeros(batch_rows + 1, dtype=torch.int64, device=dev)
    rc = torch.zeros(batch_rows, dtype=torch.int32, device=dev)
    keys = torch.zeros(cap, dtype=torch.int32 if key_bytes == 4 else torch.int64, device=dev)
    cnt = torch.zeros(cap, dtype=torch.int32, device=dev)
    batches = [(a, min(nS_local, a + batch_rows)) for a in range(0, nS_local, batch_rows)]
    distinct = [0]

    def one_batch(a, b):
        n = b - a
        off = a * (L + 1)
        # start[] of the batch is relative to the batch's first byte (16-byte aligned: (L+1)*a is not
        # in general, so the batch keeps the GLOBAL buffer and a start[] slice instead)
        cf.count_sparse_device(flat.data_ptr(), start[a:b].data_ptr(), length[a:b].data_ptr(), flat.numel() - 16, n, k,
                               rb.data_ptr(), rc.data_ptr(), keys.data_ptr(), cnt.data_ptr(), cap, key_bytes=key_bytes,
                               fmt=cf.FMT_ASCII, stream=stream)
        return off

    def one_pass():
        for a, b in batches:
            one_batch(a, b)

    launches0 = cf.launch_count()
    ms = timed(torch, one_pass, warmup, reps)
    launches = (cf.launch_count() - launches0) // (warmup + reps)
    # checks on the first batch (left in the buffers by a fresh call)
    a, b = batches[0]
    one_batch(a, b)
    torch.cuda.synchronize()
    n = b - a
    ok_rows, nvalid, ndistinct = True, 0, 0
    step = max(1, min(n, 1_000_000 if L <= 1000 else 8))
    for lo in range(0, min(n, check_rows), step):
        hi = min(n, lo + step)
        okr, v, d = sparse_rows_check(torch, flat, nS_local, L, k, key_bytes, rc, keys, cnt, lo, hi)
        ok_rows = ok_rows and okr
        nvalid += v
        ndistinct += d
    ok_oracle = sparse_oracle_check(torch, flat, start, length, k, key_bytes, rb, rc, keys, cnt, min(oracle_rows, n))
    # distinct pairs of the whole pass (for the algorithmic bytes): count per batch
    total_distinct = 0
    for a, b in batches:
        one_batch(a, b)
        total_distinct += int(rc[: b - a].sum(dtype=torch.int64))
    t = torch.tensor([ms, float(total_distinct), 1.0 if (ok_rows and ok_oracle) else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        tn = t.clone()
        dist.all_reduce(tn, op=dist.ReduceOp.MIN)
        ms, total_distinct, all_ok = float(tm[0]), float(ts[1]), float(tn[2]) > 0.5
    else:
        all_ok = ok_rows and ok_oracle
    del rb, rc, keys, cnt
    bases = total_units_all_ranks * L
    alg = total_units_all_ranks * (L + 8) + total_distinct * (key_bytes + 4)
    return {"config": name, "k": k, "key_bytes": key_bytes, "rows_all_ranks": int(total_units_all_ranks), "row_len": L,
            "rows_this_rank": int(nS_local), "batches_per_pass": len(batches), "ms_per_pass": round(ms, 3),
            "gbases_s": round(bases / ms / 1e6, 2), "alg_bytes": int(alg), "alg_gb_s": round(alg / ms / 1e6, 1),
            "frac_of_peak_per_gpu": round(alg / ms / 1e6 / peak / world, 4), "scaling": "strong",
            "launches_per_pass": int(launches),
            "check": {"rows_checked": int(min(n, check_rows)), "counts_sum_to_valid_windows_sorted_keys": bool(ok_rows),
                      "oracle_rows": int(min(oracle_rows, n)), "oracle_ok": bool(ok_oracle), "all_ranks_ok": bool(all_ok),
                      "valid_windows_checked": int(nvalid), "distinct_pairs_checked": int(ndistinct)}}


def bind_to_gpu_numa_node(torch, local):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned buffer is
    allocated (first touch puts the pages there): the e2e leg moves gigabytes of rows per step over
    PCIe into host memory, and a rank whose buffers sit on the other socket pays the inter-socket
    link as well.  Returns the node, or None when the topology is not visible (containers)."""
    if os.environ.get("CFRK_BENCH_NO_NUMA"):
        return None
    try:
        p = torch.cuda.get_device_properties(local)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cfrk_b200 as cf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or cf.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # the host threads of the operator (e2e leg) share the box's cores with the other ranks
        cf.lib().cfrk_set_host_threads(max(1, min(16, len(os.sched_getaffinity(0)) // world)))

    ks = [int(x) for x in args.k.split(",")]
    nS, L = args.reads, args.read_len
    packed = args.fmt == "packed"
    fmt = cf.FMT_CODES if args.fmt == "codes" else cf.FMT_ASCII
    mode = cf.MODE_COMPAT if args.mode == "compat" else cf.MODE_EXACT
    flat, start, length = make_reads_device(torch, nS, L, 42 + rank, args.n_frac, "codes" if args.fmt == "codes" else "ascii", dev)
    nN = nS * (L + 1)
    if packed:
        nblk = (nN + 15) // 16 + 1
        p_codes = torch.zeros(nblk, dtype=torch.int32, device=dev)
        p_valid = torch.zeros(nblk, dtype=torch.int16, device=dev)
    ring_bytes = int(args.ring_gib * (1 << 30))
    need = max(min(ring_bytes, nS * 4 ** k * 4) for k in ks)
    ring = torch.empty(need // 4, dtype=torch.int32, device=dev)
    ring_gib = ring.numel() * 4 / 2 ** 30
    stream = torch.cuda.current_stream().cuda_stream

    def plan(k):
        row = 4 ** k * 4
        rpt = cf.dense_reads_per_tile(k)
        per = max(rpt, (ring.numel() * 4 // row) // rpt * rpt)
        return [(a, min(nS, a + per)) for a in range(0, nS, per)]

    want_cfg = set() if args.no_configs else set(x for x in args.configs.split(",") if x)
    ks_small = [k for k in (1, 2, 3) if k not in ks] if "small_k" in want_cfg else []
    plans = {k: plan(k) for k in ks + ks_small}

    def encode():
        if packed:
            cf.encode_2bit_device(flat.data_ptr(), nN, p_codes.data_ptr(), p_valid.data_ptr(), fmt=cf.FMT_ASCII, stream=stream)

    def launch_rows(k, a, b, src=None):
        fl, st_, ln = src if src is not None else (flat, start, length)
        # encode-once mode keeps both layouts resident; k = 8 reads the ASCII bases: its kernel is a write
        # stream with sparse patches, and the two narrow loads per block of the packed layout cost it 8 %
        if packed and src is None and k < 8:
            cf.count_dense_packed_device(p_codes.data_ptr(), p_valid.data_ptr(), st_.data_ptr(), ln.data_ptr(),
                                         nN, nS, k, ring.data_ptr(), mode=mode, read_begin=a, read_end=b, stream=stream)
        else:
            cf.count_dense_device(fl.data_ptr(), st_.data_ptr(), ln.data_ptr(), nN, nS, k,
                                  ring.data_ptr(), mode=mode, fmt=fmt, read_begin=a, read_end=b, stream=stream)

    def sweep_k(k, src=None):
        for a, b in plans[k]:
            launch_rows(k, a, b, src)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        encode()
        for k in ks:
            sweep_k(k)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ks]
          for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = cf.launch_count()
    barrier()
    e0.record()
    for s in range(args.steps):
        encode()
        for i, k in enumerate(ks):
            ev[s][i][0].record()
            sweep_k(k)
            ev[s][i][1].record()
    e1.record()
    barrier()
    launches = cf.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    per_k_ms = {k: float(np.mean([ev[s][i][0].elapsed_time(ev[s][i][1]) for s in range(args.steps)]))
                for i, k in enumerate(ks)}
    # context for write-dominated kernels: what a plain fill of the same ring achieves on this GPU
    # (MEASURED_PEAKS.json's hbm_gbs is a COPY, read + write; a write-only stream can exceed it)
    wbest = 1e9
    for _ in range(3):
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(); ring.fill_(0); w1.record(); torch.cuda.synchronize()
        wbest = min(wbest, w0.elapsed_time(w1))
    write_only_gbs = ring.numel() * 4 / wbest / 1e6
    peak, peak_src = measured_peak()
    def per_k_entry(k, ms):
        bytes_k = nS * alg_bytes_per_read(L, k)
        return {"k": k, "ms": round(ms, 3), "gbases_s": round(nS * L / ms / 1e6, 2),
                "alg_gb_s": round(bytes_k / ms / 1e6, 1), "frac_of_peak": round(bytes_k / ms / 1e6 / peak, 4),
                "launches": len(plans[k])}

    per_k = [per_k_entry(k, per_k_ms[k]) for k in ks]
    # k = 1..3 (the reference's own test config is k = 2, test/test.sh:13): same reads, own events,
    # not part of `value` (BASELINE configs[1] names k = 4..8)
    for k in ks_small:
        per_k.append(dict(per_k_entry(k, timed(torch, lambda: sweep_k(k), 3, 5)), in_value=False))
    per_k.sort(key=lambda d: d["k"])
    min_frac = min(d["frac_of_peak"] for d in per_k)
    dom = max(ks, key=lambda k: per_k_ms[k])
    dom_launches = len(plans[dom])
    dom_bytes_per_launch = nS * alg_bytes_per_read(L, dom) / dom_launches
    dom_s_per_launch = per_k_ms[dom] / 1e3 / dom_launches
    achieved = dom_bytes_per_launch / dom_s_per_launch / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "kernel": {8: "dense_bigrow_kernel<8,ascii,65536>", 7: "dense_bigrow_kernel<7,FMT,32768>", 6: "dense_warp_kernel<6,FMT,1,6>",
                           5: "dense_warp_kernel<5,FMT,3,3>"}.get(dom, f"dense_lane_kernel<{dom},FMT>").replace("FMT", args.fmt), "share_of_step": round(per_k_ms[dom] / sum(per_k_ms.values()), 3),
                "bytes_per_launch": int(dom_bytes_per_launch), "us_per_launch": round(dom_s_per_launch * 1e6, 1),
                "write_only_fill_gbs": round(write_only_gbs, 1),
                "note": "peak = measured COPY bandwidth (read+write); the dominant kernel only writes, and a plain "
                        "fill_ of the same ring reaches write_only_fill_gbs on this GPU, so frac can exceed 1"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tr = json.load(f)
            roofline["traffic"] = tr.get(str(dom), {}).get("dram_bytes_per_launch")
            roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    steps = args.steps
    bases_per_step = world * nS * L * len(ks)
    value = bases_per_step * steps / (total_ms / 1e3) / 1e9

    # ---- e2e: host buffers through the reference-facing operator, rank-local chunk ----------
    e2e = None
    if not args.no_e2e:
        cn = args.e2e_reads
        hb, hs, hl = make_reads_host(cn, L, 1000 + rank, "codes")
        hb_t = torch.from_numpy(hb).pin_memory(); hs_t = torch.from_numpy(hs).pin_memory()
        hl_t = torch.from_numpy(hl).pin_memory()
        houts = {k: torch.empty((cn, 4 ** k), dtype=torch.int32).pin_memory() for k in ks}

        def e2e_sweep():
            for k in ks:
                cf.count_dense_host(hb_t.numpy(), hs_t.numpy(), hl_t.numpy(), k, mode, cf.FMT_CODES, local,
                                    out=houts[k].numpy())
        for _ in range(max(1, args.warmup)):
            e2e_sweep()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_sweep()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        host_v, host_ms = world * cn * L * len(ks) * steps / dt / 1e9, dt / steps * 1e3

        # the same sweep through the EXPORTED REFERENCE SYMBOL kmer_main(struct read*, nN, nS, k, device)
        # (src/kmer.cuh:6): the callee allocates rd->Freq pinned, exactly what the reference arm times;
        # the rows go back to the library's pinned arena with cfrk_free_host (the reference arm frees
        # with cudaFreeHost)
        km = getattr(cf.lib(), "_Z9kmer_mainP4readllit")
        km.argtypes = [C.POINTER(_RefRead), C.c_long, C.c_long, C.c_int, C.c_ushort]
        km.restype = None

        def km_sweep():
            for k in ks:
                rd = _RefRead(hb_t.data_ptr(), hl_t.data_ptr(), hs_t.data_ptr(), None, None)
                km(C.byref(rd), hb.nbytes, cn, k, local)
                if not rd.Freq:
                    raise SystemExit("bench.py: kmer_main returned no rows")
                cf.lib().cfrk_free_host(rd.Freq)
        for _ in range(max(1, args.warmup)):
            km_sweep()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            km_sweep()
        torch.cuda.synchronize()
        dtk = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dtk], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtk = float(t.item())
        e2e = {"value": round(world * cn * L * len(ks) * steps / dtk / 1e9, 4), "unit": UNIT,
               "h2d_bytes_per_step": int(len(ks) * (hb.nbytes + hs.nbytes + hl.nbytes)),
               "d2h_bytes_per_step": int(sum(cn * 4 ** k * 4 for k in ks)),
               "sample": f"{cn} reads x {L} bp per k (one reference chunk) through the exported reference symbol kmer_main "
                         f"(struct read in pinned host memory, rd->Freq allocated by the callee, returned with "
                         f"cfrk_free_host), codes layout", "ms_per_step": round(dtk / steps * 1e3, 2),
               "dense_host": {"value": round(host_v, 4), "ms_per_step": round(host_ms, 2),
                              "what": "cfrk_count_dense_host with a caller-owned pinned output buffer"}}

    # ---- the same reference chunk resident in HBM: what sits next to the reference arm's kernels_only
    chunk_resident = None
    if not args.no_e2e and rank == 0:
        db, ds, dl = hb_t.to(dev), hs_t.to(dev), hl_t.to(dev)
        rows = max(cn * 4 ** k for k in ks)
        dout = ring if ring.numel() >= rows else torch.empty(rows, dtype=torch.int32, device=dev)
        per, tot = [], 0.0
        for k in ks:
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for r in range(6):
                if r == 1:
                    c0.record()
                cf.count_dense_device(db.data_ptr(), ds.data_ptr(), dl.data_ptr(), hb.nbytes, cn, k, dout.data_ptr(),
                                      mode=mode, fmt=cf.FMT_CODES, stream=stream)
            c1.record(); torch.cuda.synchronize()
            ms = c0.elapsed_time(c1) / 5
            per.append({"k": k, "ms": round(ms, 4), "gbases_s": round(cn * L / ms / 1e6, 2)})
            tot += ms
        chunk_resident = {"value": round(cn * L * len(ks) / tot / 1e6, 2), "unit": UNIT, "ms_per_sweep": round(tot, 3), "per_k": per,
                          "what": f"{cn} reads x {L} bp (one reference chunk, codes layout) resident in HBM, rows left in HBM: "
                                  f"the counterpart of the reference arm's kernels_only"}

    # ---- on-GPU checks of the dense rows: every k of this run, all nS rows of this rank -------------
    checks = None
    if not args.no_checks:
        checks = {"dense": [], "what": "colsum over ALL rows == cfrk_global_hist_device(visited windows) + spill total "
                                       "from the lengths; first rows bit-exact vs the CPU oracle (oracle_count_fast_mt)"}
        for k in sorted(ks + ks_small):
            checks["dense"].append(dense_check(torch, cf, flat, start, length, nN, nS, k, mode, args.fmt != "codes", ring,
                                               plans[k], launch_rows, stream, 60000))
        ok = all(c["colsum_eq_hist_plus_spill"] and c["oracle_prefix_ok"] for c in checks["dense"])
        if world > 1:
            t = torch.tensor([1.0 if ok else 0.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = float(t.item()) > 0.5
        checks["all_ok_all_ranks"] = ok

    # ---- the other north-star configs ---------------------------------------------------------------
    configs = {}
    if "c2b" in want_cfg:
        # C2 variant B: 0.1 % of the bases are N (seed 43): in-read invalid windows, data-dependent spill
        fb, sb, lb = make_reads_device(torch, nS, L, 43 + rank, 0.001, "codes" if args.fmt == "codes" else "ascii", dev)
        srcb = (fb, sb, lb)
        entry = {"what": f"{nS} reads x {L} bp, 0.1 % N bases (seed 43+rank), same kernels, same ring", "per_k": []}
        tot = 0.0
        for k in ks:
            ms = timed(torch, lambda: sweep_k(k, srcb), 1, 2)
            tot += ms
            entry["per_k"].append(per_k_entry(k, ms))
        entry["gbases_s"] = round(nS * L * len(ks) / tot / 1e6, 3)
        if not args.no_checks:
            entry["check"] = [dense_check(torch, cf, fb, sb, lb, nN, nS, k, mode, args.fmt != "codes", ring, plans[k],
                                          lambda kk, a, b: launch_rows(kk, a, b, srcb), stream, 60000) for k in ks]
        configs["C2_variant_B"] = entry
        del fb, sb, lb, srcb
        torch.cuda.empty_cache()
    del ring
    if packed:
        del p_codes, p_valid
    torch.cuda.empty_cache()
    if "c3" in want_cfg:
        n3 = split_even(args.c3_reads, world, rank)
        f3, s3, l3 = make_reads_device(torch, n3, L, 44 + rank, 0.001, "ascii", dev)
        configs["C3_sparse_k12"] = run_sparse_config(torch, dist, cf, dev, rank, world, peak, "C3: 100 M x 150 bp, 0.1 % N, sparse "
                                                     "per-read rows, read range split over the ranks", f3, s3, l3, n3, L, 12, 4,
                                                     min(n3, 10_000_000), args.c3_reads, 1, 3, 20000, 2_000_000)
        del f3, s3, l3
        torch.cuda.empty_cache()
    if "c4" in want_cfg:
        n4 = split_even(args.c4_seqs, world, rank)
        if n4 > 0:
            f4, s4, l4 = make_genome_like(torch, n4, args.c4_len, 45 + rank, dev)
            for k4 in (16, 21, 31):
                configs[f"C4_sparse_k{k4}"] = run_sparse_config(
                    torch, dist, cf, dev, rank, world, peak, "C4: 1000 x 5 Mbp (80 % uniform + 10 copies of a 2 % unit), sparse "
                    "sort-compact rows, sequences split over the ranks", f4, s4, l4, n4, args.c4_len, k4, 4 if k4 <= 16 else 8,
                    min(n4, 40), args.c4_seqs, 1, 2, 1, 40)
            del f4, s4, l4
            torch.cuda.empty_cache()
    if "c5" in want_cfg:
        k5 = 12
        n5 = split_even(args.c5_reads, world, rank)
        f5, s5, l5 = make_reads_device(torch, n5, L, 46 + rank, 0.0, "ascii", dev)
        # N > 1: the table lives in memory the peers can address and is summed in place by cfrk_hist_allreduce_device
        # (one kernel per GPU over NVLink peer memory, hist_reduce.cu); NCCL only if that cannot be set up
        reducer, reduce_how = None, "none (1 GPU)"
        if world > 1:
            try:
                from cfrk_b200.sharding import HistReducer
                reducer = HistReducer(4 ** k5, dev)
                reduce_how = f"cfrk_hist_allreduce_device ({reducer.mode})"
            except Exception as e:  # noqa: BLE001
                reducer, reduce_how = None, f"NCCL all-reduce (peer memory unavailable: {type(e).__name__})"
            flag = torch.tensor([1 if reducer is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)     # all ranks or none
            if int(flag) == 0 and reducer is not None:
                reducer, reduce_how = None, "NCCL all-reduce (peer memory unavailable on a peer)"
        hist = reducer.table if reducer is not None else torch.zeros(4 ** k5, dtype=torch.int32, device=dev)
        W5, K5 = 3, 10
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K5)]

        def c5_step(ev=None):
            hist.zero_()
            if ev:
                ev[0].record()
            cf.global_hist_device(f5.data_ptr(), s5.data_ptr(), l5.data_ptr(), n5 * (L + 1), n5, k5, hist.data_ptr(),
                                  fmt=cf.FMT_ASCII, stream=stream)
            if ev:
                ev[1].record()
            if reducer is not None:
                reducer.allreduce(stream)                     # the path's one exchange step
            elif world > 1:
                dist.all_reduce(hist, op=dist.ReduceOp.SUM)
            if ev:
                ev[2].record()
        for _ in range(W5):
            c5_step()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(K5):
            c5_step(evs[i])
        c1.record()
        barrier()
        t = torch.tensor([c0.elapsed_time(c1) / K5, sum(e[0].elapsed_time(e[1]) for e in evs) / K5,
                          sum(e[1].elapsed_time(e[2]) for e in evs) / K5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms5, cnt5, red5 = (float(x) for x in t)
        windows = args.c5_reads * (L - k5 + 1)
        nccl5 = None
        if world > 1:        # the library collective on the same table, for the record (outside the timed steps)
            tmp = hist.clone()
            for _ in range(3):
                dist.all_reduce(tmp)
            barrier()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0.record()
            for _ in range(K5):
                dist.all_reduce(tmp)
            n1.record()
            torch.cuda.synchronize()
            tn5 = torch.tensor([n0.elapsed_time(n1) / K5], dtype=torch.float64, device=dev)
            dist.all_reduce(tn5, op=dist.ReduceOp.MAX)
            nccl5 = round(float(tn5), 3)
            del tmp
        status5 = reducer.status() if reducer is not None else 0
        configs["C5_global_hist_k12"] = {
            "config": "C5: 1 Gbase (6 666 667 x 150 bp) whole-dataset histogram, k = 12, reads split over the ranks, "
                      "64 MiB table summed over the ranks in place", "reduce": reduce_how, "nccl_allreduce_ms": nccl5, "k": k5, "reads_all_ranks": args.c5_reads, "scaling": "strong",
            "ms_per_step": round(ms5, 3), "count_ms": round(cnt5, 3), "allreduce_ms": round(red5, 3),
            "gbases_s": round(args.c5_reads * L / ms5 / 1e6, 1), "red_global_per_s_per_gpu_G": round(windows / world / cnt5 / 1e6, 1),
            "alg_bytes": int(args.c5_reads * (L + 8) + 4 ** k5 * 4),
            "check": {"sum_eq_windows": int(hist.sum(dtype=torch.int64)) == windows, "all_peers_met": status5 == 0}}
        del f5, s5, l5, hist, reducer
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if args.no_cpu else cpu_baseline(args, ks, L, args.cpu_seconds)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": round(total_ms / steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"C2: {nS} reads x {L} bp uniform ACGT per GPU (seed 42+rank), dense per-read int32 "
                               f"counts, sweep k={ks}, {args.mode} semantics, {'ASCII bases resident in HBM, encoded ONCE per step to packed 2-bit words + validity (inside the timed region) and counted from those for k < 8' if packed else args.fmt + ' bases resident in HBM'}, rows to a "
                               f"{ring_gib:.1f} GiB HBM ring",
                   "l2_policy": "inputs (1.5 GB) and outputs (>= 10 GB per k) larger than L2 (126 MB); no flush needed",
                   "reads_per_gpu": nS, "read_len": L, "k": ks, "parallelism": f"read-range shards x{world}",
                   "host_numa_node_rank0": numa_node},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "per_k": per_k, "min_frac": min_frac, "chunk_resident": chunk_resident,
        "checks": checks, "configs": configs,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
class _RefRead(C.Structure):   # reference ABI, src/tipos.h:23-30
    _fields_ = [("data", C.c_void_p), ("length", C.c_void_p), ("start", C.c_void_p),
                ("Freq", C.c_void_p), ("next", C.c_void_p)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ks = [int(x) for x in args.k.split(",")]
    L, cn = args.read_len, args.e2e_reads
    steps = args.steps
    so = os.path.join(ROOT, "oracle", "_ref", "libcfrk_ref_gpu.so")
    cpu = None if args.no_cpu else cpu_baseline(args, ks, L, args.cpu_seconds)
    gpu_ok = False
    try:
        import torch
        gpu_ok = torch.cuda.is_available() and os.path.exists(so)
    except Exception:
        gpu_ok = False
    cfg = {"workload": f"C2 sample: {cn} reads x {L} bp uniform ACGT per k (one reference chunk), sweep k={ks}, "
                       f"host buffers in, host rows out", "reads": cn, "read_len": L, "k": ks}
    if gpu_ok:
        import torch
        lib = C.CDLL(so)
        km = getattr(lib, "_Z9kmer_mainP4readllit")
        km.argtypes = [C.POINTER(_RefRead), C.c_long, C.c_long, C.c_int, C.c_ushort]
        km.restype = None
        lib.ref_free_host.argtypes = [C.c_void_p]
        torch.cuda.set_device(0)
        bind_to_gpu_numa_node(torch, 0)
        hb, hs, hl = make_reads_host(cn, L, 1000, "codes")
        hb_t = torch.from_numpy(hb).pin_memory(); hs_t = torch.from_numpy(hs).pin_memory()
        hl_t = torch.from_numpy(hl).pin_memory()
        nN = hb.nbytes

        def sweep():
            for k in ks:
                rd = _RefRead(hb_t.data_ptr(), hl_t.data_ptr(), hs_t.data_ptr(), None, None)
                km(C.byref(rd), nN, cn, k, 0)
                if rd.Freq:
                    lib.ref_free_host(rd.Freq)   # the reference never frees it (src/kmer_main.cu:115)
        # the reference reports CUDA errors with printf on stdout (src/kmer_main.cu:59-63): keep
        # stdout for the one JSON line by pointing fd 1 at stderr while its code runs
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            for _ in range(max(1, args.warmup)):
                sweep()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                sweep()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        v = cn * L * len(ks) * steps / dt / 1e9
        # SURVEY 8(d) baseline 1b: the reference's four kernel launches alone, inputs resident in HBM
        # (oracle/ref_helper.cu restates the launch shapes of src/kmer_main.cu:66-111)
        kernels_only = None
        if hasattr(lib, "ref_kernels_time"):
            lib.ref_kernels_time.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int,
                                             C.POINTER(C.c_float)]
            db, ds, dl = hb_t.cuda(), hs_t.cuda(), hl_t.cuda()
            per_k, tot_ms = [], 0.0
            for k in ks:
                ms = C.c_float(0)
                rc = lib.ref_kernels_time(db.data_ptr(), ds.data_ptr(), dl.data_ptr(), nN, cn, k, 5, C.byref(ms))
                if rc != 0:
                    per_k = None
                    break
                per_k.append({"k": k, "ms": round(ms.value, 4), "gbases_s": round(cn * L / ms.value / 1e6, 3)})
                tot_ms += ms.value
            if per_k:
                kernels_only = {"value": round(cn * L * len(ks) / tot_ms / 1e6, 3), "unit": UNIT, "ms_per_sweep": round(tot_ms, 3),
                                "per_k": per_k,
                                "what": "SetMatrix x2 + ComputeIndex + ComputeFreqNew (the reference's objects, its launch shapes), "
                                        f"{cn} reads x {L} bp resident in HBM, CUDA events, no allocation or copy in the timed region"}
        line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": 1,
                "steps": steps, "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/i32",
                "data": "synthetic", "config": cfg,
                "reference_kind": "the reference's own kmer_main (unmodified kmer_main.cu + kmer_kernel.cu, nvcc "
                                  "sm_100) on one B200 -- the reference has no CPU path",
                "cpu_baseline": cpu, "kernels_only": kernels_only,
                "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        v = cpu["value"] if cpu else None
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": steps,
                "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "i32", "data": "synthetic", "config": cfg,
                "reference_kind": "oracle port on host cores (oracle/_ref/libcfrk_ref_gpu.so or GPU unavailable)",
                "cpu_baseline": cpu,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class QuietStdout:
    """Everything but the final JSON line goes to stderr: NCCL prints its version banner on stdout,
    the reference prints its errors there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


if __name__ == "__main__":
    a = parse_args()
    out_fd = os.dup(1)
    with QuietStdout():
        real_print = print

        def emit(*args, **kw):          # the one JSON line goes to the real stdout
            os.write(out_fd, (" ".join(str(x) for x in args) + "\n").encode())
        import builtins
        builtins.print = emit
        try:
            if a.impl == "reference":
                run_reference(a)
            else:
                run_ours(a)
        finally:
            builtins.print = real_print
