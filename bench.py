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
    ap.add_argument("--fmt", default="ascii", choices=["ascii", "codes", "packed"],
                    help="packed: every step encodes the ASCII bases once (2-bit + validity) and counts every k from that")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    stream = torch.cuda.current_stream().cuda_stream

    def plan(k):
        row = 4 ** k * 4
        rpt = cf.dense_reads_per_tile(k)
        per = max(rpt, (ring.numel() * 4 // row) // rpt * rpt)
        return [(a, min(nS, a + per)) for a in range(0, nS, per)]

    plans = {k: plan(k) for k in ks}

    def encode():
        if packed:
            cf.encode_2bit_device(flat.data_ptr(), nN, p_codes.data_ptr(), p_valid.data_ptr(), fmt=cf.FMT_ASCII, stream=stream)

    def sweep_k(k):
        for a, b in plans[k]:
            if packed:
                cf.count_dense_packed_device(p_codes.data_ptr(), p_valid.data_ptr(), start.data_ptr(), length.data_ptr(),
                                             nN, nS, k, ring.data_ptr(), mode=mode, read_begin=a, read_end=b, stream=stream)
            else:
                cf.count_dense_device(flat.data_ptr(), start.data_ptr(), length.data_ptr(), nN, nS, k,
                                      ring.data_ptr(), mode=mode, fmt=fmt, read_begin=a, read_end=b, stream=stream)

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
    per_k = []
    for k in ks:
        ms = per_k_ms[k]
        bytes_k = nS * alg_bytes_per_read(L, k)
        per_k.append({"k": k, "ms": round(ms, 3), "gbases_s": round(nS * L / ms / 1e6, 2),
                      "alg_gb_s": round(bytes_k / ms / 1e6, 1), "frac_of_peak": round(bytes_k / ms / 1e6 / peak, 4),
                      "launches": len(plans[k])})
    dom = max(ks, key=lambda k: per_k_ms[k])
    dom_launches = len(plans[dom])
    dom_bytes_per_launch = nS * alg_bytes_per_read(L, dom) / dom_launches
    dom_s_per_launch = per_k_ms[dom] / 1e3 / dom_launches
    achieved = dom_bytes_per_launch / dom_s_per_launch / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "kernel": f"dense_count_kernel<K={dom},{args.fmt}>", "share_of_step": round(per_k_ms[dom] / sum(per_k_ms.values()), 3),
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
        e2e = {"value": round(world * cn * L * len(ks) * steps / dt / 1e9, 4), "unit": UNIT,
               "h2d_bytes_per_step": int(len(ks) * (hb.nbytes + hs.nbytes + hl.nbytes)),
               "d2h_bytes_per_step": int(sum(cn * 4 ** k * 4 for k in ks)),
               "sample": f"{cn} reads x {L} bp per k (one reference chunk), cfrk_count_dense_host, pinned host "
                         f"buffers, codes layout", "ms_per_step": round(dt / steps * 1e3, 2)}

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
                               f"counts, sweep k={ks}, {args.mode} semantics, {args.fmt} bases resident in HBM, rows to a "
                               f"{ring.numel() * 4 / 2**30:.1f} GiB HBM ring",
                   "l2_policy": "inputs (1.5 GB) and outputs (>= 10 GB per k) larger than L2 (126 MB); no flush needed",
                   "reads_per_gpu": nS, "read_len": L, "k": ks, "parallelism": f"read-range shards x{world}",
                   "host_numa_node_rank0": numa_node},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "per_k": per_k, "chunk_resident": chunk_resident,
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
