#!/usr/bin/env python
"""End-to-end time of the `cfrk` command on a synthetic FASTA file (reported separately
from the kernel-level roofline).  python tools/bench_cli.py [--reads N] [--k K] [--nt T]"""
import argparse, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=2_000_000)
ap.add_argument("--read-len", type=int, default=150)
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--nt", type=int, default=os.cpu_count() or 8)
ap.add_argument("--dir", default="/dev/shm")
ap.add_argument("--devices", default="", help="e.g. 0,1 or all: several GPUs (cfrk --devices=...)")
ap.add_argument("--runs", default="all_rows_dense,all_rows_sparse,tail_only")
ap.add_argument("--envs", default="", help='A/B: "name:VAR=x,VAR2=y;name2:..." -- every run once per environment, one JSON line each')
ap.add_argument("--md5", action="store_true", help="md5 of every output (A/B runs must agree)")
a = ap.parse_args()
d = tempfile.mkdtemp(dir=a.dir if os.path.isdir(a.dir) else None)
fa, out = os.path.join(d, "in.fa"), os.path.join(d, "out.cfrk")
rng = np.random.default_rng(1)
with open(fa, "wb") as f:
    for s in range(0, a.reads, 100_000):
        n = min(100_000, a.reads - s)
        blk = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, a.read_len))]
        hdr = np.frombuffer("".join(f">{i:09d}\n" for i in range(s, s + n)).encode(), dtype=np.uint8).reshape(n, 11)
        f.write(np.concatenate([hdr, blk, np.full((n, 1), 10, np.uint8)], axis=1).tobytes())
dev = [f"--devices={a.devices}"] if a.devices else []
variants = [("", {})]
if a.envs:
    variants = []
    for item in a.envs.split(";"):
        name, _, kv = item.partition(":")
        variants.append((name, dict(x.split("=", 1) for x in kv.split(",") if x)))
for vname, venv in variants:
    res = {}
    for label, extra in (("all_rows_dense", ["--all-rows"]), ("all_rows_sparse", ["--all-rows", "--sparse"]), ("tail_only", [])):
        if label not in a.runs.split(","):
            continue
        extra = extra + dev
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "bin", "cfrk"), fa, out, str(a.k), str(a.nt), "8192", *extra],
                           capture_output=True, env=dict(os.environ, CFRK_TRACE="1", **venv))
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr.decode()[-500:]
        res[label] = {"seconds": round(dt, 3), "out_bytes": os.path.getsize(out),
                      "gbases_s": round(a.reads * a.read_len / dt / 1e9, 4),
                      "out_gb_s": round(os.path.getsize(out) / dt / 1e9, 3)}
        if os.environ.get("CFRK_BENCH_CLI_TRACE"):
            sys.stderr.write(f"== {vname} {label} k={a.k}\n" + r.stderr.decode())
        last = [l for l in r.stderr.decode().splitlines() if "trace" in l][-1:]
        res[label]["pipeline_ms"] = float(last[0].split()[2]) if last else None
        if a.md5:
            import hashlib
            h = hashlib.md5()
            with open(out, "rb") as f:
                for blk in iter(lambda: f.read(1 << 24), b""):
                    h.update(blk)
            res[label]["md5"] = h.hexdigest()
        os.remove(out)          # the next run must not pay for truncating this one's pages
        time.sleep(0.5)
    print(json.dumps({"variant": vname, "env": venv, "reads": a.reads, "read_len": a.read_len, "k": a.k, "nt": a.nt, "devices": a.devices or "0",
                      "fasta_bytes": os.path.getsize(fa), **res}), flush=True)
os.remove(fa)
os.rmdir(d)
