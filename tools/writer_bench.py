"""Host-only timing of the .cfrk text writer (cfrk_write_rows): rows of 150-bp reads, k = 4.
usage: python tools/writer_bench.py [rows] [nt] [out_dir]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import cfrk_b200 as cf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 414648
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out_dir = sys.argv[3] if len(sys.argv) > 3 else "/dev/shm"
k, bins = 4, 256
rng = np.random.default_rng(1)
rows = np.zeros((n, bins), dtype=np.int32)
idx = rng.integers(0, bins, size=(n, 147))
np.add.at(rows, (np.arange(n)[:, None], idx), 1)
L = cf.lib()
path = os.path.join(out_dir, "cfrk_writer_bench.out").encode()
for flags, name in ((0, "dense"), (cf.RUN_SPARSE, "sparse")):
    for rep in range(3):
        if os.path.exists(path): os.unlink(path)
        t0 = time.perf_counter()
        rc = L.cfrk_write_rows(path, rows.ctypes.data, n, k, nt, flags)
        dt = time.perf_counter() - t0
        sz = os.path.getsize(path)
        print(f"{name} nt={nt} rc={rc} {dt*1e3:.1f} ms, {sz/1e6:.0f} MB, {sz/dt/1e9:.2f} GB/s", flush=True)
os.unlink(path)
