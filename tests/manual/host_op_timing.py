#!/usr/bin/env python
"""per-call wall time of cfrk_count_dense_host with different kinds of output buffers (diagnostic)"""
import ctypes as C, os, sys, time
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cfrk_b200 as cf
from bench import make_reads_host
cn, L = 8192, 150
hb, hs, hl = make_reads_host(cn, L, 1000, "codes")
hb_t, hs_t, hl_t = (torch.from_numpy(x).pin_memory() for x in (hb, hs, hl))
ks = [4, 5, 6, 7, 8]
def bufs(kind):
    out = {}
    for k in ks:
        if kind == "torch_pinned":
            out[k] = torch.empty((cn, 4 ** k), dtype=torch.int32).pin_memory().numpy()
        elif kind == "pageable":
            out[k] = np.empty((cn, 4 ** k), dtype=np.int32)
        else:   # cudaMallocHost through torch's runtime binding
            n = cn * 4 ** k * 4
            p = C.c_void_p()
            rt = C.CDLL("libcudart.so.12")
            assert rt.cudaMallocHost(C.byref(p), C.c_size_t(n)) == 0
            out[k] = np.ctypeslib.as_array((C.c_int32 * (n // 4)).from_address(p.value)).reshape(cn, 4 ** k)
    return out
for kind in ("cuda_malloc_host", "torch_pinned", "pageable", "torch_pinned"):
    o = bufs(kind)
    for rep in range(6):
        t = []
        for k in ks:
            t0 = time.perf_counter()
            cf.count_dense_host(hb_t.numpy(), hs_t.numpy(), hl_t.numpy(), k, cf.MODE_COMPAT, cf.FMT_CODES, 0, out=o[k])
            t.append((time.perf_counter() - t0) * 1e3)
        print(kind, rep, " ".join(f"{x:7.2f}" for x in t), f" sum {sum(t):7.2f} ms", flush=True)
    del o
