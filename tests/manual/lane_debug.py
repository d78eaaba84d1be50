#!/usr/bin/env python
"""debug helper: every edge fixture x k x mode in its own process (a device fault kills the context)"""
import os, subprocess, sys, json
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, torch
    import cfrk_b200 as cf, fixtures as fx, oracle_binding as ob
    name, k, mode = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    text = dict((n, t) for n, t, _ in fx.EDGE_SET)[name]
    data, start, length = ob.parse_fasta(text=text)
    want = ob.count_dense_fast(data, start, length, k, mode)
    got = cf.count_dense_host(data, start, length, k, mode, cf.FMT_CODES)
    bad = np.argwhere(got != want)
    print(json.dumps({"fixture": name, "k": k, "mode": mode, "ok": bool(len(bad) == 0), "nbad": int(len(bad)),
                      "first_bad": bad[:4].tolist(), "lens_of_bad_rows": [int(length[b[0]]) for b in bad[:4]],
                      "got": [int(got[tuple(b)]) for b in bad[:4]], "want": [int(want[tuple(b)]) for b in bad[:4]]}))
    sys.exit(0)
import fixtures as fx
for k in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,3").split(",")]:
    for name, _, _ in fx.EDGE_SET:
        for mode in (0, 1):
            r = subprocess.run([sys.executable, __file__, "one", name, str(k), str(mode)], capture_output=True, text=True, timeout=120)
            out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAULT rc=%d %s" % (r.returncode, r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else ""))
            print(name, k, mode, out, flush=True)
