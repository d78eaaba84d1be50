"""On the GPU box: the real reference binary (oracle/_ref/cfrk_ref_gpu, built from the unmodified
sources for sm_100) vs the oracle vs our CLI, on the edge-case fixtures. SURVEY 8c [verify on GPU]."""
import filecmp
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures as fx  # noqa: E402
import oracle_binding as ob  # noqa: E402

ref = os.path.join(ROOT, "oracle", "_ref", "cfrk_ref_gpu")
ours = os.path.join(ROOT, "bin", "cfrk")
env = dict(os.environ, CUDA_VISIBLE_DEVICES="0")
tmp = tempfile.mkdtemp()
bad = 0
cases = [(n, t, ks, (8192,)) for n, t, ks in fx.EDGE_SET if n != "R_ragged"] + list(fx.CHUNK_SET)
for name, text, ks, chunks in cases:
    fa = os.path.join(tmp, name + ".fa")
    open(fa, "w").write(text)
    for k in ks:
        if k > 6:
            continue
        for ch in chunks:
            o_ref, o_or, o_us = (os.path.join(tmp, f"{name}.{k}.{ch}.{w}") for w in ("ref", "or", "us"))
            r = subprocess.run([ref, fa, o_ref, str(k), "12", str(ch)], env=env, capture_output=True, timeout=120)
            ob.run_cli(fa, o_or, k, ch)
            u = subprocess.run([ours, fa, o_us, str(k), "12", str(ch)], env=env, capture_output=True, timeout=120)
            same_ref = os.path.exists(o_ref) and filecmp.cmp(o_ref, o_or, shallow=False)
            same_us = os.path.exists(o_us) and filecmp.cmp(o_us, o_or, shallow=False)
            if not (same_ref and same_us):
                bad += 1
            print(f"{name:12s} k={k} chunk={ch:5d} ref_rc={r.returncode} ref==oracle:{same_ref} "
                  f"ours_rc={u.returncode} ours==oracle:{same_us} {u.stderr.decode()[:200]}")
print("MISMATCHES:", bad)
sys.exit(1 if bad else 0)
