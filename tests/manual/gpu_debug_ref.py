import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures as fx, oracle_binding as ob
out = os.path.join(ROOT, "gpurun_out", "dbg"); os.makedirs(out, exist_ok=True)
env = dict(os.environ, CUDA_VISIBLE_DEVICES="0")
for name, text in (("A_basic", fx.fx_basic()), ("G_short", fx.fx_short())):
    fa = os.path.join(out, name + ".fa"); open(fa, "w").write(text)
    for k in (2, 3):
        r = subprocess.run([os.path.join(ROOT, "oracle/_ref/cfrk_ref_gpu"), fa, os.path.join(out, f"{name}.{k}.ref"), str(k), "12", "8192"], env=env, capture_output=True)
        print(name, k, r.returncode, r.stdout[:300], r.stderr[:300])
        ob.run_cli(fa, os.path.join(out, f"{name}.{k}.or"), k, 8192)
