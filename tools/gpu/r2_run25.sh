#!/bin/bash
# final: full GPU suite, bench N=1 (both arms), ncu of the k=2 lane kernel (one ncu invocation)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_t25.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t25.log
timeout 900 python bench.py > $O/r2_b25.json 2> $O/r2_b25.err; echo "bench rc=$?"; tail -2 $O/r2_b25.err
timeout 600 python bench.py --impl reference > $O/r2_b25_ref.json 2> $O/r2_b25_ref.err; echo "bench ref rc=$?"
B="python bench.py --k 3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks"
timeout 300 $B > $O/plain_r2_k3b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_lane" -s 1 -c 1 -o $O/prof_r2_k3_final -f $B > $O/ncu_r2_k3b.log 2>&1
echo "ncu rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2_b25.json").read().strip().splitlines()[-1])
print("value", d["value"], "min_frac", d.get("min_frac"), "e2e", d["e2e"]["value"], d["e2e"].get("dense_host"))
for v in d["per_k"]: print(v.get("k"), v.get("gbases_s"), v.get("frac_of_peak"))
print("checks", d["checks"]["all_ok_all_ranks"])
for k,v in d["configs"].items(): print(k, v.get("gbases_s"), v.get("value"))
print(open("$O/r2_b25_ref.json").read()[:600])
PY
