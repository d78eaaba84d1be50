#!/bin/bash
# closing run: full GPU suite + bench N=1 (both arms)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_t29.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t29.log
timeout 900 python bench.py > $O/r2_b29.json 2> $O/r2_b29.err; echo "bench rc=$?"; tail -2 $O/r2_b29.err
timeout 600 python bench.py --impl reference > $O/r2_b29_ref.json 2> $O/r2_b29_ref.err; echo "bench ref rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2_b29.json").read().strip().splitlines()[-1])
print("value", d["value"], "min_frac", d.get("min_frac"), "e2e", d["e2e"]["value"], d["e2e"].get("dense_host",{}).get("value"), "launches", d.get("gpu_launches"))
for v in d["per_k"]: print(v.get("k"), v.get("gbases_s"), v.get("frac_of_peak"))
print("checks", d["checks"]["all_ok_all_ranks"], "roofline", d["roofline"])
for k,v in d["configs"].items(): print(k, v.get("gbases_s"), v.get("value"))
print(open("$O/r2_b29_ref.json").read()[:200])
PY
