#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane.py tests/test_gpu_parity.py -x -q > $O/r2_t23.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t23.log
timeout 600 python bench.py --steps 3 --warmup 3 --configs small_k,c2b --no-cpu > $O/r2_b23.json 2> $O/r2_b23.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2_b23.json").read().strip().splitlines()[-1])
print("value", d["value"], "min_frac", d.get("min_frac"))
for v in d["per_k"]: print(v.get("k"), v.get("gbases_s"), v.get("frac_of_peak"))
print(d.get("checks"))
print({k: (v.get("gbases_s"), v.get("check")) for k,v in d.get("configs",{}).items()})
PY
