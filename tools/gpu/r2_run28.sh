#!/bin/bash
# sparse: in-lane mergers instead of full lane networks after every merge level
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py -x -q > $O/r2_t28.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t28.log
timeout 600 python bench.py --steps 2 --warmup 3 --configs c3,c4 --no-cpu --no-e2e > $O/r2_b28.json 2> $O/r2_b28.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/r2_b28.json").read().strip().splitlines()[-1])
for k,v in d["configs"].items(): print(k, v.get("gbases_s"), v.get("check"))
PY
