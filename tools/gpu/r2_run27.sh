#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_lane.py tests/test_gpu_parity.py tests/test_gpu_stages.py -x -q > $O/r2_t27.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t27.log
for v in 1 0; do
  CFRK_K3_TRANSPOSED=$v timeout 300 python bench.py --k 3 --steps 5 --warmup 3 --no-cpu --no-e2e --no-configs > $O/r2_b27_$v.json 2> $O/r2_b27_$v.err; echo "bench transposed=$v rc=$?"
  python - <<PY
import json
d=json.loads(open("$O/r2_b27_$v.json").read().strip().splitlines()[-1])
for v in d["per_k"]: print(v.get("k"), v.get("gbases_s"), v.get("frac_of_peak"))
print(d["checks"]["all_ok_all_ranks"])
PY
done
