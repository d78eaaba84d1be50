#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
for v in "" "CFRK_LANE_SPLIT_K3=2" "CFRK_LANE_SPLIT_K3=4"; do
env $v timeout 300 python bench.py --steps 3 --warmup 3 --k 1,2,3 --no-cpu --no-e2e --no-configs --no-checks > $O/r2_b12.json 2>> $O/r2_b12.err
python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2_b12.json').read().strip().splitlines()[-1])
print(sys.argv[1] or 'default', [(p['k'],p['gbases_s'],p['frac_of_peak']) for p in d['per_k']])
PY
done
