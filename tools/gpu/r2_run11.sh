#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 300 python tests/manual/lane_variants.py > $O/r2_lane_default3.log 2>&1; echo "lane parity rc=$?"; tail -4 $O/r2_lane_default3.log | cut -c1-80
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-configs --no-checks --configs small_k > $O/r2_b11.json 2> $O/r2_b11.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b11.json').read().strip().splitlines()[-1])
for p in d['per_k']: print(p['k'],p['gbases_s'],p['frac_of_peak'])
PY
CFRK_LANE_SPLIT_K3=2 timeout 300 python bench.py --steps 3 --warmup 3 --k 3 --no-cpu --no-e2e --no-configs --no-checks > $O/r2_b11_k3s2.json 2>> $O/r2_b11.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b11_k3s2.json').read().strip().splitlines()[-1])
for p in d['per_k']: print('split2',p['k'],p['gbases_s'],p['frac_of_peak'])
PY
