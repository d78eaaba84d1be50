#!/bin/bash
run() { echo "== $1 k=$2 $3"; env $1 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --k $2 $3 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for x in d['per_k']: print('  ', x['k'], x['ms'], x['gbases_s'], x['frac_of_peak'])"; }
run "CFRK_HANDOFF=0" 4,5
run "CFRK_HANDOFF=1 CFRK_K4=8 CFRK_K5=6" 4,5
run "CFRK_HANDOFF=1 CFRK_K4=9 CFRK_K5=7" 4,5
run "CFRK_HANDOFF=1 CFRK_K4=8 CFRK_K5=4" 4,5
