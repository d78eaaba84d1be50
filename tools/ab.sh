#!/bin/bash
run() { echo "== $1 k=$2 $3"; env $1 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --k $2 $3 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for x in d['per_k']: print('  ', x['k'], x['ms'], x['gbases_s'], x['frac_of_peak'])"; }
run "CFRK_K4=0" 4
run "CFRK_K4=6" 4
run "CFRK_K4=7" 4
run "CFRK_K4=8" 4
run "CFRK_K4=0" 4
