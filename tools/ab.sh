#!/bin/bash
run() { echo "== $1 k=$2 $3"; env $1 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --k $2 $3 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms/step', d['ms_per_step'])
for x in d['per_k']: print('  ', x['k'], x['ms'], x['gbases_s'], x['frac_of_peak'])"; }
run "X=1" 4,5,6 "--fmt ascii"
run "X=1" 4,5,6 "--fmt packed"
run "X=1" 4,5,6 "--fmt codes"
