#!/bin/bash
run() { echo "== $1 k=$2"; env $1 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --k $2 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for x in d['per_k']: print('  ', x['k'], x['ms'], x['frac_of_peak'])"; }
run "CFRK_BIGROW_MIN_K=5 CFRK_BIG_TILE_KB=64 CFRK_BIG_CTAS=3" 5,6
run "CFRK_BIGROW_MIN_K=5 CFRK_BIG_TILE_KB=32 CFRK_BIG_CTAS=6" 5,6
run "CFRK_BIGROW_MIN_K=5 CFRK_BIG_TILE_KB=128 CFRK_BIG_CTAS=2" 5,6
run "CFRK_BIG_TILE_KB=64 CFRK_BIG_CTAS=3" 7,8
run "CFRK_BIG_TILE_KB=64 CFRK_BIG_CTAS=3" 7,8
run "CFRK_BIG_TILE_KB=32 CFRK_BIG_CTAS=5" 7,8
