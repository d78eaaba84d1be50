#!/bin/bash
run() { echo "== $1 k=$2 $3"; env $1 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --k $2 $3 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for x in d['per_k']: print('  ', x['k'], x['ms'], x['gbases_s'], x['frac_of_peak'])"; }
run "CFRK_K7=1" 7
run "CFRK_K7=0 CFRK_BIG_TILE_KB=64 CFRK_BIG_CTAS=3" 7
run "CFRK_K7=0 CFRK_BIG_TILE_KB=32 CFRK_BIG_CTAS=4" 7
run "CFRK_K7=0 CFRK_BIG_TILE_KB=32 CFRK_BIG_CTAS=5" 7
run "CFRK_K7=0 CFRK_BIG_TILE_KB=32 CFRK_BIG_CTAS=6" 7
run "CFRK_K7=0 CFRK_BIG_TILE_KB=16 CFRK_BIG_CTAS=6" 7
run "CFRK_K7=1" 7
