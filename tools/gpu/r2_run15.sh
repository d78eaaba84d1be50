#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse.py -x -q > $O/r2_t15.log 2>&1; echo "sparse tests rc=$?"; tail -3 $O/r2_t15.log
for h in 1 0; do for kk in "16 4" "21 8" "31 8"; do set -- $kk; CFRK_SPARSE_HALF=$h timeout 300 python tools/bench_sparse.py --reads 40 --read-len 5000000 --k $1 --key-bytes $2 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('half=$h k=',d['k'],d['value'],d['ms'])"; done; done
