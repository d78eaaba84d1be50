#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q > $O/r2_t17.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t17.log
CFRK_BENCH_CLI_TRACE=1 timeout 900 python tools/bench_cli.py --nt 16 --md5 > $O/r2_cli17.json 2> $O/r2_cli17_trace.log; echo "cli rc=$?"
cat $O/r2_cli17.json | cut -c1-1100
for k in 6 8; do
CFRK_BENCH_CLI_TRACE=1 timeout 600 python tools/bench_cli.py --nt 16 --k $k --reads $([ $k = 8 ] && echo 20000 || echo 500000) --runs all_rows_dense,all_rows_sparse > $O/r2_cli17_k$k.json 2> $O/r2_cli17_k${k}_trace.log; echo "cli k=$k rc=$?"
cat $O/r2_cli17_k$k.json | cut -c1-700
done
