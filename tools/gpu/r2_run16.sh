#!/bin/bash
# 16 MiB windows and row slots by default, workers set up while the reader opens
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_writer.py -x -q > $O/r2_t16.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t16.log
CFRK_BENCH_CLI_TRACE=1 timeout 900 python tools/bench_cli.py --nt 16 --md5 \
  --envs "default:;again:;pwrite:CFRK_WRITER=pwrite" > $O/r2_cli16.json 2> $O/r2_cli16_trace.log; echo "cli rc=$?"
cat $O/r2_cli16.json | cut -c1-1100
for k in 2 6 8; do
CFRK_BENCH_CLI_TRACE=1 timeout 600 python tools/bench_cli.py --nt 16 --k $k --reads $([ $k = 8 ] && echo 20000 || echo 500000) --runs all_rows_dense,all_rows_sparse > $O/r2_cli16_k$k.json 2> $O/r2_cli16_k${k}_trace.log; echo "cli k=$k rc=$?"
cat $O/r2_cli16_k$k.json | cut -c1-700
done
