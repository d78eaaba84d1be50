#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-8}
: > $O/r2_hr21_n$N.log
for v in "2 4" "2 8" "1 8"; do
  set -- $v
  echo "== ctas/SM $1 unroll $2" >> $O/r2_hr21_n$N.log
  CFRK_HIST_REDUCE_CTAS=$1 CFRK_HIST_REDUCE_UNROLL=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29731 tests/dist_hist_reduce.py 2>&1 | grep "HIST_REDUCE_OK\|Error\|error" >> $O/r2_hr21_n$N.log
done
cat $O/r2_hr21_n$N.log
