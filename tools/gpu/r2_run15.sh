#!/bin/bash
# writer: rows formatted into the mapped output (text sizes from the GPU) vs buffers + pwrite
cd "$(dirname "$0")/../.."
O=gpurun_out
nproc > $O/r2_w15.log; free -g | head -2 >> $O/r2_w15.log
gcc -O2 -pthread -fopenmp -o /tmp/pwrite_scaling tools/host/pwrite_scaling.c && /tmp/pwrite_scaling /dev/shm 64 >> $O/r2_w15.log 2>&1
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_writer.py -x -q > $O/r2_t15.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t15.log
M=$((32<<20)); M16=$((16<<20))
CFRK_BENCH_CLI_TRACE=1 timeout 900 python tools/bench_cli.py --nt 16 --md5 --runs all_rows_dense,all_rows_sparse \
  --envs "mapped:;pwrite:CFRK_WRITER=pwrite;slot32:CFRK_ROW_SLOT_BYTES=$M;slot32win32:CFRK_ROW_SLOT_BYTES=$M,CFRK_WINDOW_BYTES=$M;slot16win16:CFRK_ROW_SLOT_BYTES=$M16,CFRK_WINDOW_BYTES=$M16" \
  > $O/r2_cli15.json 2> $O/r2_cli15_trace.log; echo "cli rc=$?"
cat $O/r2_cli15.json | cut -c1-900
CFRK_TRACE=1 python tools/writer_bench.py 414648 16 >> $O/r2_w15.log 2>&1
CFRK_TRACE=1 python tools/writer_bench.py 414648 32 >> $O/r2_w15.log 2>&1
cat $O/r2_w15.log | tail -60
