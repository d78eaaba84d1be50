#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host or concurrent or pageable" > $O/r2_t10.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_t10.log
CFRK_TRACE=1 timeout 300 python tests/manual/host_op_timing.py > $O/r2_hostop3.log 2> $O/r2_hostop3.err; grep -v "host op" $O/r2_hostop3.log | head -14
grep "host op" $O/r2_hostop3.err | sed -n 26,30p
