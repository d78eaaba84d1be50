#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -x -q > $O/r2_t9.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_t9.log
timeout 120 python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_smoke.log
CFRK_TRACE=1 timeout 300 python tests/manual/host_op_timing.py > $O/r2_hostop2.log 2> $O/r2_hostop2.err; grep -v "host op" $O/r2_hostop2.log | head -14
grep "host op" $O/r2_hostop2.err | sed -n 26,30p
