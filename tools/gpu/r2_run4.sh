#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 400 python tests/manual/lane_variants.py --time-reads 4000000 > $O/r2_lane_default.log 2>&1; L=$?
echo "lane default rc=$L"; tail -6 $O/r2_lane_default.log
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "small_spans or larger_than or gzip or never_cross" > $O/r2_t_cli2.log 2>&1; echo "cli tests rc=$?"; tail -4 $O/r2_t_cli2.log
[ $L -eq 0 ] && bash tools/gpu/r2_ncu_lane.sh
