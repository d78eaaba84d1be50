#!/bin/bash
# round 2, GPU call 3: lane kernel parity + timing, sparse tests, CLI tests, ncu of the lane kernels
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 400 python tests/manual/lane_variants.py --time-reads 4000000 > $O/r2_lane_default.log 2>&1; L=$?
echo "lane default rc=$L"; tail -6 $O/r2_lane_default.log
timeout 600 python -m pytest tests/test_gpu_sparse.py -x -q > $O/r2_t_sparse.log 2>&1; echo "sparse tests rc=$?"; tail -4 $O/r2_t_sparse.log
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q > $O/r2_t_cli.log 2>&1; echo "cli tests rc=$?"; tail -4 $O/r2_t_cli.log
for h in 1 0; do CFRK_SPARSE_HALF=$h timeout 300 python tools/bench_sparse.py --reads 10000000 --read-len 150 --k 12 --key-bytes 4; done > $O/r2_sparse_half_ab.log 2>&1; cat $O/r2_sparse_half_ab.log
CFRK_SPARSE_HALF=1 timeout 300 python tools/bench_sparse.py --reads 10000000 --read-len 150 --k 21 --key-bytes 8 >> $O/r2_sparse_half_ab.log 2>&1; tail -1 $O/r2_sparse_half_ab.log
if [ $L -eq 0 ]; then
  for k in 2 4; do
    timeout 200 python tests/manual/lane_variants.py --no-parity --ks $k --time-reads 2000000 > $O/plain_lane_k$k.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_lane -s 3 -c 1 -o $O/prof_r2_lane_k$k -f \
        python tests/manual/lane_variants.py --no-parity --ks $k --time-reads 2000000 > $O/ncu_lane_k$k.log 2>&1
    echo "ncu k=$k rc=$?"
  done
fi
