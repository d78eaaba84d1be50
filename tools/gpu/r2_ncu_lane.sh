#!/bin/bash
# ncu --set full of the lane kernels (k=2: bit planes, k=4: shared rows) and the sparse half-warp kernel
cd "$(dirname "$0")/../.."
O=gpurun_out
for k in 2 4; do
  timeout 200 python tests/manual/lane_variants.py --no-parity --ks $k --time-reads 2000000 > $O/plain_lane_k$k.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_lane -s 3 -c 1 -o $O/prof_r2_lane_k$k -f \
      python tests/manual/lane_variants.py --no-parity --ks $k --time-reads 2000000 > $O/ncu_lane_k$k.log 2>&1
  echo "ncu lane k=$k rc=$?"
done
timeout 200 python tools/bench_sparse.py --reads 2000000 --read-len 150 --k 12 --key-bytes 4 --steps 2 > $O/plain_half.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_half -s 1 -c 1 -o $O/prof_r2_sparse_half -f \
    python tools/bench_sparse.py --reads 2000000 --read-len 150 --k 12 --key-bytes 4 --steps 2 > $O/ncu_half.log 2>&1
echo "ncu half rc=$?"
