#!/bin/bash
# final: bench N=1 (both arms), launch list, ncu captures of the lane kernels and the dominant kernel
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_b17.json 2> $O/r2_b17.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/r2_b17_ref.json 2> $O/r2_b17_ref.err; echo "bench ref rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks"
timeout 300 $B > $O/plain_r2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_raw.csv $B > $O/ncu_r2_launches.log 2>&1
echo "launch list rc=$?"
for k in 1 2 3 4 8; do
  timeout 300 python bench.py --k $k --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks > $O/plain_r2_k$k.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_lane|dense_bigrow" -s 1 -c 1 -o $O/prof_r2_final_k$k -f \
      python bench.py --k $k --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks > $O/ncu_r2_final_k$k.log 2>&1
  echo "ncu k=$k rc=$?"
done
timeout 200 python tools/bench_sparse.py --reads 2000000 --read-len 150 --k 12 --key-bytes 4 --steps 2 > $O/plain_half.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_half -s 1 -c 1 -o $O/prof_r2_sparse_half -f \
    python tools/bench_sparse.py --reads 2000000 --read-len 150 --k 12 --key-bytes 4 --steps 2 > $O/ncu_half.log 2>&1
echo "ncu half rc=$?"
