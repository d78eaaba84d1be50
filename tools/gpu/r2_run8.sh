#!/bin/bash
# FASTQ tests, CLI end-to-end times, launch list + ncu captures of the bench kernels, bench N=1 (both arms)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "fastq or gzip or small_spans" > $O/r2_t8.log 2>&1; echo "fastq tests rc=$?"; tail -3 $O/r2_t8.log
for k in 2 4; do timeout 300 python tools/bench_cli.py --reads 2000000 --k $k; done > $O/r2_cli.jsonl 2> $O/r2_cli.err; cat $O/r2_cli.jsonl
timeout 300 python tools/bench_cli.py --reads 200000 --k 8 --runs all_rows_sparse,tail_only >> $O/r2_cli.jsonl 2>> $O/r2_cli.err; tail -1 $O/r2_cli.jsonl
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_b8.json 2> $O/r2_b8.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/r2_b8_ref.json 2> $O/r2_b8_ref.err; echo "bench ref rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks"
timeout 300 $B > $O/plain_r2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_raw.csv $B > $O/ncu_r2_launches.log 2>&1
echo "launch list rc=$?"
for k in 2 4 8; do
  timeout 300 python bench.py --k $k --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks > $O/plain_r2_k$k.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_lane|dense_bigrow" -s 1 -c 1 -o $O/prof_r2_final_k$k -f \
      python bench.py --k $k --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-checks > $O/ncu_r2_final_k$k.log 2>&1
  echo "ncu k=$k rc=$?"
done
