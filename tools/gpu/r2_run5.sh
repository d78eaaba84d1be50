#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_t5.log 2>&1; echo "tests rc=$?"; tail -6 $O/r2_t5.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_b5.json 2> $O/r2_b5.err; echo "bench rc=$?"; tail -c 800 $O/r2_b5.err
timeout 300 python bench.py --steps 3 --warmup 3 --fmt ascii --no-configs --no-checks --no-cpu > $O/r2_b5_ascii.json 2> $O/r2_b5_ascii.err; echo "bench ascii rc=$?"
