#!/bin/bash
# 2-GPU box: link tests, lane timing after the hand-off change, two-GPU CLI test, bench at N=2
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_vs_reference.py -x -q > $O/r2_t7_ref.log 2>&1; echo "ref tests rc=$?"; tail -4 $O/r2_t7_ref.log
timeout 300 python tests/manual/lane_variants.py --time-reads 4000000 > $O/r2_lane_default2.log 2>&1; echo "lane rc=$?"; tail -4 $O/r2_lane_default2.log
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "two_gpus" > $O/r2_t7_2gpu.log 2>&1; echo "2gpu test rc=$?"; tail -4 $O/r2_t7_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_b7_n2.json 2> $O/r2_b7_n2.err; echo "bench n2 rc=$?"; tail -c 600 $O/r2_b7_n2.err
