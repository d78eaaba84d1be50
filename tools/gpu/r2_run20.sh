#!/bin/bash
# N GPUs: histogram sum over NVLink peer memory (p2p / nvls / NCCL)
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29731 tests/dist_hist_reduce.py 2>&1 | grep "HIST_REDUCE_OK\|Error\|error\|assert" > $O/r2_hr20_n$N.log
cat $O/r2_hr20_n$N.log
