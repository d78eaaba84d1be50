#!/bin/bash
# 2 GPUs: histogram sum over NVLink peer memory
cd "$(dirname "$0")/../.."
O=gpurun_out
nvidia-smi topo -m > $O/r2_topo18.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29731 tests/dist_hist_reduce.py > $O/r2_hr18.log 2>&1; echo "rc=$?"
tail -25 $O/r2_hr18.log
