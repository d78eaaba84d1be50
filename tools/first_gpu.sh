#!/bin/bash
# first contact with the GPU: parity tests, then a quick timing
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
