#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_t6.log 2>&1; echo "tests rc=$?"; tail -6 $O/r2_t6.log
