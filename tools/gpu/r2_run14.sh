#!/bin/bash
# full GPU suite on the final code
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_t14.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2_t14.log
