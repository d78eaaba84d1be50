#!/bin/bash
# last check of the shipped library: CLI + writer tests (runfile.cu was rebuilt), smoke
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cli.py -x -q -k "not two_gpus" > $O/r2_t30.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2_t30.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
