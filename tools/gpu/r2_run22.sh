#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 2 --warmup 3 --configs c5 --no-cpu > $O/r2_b22_n$N.json 2> $O/r2_b22_n$N.err; echo "bench rc=$?"
tail -3 $O/r2_b22_n$N.err
python - <<PY
import json
d=json.loads(open("$O/r2_b22_n$N.json").read().strip().splitlines()[-1])
print(json.dumps(d["configs"]["C5_global_hist_k12"], indent=1))
PY
