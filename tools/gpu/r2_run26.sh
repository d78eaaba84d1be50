#!/bin/bash
# full bench at N GPUs (as the driver launches it)
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N > $O/r2_b26_n$N.json 2> $O/r2_b26_n$N.err; echo "bench rc=$?"
tail -2 $O/r2_b26_n$N.err
python - <<PY
import json
d=json.loads(open("$O/r2_b26_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "min_frac", d.get("min_frac"), "e2e", d["e2e"]["value"], "checks", d["checks"]["all_ok_all_ranks"])
for k,v in d["configs"].items(): print(k, v.get("gbases_s"), v.get("value"), v.get("ms_per_step"), v.get("count_ms"), v.get("allreduce_ms"), v.get("reduce"))
PY
