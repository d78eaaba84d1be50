#!/bin/bash
# N GPUs: bench with C3 + C5 only (histogram exchange over peer memory inside bench.py)
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 2 --warmup 3 --configs c3,c5 --no-cpu > $O/r2_b24_n$N.json 2> $O/r2_b24_n$N.err; echo "bench rc=$?"
tail -3 $O/r2_b24_n$N.err
python - <<PY
import json
d=json.loads(open("$O/r2_b24_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ("config","check")}))
print(d["configs"]["C5_global_hist_k12"]["check"])
PY
