#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", json.dumps(d["e2e"])[:900])
except Exception as e:
    print("unreadable", e)
PY
tail -8 gpurun_out/r2_bench_n$N.err
