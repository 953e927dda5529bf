#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29644 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "n4 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n4.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n4', d['ms_per_step'], d['value'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d.get('result_collection_verified'))
PY
