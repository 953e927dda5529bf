#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "n2 rc=$?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r2_bench_n2.err | tail -8
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n2', d['ms_per_step'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d['shards']['balance'], d.get('result_collection_verified'), d.get('result_collection'))
PY
