#!/bin/bash
# round 2, final scaling lines: N = 8, 4, 2 with the defaults the driver will run
set -x
mkdir -p gpurun_out
for np in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + np)) bench.py --gpus $np --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$np.json 2> gpurun_out/r2_bench_n$np.err; echo "n$np rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2_bench_n$np.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n$np', d['ms_per_step'], d['value'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d.get('result_collection_verified'), d['shards']['balance']['max_over_mean_equal_cost'])
PY
done
