#!/bin/bash
# N = 8: side-stream dispatch with small persistent grids
set -x
mkdir -p gpurun_out
run() {  # name nproc port [env...]
  local name=$1 np=$2 port=$3; shift 3
  env "$@" GVPM_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_$name.json 2> gpurun_out/r2z_bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2z_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', d['ms_per_step'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d.get('result_collection_verified'))
PY
}
run n8_side32 8 29581 GVPM_DISPATCH_STREAM=side GVPM_DISPATCH_CTAS=32
run n8_side96 8 29582 GVPM_DISPATCH_STREAM=side GVPM_DISPATCH_CTAS=96
timeout 300 python -m pytest tests/test_gpu_dispatch.py -m gpu -x -q 2>&1 | tail -3
