#!/bin/bash
# N = 8 / 4: copy-engine result collection, owner-map dispatch; inline vs side-stream dispatch
set -x
mkdir -p gpurun_out
run() {  # name nproc port [env...]
  local name=$1 np=$2 port=$3; shift 3
  env "$@" GVPM_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_$name.json 2> gpurun_out/r2y_bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2y_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', d['ms_per_step'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d.get('result_collection_verified'))
PY
}
run n8_inline 8 29571 GVPM_X=1
run n8_side 8 29572 GVPM_DISPATCH_STREAM=side
run n4_inline 4 29573 GVPM_X=1
