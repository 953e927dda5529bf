#!/bin/bash
# N = 8 / 4 with cost-balanced bands
set -x
mkdir -p gpurun_out
run() {  # name nproc port [env...]
  local name=$1 np=$2 port=$3; shift 3
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_$name.json 2> gpurun_out/r2q_bench_$name.err; echo "$name rc=$?"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r2q_bench_$name.err | tail -4
  python - <<PY
import json
for l in open('gpurun_out/r2q_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', d['ms_per_step'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d['shards']['photons_in_hierarchy_per_rank'], d['shards']['rays_per_rank'], d['shards']['balance'])
PY
}
run n8 8 29541 GVPM_X=1
run n8_c4 8 29542 GVPM_BAND_CYCLES=4
run n4 4 29543 GVPM_X=1
