#!/bin/bash
# N = 8 bench of the value leg: photon dispatch (default) vs the whole-set exchange; N = 4 dispatch
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() {  # name nproc port [env...]
  local name=$1 np=$2 port=$3; shift 3
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_$name.json 2> gpurun_out/r2o_bench_$name.err; echo "$name rc=$?"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/r2o_bench_$name.err | tail -4
  python - <<PY
import json
for l in open('gpurun_out/r2o_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', d['ms_per_step'], d['phases_ms'], 'e2e', d['e2e']['ms_per_step'], d['shards']['photons_in_hierarchy_per_rank'], d.get('value_exchange','')[:30])
PY
}
run n8 8 29521 GVPM_X=1
run n8_allgather 8 29522 GVPM_VALUE_EXCHANGE=allgather
run n4 4 29523 GVPM_X=1
