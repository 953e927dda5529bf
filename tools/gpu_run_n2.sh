#!/bin/bash
# N = 2 bench of the value leg: photon dispatch (default) vs the whole-set exchange
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "n2 dispatch rc=$?"
tail -5 gpurun_out/r2l_bench_n2.err
python - <<'PY'
import json
for f in ['gpurun_out/r2l_bench_n2.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'], d['shards']['photons_in_hierarchy_per_rank'], d.get('value_exchange','')[:40])
PY
GVPM_VALUE_EXCHANGE=allgather timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_n2_allgather.json 2> gpurun_out/r2l_bench_n2_allgather.err; echo "n2 allgather rc=$?"
tail -3 gpurun_out/r2l_bench_n2_allgather.err
python - <<'PY'
import json
for f in ['gpurun_out/r2l_bench_n2_allgather.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'], d['shards']['photons_in_hierarchy_per_rank'])
PY
