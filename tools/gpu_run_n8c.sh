#!/bin/bash
# N = 8 with per-step event trace (stderr)
set -x
mkdir -p gpurun_out
GVPM_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.err; echo "n8 rc=$?"
grep "trace rank" gpurun_out/r2v_bench_n8.err | head -20
python - <<'PY'
import json
for l in open('gpurun_out/r2v_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n8 inline', d['ms_per_step'], d['phases_ms'])
PY
GVPM_DISPATCH_STREAM=side GVPM_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_n8_side.json 2> gpurun_out/r2v_bench_n8_side.err; echo "n8 side rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2v_bench_n8_side.json'):
    if l.startswith('{'):
        d=json.loads(l); print('n8 side', d['ms_per_step'], d['phases_ms'])
PY
timeout 300 python -m pytest tests/test_gpu_dispatch.py -m gpu -x -q 2>&1 | tail -3
