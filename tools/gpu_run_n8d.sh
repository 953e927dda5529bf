#!/bin/bash
# N = 8: what does the concurrent NCCL result gather cost the build / dispatch kernels?
set -x
mkdir -p gpurun_out
run() {  # name port [env...]
  local name=$1 port=$2; shift 2
  env "$@" GVPM_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_$name.json 2> gpurun_out/r2w_bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2w_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', d['ms_per_step'], d['phases_ms'])
PY
}
run nocollect 29561 GVPM_COLLECT=none
run nch2 29562 NCCL_MAX_NCHANNELS=2
run nch2_side 29563 NCCL_MAX_NCHANNELS=2 GVPM_DISPATCH_STREAM=side
