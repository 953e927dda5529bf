#!/bin/bash
# round 2, run K: grid traverse (single flush site), plane functor (fast reciprocals, 4 CTAs/SM), beam reconnection; cfg5 / cfg4 / cfg3 bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2k_pytest_gpu.log
for wl in cfg5 cfg4 cfg3; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_$wl.json 2> gpurun_out/r2k_bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2k_bench_$wl.json'))
print('$wl', d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])
PY
  tail -3 gpurun_out/r2k_bench_$wl.err
done
