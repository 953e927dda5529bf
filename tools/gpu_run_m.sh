#!/bin/bash
# round 2, run M: bounded (sync-free) sharded build + overflow recovery, dispatch, beam traversal with the chord pre-test; cfg3 bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2m_pytest_gpu.log
for wl in cfg3 beams1080; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_$wl.json 2> gpurun_out/r2m_bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2m_bench_$wl.json'))
print('$wl', d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline'].get('pairs_shaded'))
PY
  tail -3 gpurun_out/r2m_bench_$wl.err
done
