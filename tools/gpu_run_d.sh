#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generate.py tests/test_gpu_frustum_grid.py tests/test_gpu_pruned_build.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_cfg5.json 2> gpurun_out/r2d_bench_cfg5.err; echo "cfg5 rc=$?"
tail -5 gpurun_out/r2d_bench_cfg5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench_cfg5.json'))
print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['shards'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2d_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "launches rc=$?"
