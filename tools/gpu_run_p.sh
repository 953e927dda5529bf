#!/bin/bash
# round 2, run P: dispatch tests after the algebraic footprint / by-value grids; launch list of an 8-rank dispatch on one GPU; cfg5 bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dispatch.py tests/test_gpu_frustum_grid.py tests/test_gpu_pruned_build.py tests/test_gpu_generate.py tests/test_gpu_bre_parity.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2t_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches_dispatch8.csv python tools/time_dispatch.py --world 8 --iters 2 > gpurun_out/r2t_time_dispatch.log 2>&1; echo "dispatch8 rc=$?"
tail -4 gpurun_out/r2t_time_dispatch.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_cfg5.json 2> gpurun_out/r2t_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench_cfg5.json'))
print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
PY
