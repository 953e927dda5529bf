#!/bin/bash
# round 2, last lines for profiles/: cfg5 with the CPU arm, launch list, full ncu of the gather pair + the build kernels
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_cfg5.json'))
print('cfg5', round(d['ms_per_step'],3), d['phases_ms'], d['roofline']['frac'], d['roofline']['l2'] and d['roofline']['l2']['frac'], d['cpu_baseline']['value'], d['e2e']['ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "launches rc=$?"
