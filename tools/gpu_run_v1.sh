#!/bin/bash
# round 2, validation after the counting-sort build and the deterministic pruned compaction
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2v2_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v2_bench_cfg5.json 2> gpurun_out/r2v2_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
for f in ('r2v2_bench_cfg5',):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, 'frac', round(d['roofline']['frac'],4), 'l2', d['roofline'].get('l2',{}) and round(d['roofline']['l2']['frac'],4), 'traced', round(d['device_traced']['ms_per_step'],2), {k: round(v,3) for k,v in d['device_traced']['phases_ms'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2v2_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "launches rc=$?"
