#!/bin/bash
# round 2, run J: beam shade (templated, single batch site, transposed reduction), host drivers; cfg3 / beams1080 bench lines
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_beams_parity.py tests/test_gpu_sppm_beams_parity.py tests/test_gpu_host_drivers.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2j_pytest.log
for wl in cfg3 beams1080; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_$wl.json 2> gpurun_out/r2j_bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench_$wl.json'))
print('$wl', d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])
PY
  tail -3 gpurun_out/r2j_bench_$wl.err
done
