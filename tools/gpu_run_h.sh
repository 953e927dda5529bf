#!/bin/bash
# round 2, run H: parity tests of the restructured beam shade + plane bundle cull + dispatch, then cfg3 / cfg4 bench lines
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dispatch.py tests/test_gpu_beams_parity.py tests/test_gpu_planes_parity.py tests/test_gpu_sppm_beams_parity.py tests/test_gpu_full_size.py tests/test_gpu_host_and_gradient.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2h_pytest.log
for wl in cfg3 cfg4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_$wl.json 2> gpurun_out/r2h_bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2h_bench_$wl.json'))
print('$wl', d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])
PY
  tail -3 gpurun_out/r2h_bench_$wl.err
done
