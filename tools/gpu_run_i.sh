#!/bin/bash
# round 2, run I: whole GPU suite, cfg5 bench (transposed shade reduction), L2 peak, ncu of the new plane / beam shade kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2i_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_cfg5.json 2> gpurun_out/r2i_bench_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench_cfg5.json'))
print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
PY
python tools/measure_l2.py > gpurun_out/r2i_l2_peak.json 2> gpurun_out/r2i_l2_peak.err; cat gpurun_out/r2i_l2_peak.json; tail -3 gpurun_out/r2i_l2_peak.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_plane_gather" -c 1 -o gpurun_out/r2i_full_planes python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2i_ncu_planes.log 2>&1; echo "ncu planes rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_beam_shade" -c 1 -o gpurun_out/r2i_full_beam_shade python bench.py --workload cfg3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2i_ncu_beam_shade.log 2>&1; echo "ncu beam shade rc=$?"
