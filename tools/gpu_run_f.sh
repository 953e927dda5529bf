#!/bin/bash
# round 2, run F: all GPU tests on the committed tree, cfg5 bench line, full ncu captures of the beam / plane / vpm kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2f_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_cfg5.json 2> gpurun_out/r2f_bench_cfg5.err; echo "cfg5 rc=$?"
tail -3 gpurun_out/r2f_bench_cfg5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_beam_traverse|k_beam_shade" -c 2 -o gpurun_out/r2f_full_beams python bench.py --workload cfg3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2f_ncu_beams.log 2>&1; echo "ncu beams rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_plane_gather" -c 1 -o gpurun_out/r2f_full_planes python bench.py --workload cfg4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2f_ncu_planes.log 2>&1; echo "ncu planes rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vpm_traverse|k_vpm_shade" -c 2 -o gpurun_out/r2f_full_vpm python bench.py --workload cfg2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2f_ncu_vpm.log 2>&1; echo "ncu vpm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bre_shade|k_bre_grid_traverse" -c 2 -s 2 -o gpurun_out/r2f_full_bre python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_ncu_bre.log 2>&1; echo "ncu bre rc=$?"
ls -la gpurun_out | tail -12
