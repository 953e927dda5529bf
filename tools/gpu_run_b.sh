#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frustum_grid.py tests/test_gpu_bre_parity.py tests/test_gpu_pruned_build.py tests/test_gpu_sppm_parity.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg5.json 2> gpurun_out/r2b_bench_cfg5.err; echo "cfg5 rc=$?"
tail -5 gpurun_out/r2b_bench_cfg5.err
GVPM_PRUNE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg5_bvh.json 2> gpurun_out/r2b_bench_cfg5_bvh.err; echo "cfg5 bvh rc=$?"
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r2b_bench_cfg4.json 2> gpurun_out/r2b_bench_cfg4.err; echo "cfg4 rc=$?"
tail -3 gpurun_out/r2b_bench_cfg4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bre_grid_traverse" -c 1 -s 2 -o gpurun_out/r2b_full_grid python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_grid.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -12
