#!/bin/bash
# round 2, run G: dispatch tests (one GPU, contexts as ranks), frustum / pruned / generate regression, L2 peak, ncu of k_beam_shade
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dispatch.py tests/test_gpu_frustum_grid.py tests/test_gpu_pruned_build.py tests/test_gpu_generate.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2g_pytest.log
python tools/measure_l2.py > gpurun_out/r2g_l2_peak.json 2> gpurun_out/r2g_l2_peak.err; cat gpurun_out/r2g_l2_peak.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_beam_shade" -c 1 -o gpurun_out/r2g_full_beam_shade python bench.py --workload cfg3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2g_ncu_beam_shade.log 2>&1; echo "ncu beam shade rc=$?"
