#!/bin/bash
# round 2, run Q: shared-frame dispatch (owner map) tests; launch list of an 8-rank dispatch on one GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dispatch.py tests/test_gpu_frustum_grid.py tests/test_gpu_beams_parity.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2u_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_dispatch8.csv python tools/time_dispatch.py --world 8 --iters 2 > gpurun_out/r2u_time_dispatch.log 2>&1; echo "dispatch8 rc=$?"
tail -3 gpurun_out/r2u_time_dispatch.log
