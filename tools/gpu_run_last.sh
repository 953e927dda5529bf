#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 160 python -m pytest tests/test_gpu_frustum_grid.py tests/test_gpu_dispatch.py tests/test_gpu_pruned_build.py -m gpu -x -q 2>&1 | tail -4
