#!/bin/bash
# round 2, run A: GPU tests, all bench workloads, launch list + one full ncu capture of the cfg5 gather kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_cfg5.json 2> gpurun_out/r2a_bench_cfg5.err; echo "cfg5 rc=$?"
for wl in cfg2 cfg3 cfg4 beams1080; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2a_bench_$wl.json 2> gpurun_out/r2a_bench_$wl.err; echo "$wl rc=$?"
  tail -c 600 gpurun_out/r2a_bench_$wl.json; tail -3 gpurun_out/r2a_bench_$wl.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bre_traverse|k_bre_shade" -c 2 -s 2 -o gpurun_out/r2a_full_bre python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu_bre.log 2>&1; echo "ncu rc=$?"
for wl in cfg2 cfg3 cfg4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "launches $wl rc=$?"
done
ls -la gpurun_out
