#!/bin/bash
# round 2, final single-GPU run: whole GPU suite, smoke, every bench workload, launch list + full ncu capture of the cfg5 gather
# kernels (-> profiles/traffic.json), L2 / HBM read peaks, Poisson vs the reference's own CUDA backend
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_final_smoke.log
python tools/measure_l2.py > gpurun_out/r2_l2_peak.json 2> gpurun_out/r2_l2_peak.err; cat gpurun_out/r2_l2_peak.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
tail -c 300 gpurun_out/r2_bench_cfg5.err
for wl in cfg2 cfg3 cfg4 beams1080; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json
for wl in ('cfg5','cfg2','cfg3','cfg4','beams1080'):
    try:
        d=json.load(open(f'gpurun_out/r2_bench_{wl}.json'))
        print(wl, round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phases_ms'].items()}, 'frac', round(d['roofline']['frac'],4), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(wl, 'FAILED', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bre_shade|k_bre_grid_traverse" -c 2 -s 2 -o gpurun_out/r2_full_bre python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_bre.log 2>&1; echo "ncu bre rc=$?"
timeout 600 python tools/time_poisson.py --w 1920 --h 1080 > gpurun_out/r2_poisson_1080p.jsonl 2> gpurun_out/r2_poisson.err; echo "poisson rc=$?"; cat gpurun_out/r2_poisson_1080p.jsonl
timeout 600 python tools/time_poisson.py --w 1280 --h 720 > gpurun_out/r2_poisson_720p.jsonl 2>> gpurun_out/r2_poisson.err; cat gpurun_out/r2_poisson_720p.jsonl; tail -3 gpurun_out/r2_poisson.err
