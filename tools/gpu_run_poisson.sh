#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/time_poisson.py --w 1920 --h 1080 > gpurun_out/r2_poisson_1080p.jsonl 2> gpurun_out/r2_poisson.err; echo "poisson rc=$?"; cat gpurun_out/r2_poisson_1080p.jsonl
timeout 600 python tools/time_poisson.py --w 1280 --h 720 > gpurun_out/r2_poisson_720p.jsonl 2>> gpurun_out/r2_poisson.err; cat gpurun_out/r2_poisson_720p.jsonl; tail -3 gpurun_out/r2_poisson.err
