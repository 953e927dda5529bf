#!/usr/bin/env python
"""Time gvpm_poisson_solve against the reference's own solver (oracle/_ref, OpenMP backend on the host cores) on a
synthetic gradient-domain image set.

    python tools/time_poisson.py [--w 1920 --h 1080] [--presets L2D L1D] [--reps 3] [--no-ref]
"""
import argparse
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("mpg", os.path.join(ROOT, "tests", "golden", "make_poisson_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)

from gvpm_b200.api import Context  # noqa: E402
from oracle import poisson_ref as pr  # noqa: E402  (checker / CPU arm, never the product path)

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--presets", nargs="+", default=["L2D", "L1D"])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-ref", action="store_true")
a = ap.parse_args()

_, tp, dx, dy, direct = G.images(a.h, a.w, 11)
ctx = Context(0)
for preset in a.presets:
    ctx.poisson_solve(tp, dx, dy, direct, preset=preset)
    ms = []
    for _ in range(a.reps):
        rec = ctx.poisson_solve(tp, dx, dy, direct, preset=preset)
        ms.append(ctx.last_poisson_ms())
    line = {"solver": "gvpm_poisson_solve", "preset": preset, "w": a.w, "h": a.h, "gpu_ms_incl_copies": float(np.mean(ms))}
    if not a.no_ref and pr.available():
        t0 = time.perf_counter()
        want = pr.solve(tp, dx, dy, direct, backend="OpenMP", **pr.preset(preset))
        line["reference_openmp_ms"] = (time.perf_counter() - t0) * 1e3
        line["cores"] = os.cpu_count()
        line["max_rel_err"] = float(np.abs(rec.astype(np.float64) - want).max() / np.abs(want).max())
        line["speedup"] = line["reference_openmp_ms"] / line["gpu_ms_incl_copies"]
    if not a.no_ref and pr.cuda_available():
        # the reference's own CUDA backend (BackendCUDA.cu compiled for sm_100a) on the same GPU, host arrays in and out
        # like gvpm_poisson_solve: wall clock around the whole Solver sequence (import, setupBackend, solve, export)
        pr.solve(tp, dx, dy, direct, backend="CUDA", **pr.preset(preset))       # warm-up (context, allocations)
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            got = pr.solve(tp, dx, dy, direct, backend="CUDA", **pr.preset(preset))
            ts.append((time.perf_counter() - t0) * 1e3)
        line["reference_cuda_ms"] = float(np.mean(ts))
        line["gpu_wall_ms"] = None
        t0 = time.perf_counter()
        ctx.poisson_solve(tp, dx, dy, direct, preset=preset)
        line["gpu_wall_ms"] = (time.perf_counter() - t0) * 1e3
        line["vs_reference_cuda"] = line["reference_cuda_ms"] / line["gpu_wall_ms"]
        line["max_rel_err_vs_reference_cuda"] = float(np.abs(rec.astype(np.float64) - got).max() / np.abs(got).max())
    print(json.dumps(line), flush=True)
ctx.close()
