#!/usr/bin/env python
"""Kernel-tuning helper: time build + G-BRE gather on a cfg5-sized synthetic case (device-side CUDA events).

    GVPM_B200_LIB=build/variants/lib_w4_b6.so python tools/time_gather.py [--photons N] [--scale S] [--w W --h H]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gvpm_b200 as g  # noqa: E402
from gvpm_b200.api import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--photons", type=int, default=10_000_000)
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--block", type=int, default=-32, help="negative: Z-order inside 32x32 blocks (bench.py layout)")
ap.add_argument("--counts", action="store_true", help="time the gather WITH per-ray counts (no traversal prefilter)")
a = ap.parse_args()

med = g.make_medium()
ph, paths = g.synth_photons(a.photons, med, seed=0xC0FFEE + 5, threads=os.cpu_count() or 8)
rays = g.synth_rays(a.w, a.h, seed=0xC0FFEE + 6, block=a.block)
ctx = Context(0)
ctx.set_medium(med)
ctx.set_config(g.make_config(a.w, a.h))
ctx.set_occluders(g.synth_occluders())
ctx.upload_photons(ph)
ctx.upload_rays(rays)
r = g.bre_radius(a.scale)
bs, gs, ts, ss = [], [], [], []
ctx.build_points(r)
out_ptr, _ = ctx.gather_bre_device()
for i in range(a.reps + 1):
    ctx.build_points(r)
    if a.counts:
        ctx.gather_bre_device()
    else:
        ctx.gather_bre_into(out_ptr, None)   # what bench.py times: filters applied in the traversal
        ctx.sync()
    b, gm = ctx.last_timings()
    t, s, pairs = ctx.last_gather_detail()
    if i:
        bs.append(b)
        gs.append(gm)
        ts.append(t)
        ss.append(s)
out, counts = ctx.gather_bre()
print(f"lib={os.environ.get('GVPM_B200_LIB', 'default')} photons={a.photons} rays={rays.n} scale={a.scale} "
      f"build_ms={np.mean(bs):.3f} gather_ms={np.mean(gs):.3f} (traverse {np.mean(ts):.3f} shade {np.mean(ss):.3f} "
      f"pairs {pairs}) rays/s={rays.n / np.mean(gs) * 1e3:.3e} "
      f"H={int(counts[:, 0].sum())} C={int(counts[:, 1].sum())} checksum={float(out.astype(np.float64).sum()):.6e}")
ctx.close()
