#!/usr/bin/env python
"""Kernel-tuning helper: time build + G-Planes 0D gather on a cfg4-sized synthetic case (CUDA events).

    python tools/time_planes.py [--planes N] [--w W --h H] [--sheet]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import gvpm_testlib as H  # noqa: E402
from gvpm_b200.api import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--planes", type=int, default=200_000)
ap.add_argument("--w", type=int, default=1280)
ap.add_argument("--h", type=int, default=720)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--sheet", action="store_true")
a = ap.parse_args()

c = H.make_plane_case(n_planes=a.planes, w=a.w, h=a.h, seed=0xC0FFEE + 4, sheet=a.sheet)
ctx = Context(0)
ctx.set_medium(c.medium)
ctx.set_config(c.config)
ctx.upload_planes(c.planes)
ctx.upload_rays(c.rays)
bs, gs = [], []
for i in range(a.reps + 1):
    ctx.build_planes()
    out, counts = ctx.gather_planes()
    b, gm = ctx.last_timings()
    if i:
        bs.append(b)
        gs.append(gm)
hits = int(counts[:, 0].sum())
print(f"planes={a.planes} rays={c.rays.n} sheet={a.sheet} build_ms={np.mean(bs):.3f} gather_ms={np.mean(gs):.3f} "
      f"rays/s={c.rays.n / np.mean(gs) * 1e3:.3e} hits={hits} ({hits / c.rays.n:.1f}/ray, "
      f"{hits / (c.rays.n * a.planes) * 100:.2f}% of pairs) pair-tests/s={c.rays.n * a.planes / np.mean(gs) * 1e3:.3e} "
      f"checksum={float(out.astype(np.float64).sum()):.6e} finite={bool(np.isfinite(out).all())}")
ctx.close()
