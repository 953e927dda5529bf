#!/usr/bin/env python
"""Tuning helper: one GPU plays all `world` ranks of the photon dispatch on a cfg5-sized case (contexts wired in-process),
so that the launch list (ncu --metrics gpu__time_duration.sum) shows what one rank's classify / scan / emit kernels and
the build over an inbox cost.  The peers' inboxes are local here (no NVLink): kernel time, not link time.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/time_dispatch.py --world 8
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gvpm_b200 as g  # noqa: E402
from gvpm_b200 import shard  # noqa: E402
from gvpm_b200.api import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--photons", type=int, default=10_000_000)
ap.add_argument("--scale", type=float, default=0.1)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--no-view-dir", action="store_true")
a = ap.parse_args()

med = g.make_medium()
ph, paths = g.synth_photons(a.photons, med, seed=0xC0FFEE + 5, threads=os.cpu_count() or 8)
full = g.synth_rays(a.w, a.h, seed=0xC0FFEE + 6, block=-32)
r = g.bre_radius(a.scale)
n = ph.n - ph.n % a.world
n_slice = n // a.world
ctxs = []
for rank in range(a.world):
    c = Context(0)
    c.set_medium(med)
    c.set_config(g.make_config(a.w, a.h))
    c.set_occluders(g.synth_occluders())
    if not a.no_view_dir:
        c.set_view_direction((0.0, 0.0, 1.0))
    c.upload_rays(full.take(shard.band_indices(full.px, full.py, a.w, a.h, a.world, rank, a.cycles)))
    c.photon_staging(n)
    c.upload_photons_slice(ph.take(np.arange(rank * n_slice, (rank + 1) * n_slice)), n, rank * n_slice)
    ctxs.append(c)
blobs = [c.dispatch_export(a.world, n_slice) for c in ctxs]
for rank, c in enumerate(ctxs):
    c.dispatch_connect(blobs, rank)
for it in range(a.iters):
    b = it & 1
    t0 = time.perf_counter()
    for rank, c in enumerate(ctxs):
        c.dispatch_photons(b, n, rank * n_slice, n_slice, r, after_stream=c.stream())
    for c in ctxs:
        c.sync()
    t1 = time.perf_counter()
    kept = []
    for c in ctxs:
        kept.append(c.build_dispatched(b, r, want_kept=True))
        c.gather_bre_device()
        c.dispatch_release(b)
        c.sync()
    bt = [c.last_timings() for c in ctxs]
    print(f"iter {it}: dispatch of all ranks {1e3 * (t1 - t0):.3f} ms wall ({1e3 * (t1 - t0) / a.world:.3f} per rank), kept {kept}, "
          f"build ms {[round(x[0], 3) for x in bt]}, gather ms {[round(x[1], 3) for x in bt]}", flush=True)
for c in ctxs:
    c.close()
