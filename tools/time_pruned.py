#!/usr/bin/env python
"""Tuning helper: time gvpm_build_points_for_rays + gather for one rank of a band partition on a cfg5-sized case.

    python tools/time_pruned.py --world 8 --rank 3 [--cycles 2] [--reps 3]
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/time_pruned.py ...
"""
import argparse
import os
import sys

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
ap.add_argument("--world", type=int, default=2)
ap.add_argument("--rank", type=int, default=0, help="-1: every rank in turn")
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

med = g.make_medium()
ph, paths = g.synth_photons(a.photons, med, seed=0xC0FFEE + 5, threads=os.cpu_count() or 8)
full = g.synth_rays(a.w, a.h, seed=0xC0FFEE + 6, block=-32)
ctx = Context(0)
ctx.set_medium(med)
ctx.set_config(g.make_config(a.w, a.h))
ctx.set_occluders(g.synth_occluders())
ctx.upload_photons(ph)
r = g.bre_radius(a.scale)
for rank in (range(a.world) if a.rank < 0 else [a.rank]):
    rays = full.take(shard.band_indices(full.px, full.py, a.w, a.h, a.world, rank, a.cycles))
    ctx.upload_rays(rays)
    ctx.build_points(r)
    out_ptr, _ = ctx.gather_bre_device()
    res = {}
    for mode in ("full", "pruned"):
        bs, gs = [], []
        for i in range(a.reps + 1):
            kept = ctx.build_points_for_rays(r) if mode == "pruned" else (ctx.build_points(r) or a.photons)
            ctx.gather_bre_into(out_ptr, None)
            ctx.sync()
            b, gm = ctx.last_timings()
            if i:
                bs.append(b)
                gs.append(gm)
        out, counts = ctx.gather_bre()
        res[mode] = (np.mean(bs), np.mean(gs), kept, int(counts[:, 0].sum()), float(out.astype(np.float64).sum()))
        print(f"{mode}: world={a.world} rank={rank} cycles={a.cycles} rays={rays.n} photons_in_hierarchy={kept} "
              f"build_ms={np.mean(bs):.3f} gather_ms={np.mean(gs):.3f} H={res[mode][3]} checksum={res[mode][4]:.6e}",
              flush=True)
    assert res["full"][3] == res["pruned"][3], "neighbour totals differ"
ctx.close()
