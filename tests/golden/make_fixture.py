#!/usr/bin/env python
"""Writes tests/golden/bre_tiny.gvpmfix: a seeded synthetic G-BRE iteration (inputs + the oracle's results) in the
on-disk fixture format a Mitsuba-side dump hook produces (gvpm_b200/host/gvpm_fixture.hpp).
Self-generated: it pins the format and guards regressions, it is not a vector from the reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as ge  # noqa: E402

ge.build_cpu_libs()
import gvpm_testlib as H  # noqa: E402
from gvpm_b200 import fixture as F  # noqa: E402
from oracle import binding as ob  # noqa: E402

c = H.make_case(n_photons=2500, w=16, h=12, scale=4.0, seed=0xF1C5)
ref = ob.bre_gather(c.photons, c.rays, c.medium, c.config, c.tri, c.radius, mode="brute", neighbours=True, threads=4)
out = os.path.join(ROOT, "tests", "golden", "bre_tiny.gvpmfix")
F.save(out, c.medium, c.config, c.radius, c.tri, c.photons, c.rays, ref.out, ref.offsets, ref.idx,
       producer="gvpm_b200 oracle (tests/golden/make_fixture.py)")
print(out, os.path.getsize(out), "bytes;", int(ref.counts[:, 0].sum()), "neighbours")
