#!/usr/bin/env python
"""Check the CUDA gather (and the CPU oracle) against a fixture file (gvpm_b200/fixture.py).

    python tools/check_fixture.py FIXTURE [--no-gpu] [--rtol 1e-4]

With a fixture dumped from a real Mitsuba run of the reference the "expected" sections are the
reference's own per-ray results: this is the parity check of north_star (neighbour index sets bit-exact, radiance
within 1e-4).  Exit status 0 = everything that could be compared agrees.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gvpm_b200 import fixture as F  # noqa: E402


def rel_err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    floor = 1e-3 * np.abs(ref).max() if ref.size else 1.0
    return float((np.abs(got - ref) / np.maximum(np.abs(ref), floor if floor > 0 else 1.0)).max()) if ref.size else 0.0


def differing_rays(off_a, idx_a, off_b, idx_b):
    """ray indices whose neighbour lists differ (bit 31 ignored)"""
    a, b = idx_a & 0x7fffffff, idx_b & 0x7fffffff
    bad = []
    for i in range(len(off_a) - 1):
        if not np.array_equal(np.sort(a[int(off_a[i]):int(off_a[i + 1])]), np.sort(b[int(off_b[i]):int(off_b[i + 1])])):
            bad.append(i)
    return bad


def report_sets(what, bad):
    print(f"{what}: " + ("identical" if not bad else f"{len(bad)} rays DIFFER (first: {bad[:8]})"))
    return not bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fixture")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--rtol", type=float, default=1e-4)
    a = ap.parse_args()
    fx = F.load(a.fixture)
    print(f"fixture {a.fixture}: producer '{fx.producer}', {fx.photons.n} photons, {fx.rays.n} rays, radius {fx.radius:g}, "
          f"expected results: {'yes' if fx.expected_out is not None else 'no'}")
    ok = True
    from oracle import binding as ob   # test infrastructure: this tool is a checker, not the product path
    ref = ob.bre_gather(fx.photons, fx.rays, fx.medium, fx.config, fx.tri, fx.radius, mode="brute", neighbours=True)
    if fx.expected_out is not None:
        e = rel_err(ref.out, fx.expected_out)
        print(f"oracle vs expected radiance: max rel err {e:.3e}")
        ok &= e <= a.rtol
    if fx.expected_offsets is not None:
        ok &= report_sets("oracle vs expected neighbour sets",
                          differing_rays(ref.offsets, ref.idx, fx.expected_offsets, fx.expected_idx))
    if not a.no_gpu:
        from gvpm_b200.api import Context
        ctx = Context(0)
        ctx.set_medium(fx.medium)
        ctx.set_config(fx.config)
        ctx.set_occluders(fx.tri)
        ctx.upload_photons(fx.photons)
        ctx.build_points(fx.radius)
        ctx.upload_rays(fx.rays)
        out, counts = ctx.gather_bre()
        offsets, idx = ctx.dump_neighbours_bre()
        ctx.close()
        same = np.array_equal(offsets, ref.offsets) and np.array_equal(idx, ref.idx)
        e = rel_err(out, ref.out)
        print(f"GPU vs oracle: neighbour sets {'identical' if same else 'DIFFERENT'}, radiance max rel err {e:.3e}")
        ok &= same and e <= a.rtol
        if fx.expected_out is not None:
            e = rel_err(out, fx.expected_out)
            print(f"GPU vs expected radiance: max rel err {e:.3e}")
            ok &= e <= a.rtol
        if fx.expected_offsets is not None:
            ok &= report_sets("GPU vs expected neighbour sets",
                              differing_rays(offsets, idx, fx.expected_offsets, fx.expected_idx))
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
