#!/usr/bin/env python
"""Writes tests/golden/poisson_small.npz: a seeded gradient-domain image set and the reconstructions the REFERENCE's
own solver returns for it (oracle/_ref/libgvpm_poisson_ref.so = src/integrators/poisson_solver compiled as is, naive
single-thread backend, which is deterministic).  Needs /root/reference (run in the build container)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import poisson_ref as pr  # noqa: E402

CASES = {
    "L2D": dict(preset="L2D"),
    "L1D": dict(preset="L1D"),
    "L2D_alpha05": dict(preset="L2D", alpha=0.5),
    "L1_short": dict(preset="L1D", irls_iter_max=4, cg_iter_max=30),
    "L2_tol": dict(preset="L2Q", cg_iter_check=7, cg_tolerance=1e-3),
    "L2D_no_throughput": dict(preset="L2D", no_throughput=True),
    "L2D_no_direct": dict(preset="L2D", no_direct=True),
}


def images(h=40, w=48, seed=7):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([np.sin(xx / 5.0) + 1.5, np.cos(yy / 4.0) * np.sin(xx / 9.0) + 1.5, (xx + yy) / 40.0 + 0.5],
                   axis=-1).astype(np.float32)
    img[10:20, 12:30] += 1.0                                   # an edge
    dx, dy = np.zeros_like(img), np.zeros_like(img)
    dx[:, :-1] = img[:, 1:] - img[:, :-1]
    dy[:-1] = img[1:] - img[:-1]
    dx += rng.normal(0, 0.02, img.shape).astype(np.float32)
    dy += rng.normal(0, 0.02, img.shape).astype(np.float32)
    for _ in range(12):                                        # gradient outliers (what L1 is for)
        dx[rng.integers(h), rng.integers(w)] += np.float32(rng.normal(0, 3.0))
        dy[rng.integers(h), rng.integers(w)] += np.float32(rng.normal(0, 3.0))
    tp = (img + rng.normal(0, 0.3, img.shape)).astype(np.float32)
    direct = (0.1 * np.abs(rng.normal(0, 1.0, img.shape))).astype(np.float32)
    return img, tp, dx, dy, direct


def run(case, tp, dx, dy, direct, backend="Naive"):
    kw = dict(CASES[case])
    p = pr.preset(kw.pop("preset"))
    if kw.pop("no_throughput", False):
        tp = None
    if kw.pop("no_direct", False):
        direct = None
    p.update(kw)
    return pr.solve(tp, dx, dy, direct, backend=backend, **p)


if __name__ == "__main__":
    img, tp, dx, dy, direct = images()
    out = {"truth": img, "throughput": tp, "dx": dx, "dy": dy, "direct": direct}
    for c in CASES:
        out["rec_" + c] = run(c, tp, dx, dy, direct)
        alt = run(c, tp, dx, dy, direct, backend="OpenMP")
        print(f"{c:20s} |rec - truth - direct| = {np.abs(out['rec_' + c] - img - (0 if 'no_direct' in c else direct)).mean():.4f}"
              f"   naive vs OpenMP backend: max abs diff {np.abs(alt - out['rec_' + c]).max():.2e}")
    path = os.path.join(ROOT, "tests", "golden", "poisson_small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
