"""CPU checks of the restated on-device photon tracer (oracle/gvpm_oracle_trace.cpp, the checker of gvpm_trace_photons):
its polynomial log / exp / sin / cos against libm, determinism, bookkeeping invariants of the records,
and statistical agreement with the libm-based synthetic generator the other tests use (same walk, other arithmetic)."""
import numpy as np

import gvpm_b200 as g
from oracle import binding as ob


def test_polynomial_routines_against_libm(built):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(1e-7, 1.0, 20000), np.exp(rng.uniform(np.log(1e-7), 0, 20000))]).astype(np.float32)
    lg, ex, sn, cs = ob.pm_functions(x)
    assert np.abs(lg - np.log(x.astype(np.float64))).max() < 4e-7 * 16            # |log| <= 16: a few ulp
    assert (np.abs(ex - np.exp(-x.astype(np.float64))) / np.exp(-x.astype(np.float64))).max() < 4e-7
    big = rng.uniform(0, 80, 20000).astype(np.float32)
    _, exb, _, _ = ob.pm_functions(big)
    ref = np.exp(-big.astype(np.float64))
    assert (np.abs(exb - ref) / ref).max() < 1e-6
    frac = (x - np.floor(x)).astype(np.float64)
    assert np.abs(sn - np.sin(2 * np.pi * frac)).max() < 4e-7 and np.abs(cs - np.cos(2 * np.pi * frac)).max() < 4e-7


def test_trace_is_deterministic_and_well_formed(built):
    scene, med = g.box_scene_default(), g.make_medium()
    a, pa = ob.trace_photons(scene, med, 30000, seed=7)
    b, pb_ = ob.trace_photons(scene, med, 30000, seed=7)
    for name, _, _ in a.FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert pa == pb_ > 0
    pos = a.view("pos")
    assert (pos > 0).all() and (pos < 1).all()                                    # photons live inside the medium
    assert (a.depth >= 1).all() and (a.depth <= 11).all() and (a.parent_type <= 2).all()
    assert (np.diff(a.path_id.astype(np.int64)) >= 0).all() and a.path_id[0] == 0  # contributing paths numbered in order
    first = a.depth == 1                                                          # parent = the emitter sample
    assert (a.parent_type[first] == 0).all() and np.allclose(a.view("parent_pos")[first][:, 1], 0.999)
    assert (a.view("pred_pos")[first] == 1).all()                                 # (1, 1, 1) when there is no vertex c - 2
    assert (a.parent_pdf > 0).all() and (a.edge_pdf > 0).all() and (a.rr_weight >= 1).all()
    other = ob.trace_photons(scene, med, 30000, seed=8)[0]
    assert not np.array_equal(other.pos, a.pos)


def test_trace_statistics_match_the_libm_generator(built):
    """same random walk, polynomial instead of libm transcendental functions: distributions agree"""
    scene, med = g.box_scene_default(), g.make_medium()
    n = 200000
    a, pa = ob.trace_photons(scene, med, n, seed=11)
    b, pb_ = g.synth_photons(n, med, seed=11, threads=4)
    assert abs(pa - pb_) / pb_ < 0.02                                             # photons per light path
    assert abs(a.depth.mean() - b.depth.mean()) < 0.05
    for t in range(3):
        assert abs((a.parent_type == t).mean() - (b.parent_type == t).mean()) < 0.01
    assert np.abs(a.view("pos").mean(0) - b.view("pos").mean(0)).max() < 0.01
    assert abs(np.log(a.view("flux")[:, 0]).mean() - np.log(b.view("flux")[:, 0]).mean()) < 0.02
