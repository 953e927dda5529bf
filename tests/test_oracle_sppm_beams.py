"""CPU checks of the restated sppm primal photon-beam gather (photonmapper/beams.h:29-223,
sppm.cpp:823-860): the four beam x beam techniques.  The reference has no fixture for them, so the
restatement is checked through identities between code paths that share no arithmetic beyond the pinned geometry
routines (cylinderIntersection, rayIntersectInternal1D: tests/test_oracle_ref_pin.py)."""
import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import records as R
from oracle import binding as ob


def _case(n_beams=1500, w=24, h=16, scale=4.0, seed=5, **kw):
    kw.setdefault("rng_seed", 99)
    kw.setdefault("path_set", False)
    kw.setdefault("max_depth", -1)
    c = H.make_case(n_photons=64, w=w, h=h, scale=scale, **kw)
    c.beams, c.n_beam_paths = R.synth_beams(n_beams, c.medium, seed=seed, threads=4)
    c.rays.off_valid[:] = 0
    return c


@pytest.mark.parametrize("tech,k1d", [("beam3d", False), ("beam1d", True)])
def test_matches_gvpm_primal_up_to_epsilon_transmittance(built, tech, k1d):
    """sppm's optimized-3D / 1-D branches and gvpm's BeamKernelRecord::eval are the same estimator written twice
    (beams.h:104-170,41-68 vs shift_volume_beams.h:169-283).  The only difference: sppm evaluates the camera
    transmittance from Epsilon (cameraRay.mint = Epsilon, beams.h:201), gvpm from 0 - a constant factor
    exp(sigma_t * Epsilon).  Index sets must be identical."""
    c = _case(beam_kernel_1d=k1d)
    c.config.max_depth = 0          # gvpm: unbounded
    gv = ob.beams_gather(c.beams, c.rays, c.medium, c.config, c.tri, c.radius, threads=4, neighbours=True)
    c.config.max_depth = -1         # sppm: unbounded
    sp = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=4, neighbours=True)
    assert sp.counts[:, 0].sum() > 500
    assert np.array_equal(sp.counts[:, 0], gv.counts[:, 0])
    assert np.array_equal(sp.idx & 0x7fffffff, gv.idx & 0x7fffffff)
    sigma_t = float(c.medium.sigma_s[0] + c.medium.sigma_a[0])
    want = gv.out.reshape(-1, 9, 3)[:, 0] * np.exp(sigma_t * c.config.epsilon)
    H.assert_radiance_close(sp.out, want, 2e-5, tech)


def test_three_3d_techniques_estimate_the_same_integral(built):
    """Naive, EGSR and optimized sample the same (beam point, camera point) kernel integral with different pdfs:
    averaged over many rays and seeds their totals agree statistically, and so do their per-ray means."""
    c = _case(n_beams=3000, w=32, h=24, scale=6.0)
    tot = {t: 0.0 for t in ("beam3d_naive", "beam3d_egsr", "beam3d")}
    per_ray = {t: 0.0 for t in tot}
    n_seeds = 6
    for s in range(n_seeds):
        c.config.rng_seed = 1000 + s
        for t in tot:
            r = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, t, threads=4)
            tot[t] += float(r.out.sum()) / n_seeds
            per_ray[t] = per_ray[t] + r.out.sum(axis=1) / n_seeds
    ref = tot["beam3d"]
    assert ref > 0
    for t in ("beam3d_naive", "beam3d_egsr"):
        assert abs(tot[t] - ref) < 0.04 * ref, (t, tot[t], ref)
        # per-ray: correlation of the two noisy images of the same signal
        cc = np.corrcoef(per_ray[t], per_ray["beam3d"])[0, 1]
        assert cc > 0.8, (t, cc)


def test_naive_visits_subbeams_and_split_matches(built):
    c = _case()
    t12, beam = ob.subbeams(c.beams)
    o, e = c.beams.view("origin"), c.beams.view("end")
    length = np.sqrt(((e - o).astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    n_sub = np.bincount(beam, minlength=c.beams.n)
    assert n_sub.min() >= 1 and abs(n_sub.mean() - 10) < 1.5            # average length / 10 (beams_accel.h:104)
    last = np.cumsum(n_sub) - 1
    np.testing.assert_allclose(t12[last, 1], length, rtol=1e-6)
    first = last - n_sub + 1
    assert not t12[first, 0].any()
    r = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, "beam3d_naive", threads=4, neighbours=True)
    # a long beam can be accepted through several of its sub-beams by the naive technique, never by the others
    ids = r.idx & 0x7fffffff
    dup = sum(len(ids[r.offsets[i]:r.offsets[i + 1]]) - len(set(ids[r.offsets[i]:r.offsets[i + 1]].tolist()))
              for i in range(c.rays.n))
    assert dup > 0
    o3 = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, "beam3d", threads=4, neighbours=True)
    ids = o3.idx & 0x7fffffff
    assert all(len(set(ids[o3.offsets[i]:o3.offsets[i + 1]].tolist())) == o3.offsets[i + 1] - o3.offsets[i]
               for i in range(c.rays.n))


@pytest.mark.parametrize("tech", sorted(ob.BEAM_TECHNIQUES))
def test_depth_window_and_eye_weight(built, tech):
    c = _case()
    full = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=4)
    assert np.array_equal(full.counts[:, 0], full.counts[:, 1])
    c.config.max_depth, c.config.min_depth = 4, 3
    win = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=4, neighbours=True)
    assert np.array_equal(win.counts[:, 0], full.counts[:, 0])
    assert 0 < win.counts[:, 1].sum() < full.counts[:, 1].sum()
    # sppm.cpp:853-854: beam.depth in [max(0, minDepth - camDepth), maxDepth - camDepth]
    depth = c.beams.depth[win.idx & 0x7fffffff].astype(np.int64)
    ray_of = np.repeat(np.arange(c.rays.n), np.diff(win.offsets).astype(np.int64))
    cam = c.rays.edge_id[ray_of].astype(np.int64)
    ok = (depth <= 4 - cam) & (depth >= np.maximum(0, 3 - cam))
    assert np.array_equal(ok, (win.idx >> 31).astype(bool))
    # Li * beam.weight is linear in the weight
    c.config.max_depth, c.config.min_depth = -1, 0
    c.rays.eye_contrib[:] *= 2.0
    twice = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=4)
    np.testing.assert_allclose(twice.out, 2.0 * full.out, rtol=2e-6)


@pytest.mark.parametrize("tech", sorted(ob.BEAM_TECHNIQUES))
def test_fp64_error_budget(built, tech):
    c = _case()
    a = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=4, neighbours=True)
    b = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=4, neighbours=True, double=True)
    same = np.array([np.array_equal(a.idx[a.offsets[i]:a.offsets[i + 1]], b.idx[b.offsets[i]:b.offsets[i + 1]])
                     for i in range(c.rays.n)])
    assert same.mean() > 0.95
    H.assert_radiance_close(a.out[same], b.out[same], 1e-4, tech)


def test_empty_inputs(built):
    c = _case(n_beams=8)
    empty = R.synth_beams(8, c.medium, seed=1, threads=1)[0].take(np.zeros(0, dtype=np.int64))
    r = ob.sppm_beams_gather(empty, c.rays, c.medium, c.config, c.radius, "beam3d", threads=2)
    assert not r.out.any() and not r.counts.any()


def test_committed_regression_vectors(built):
    """tests/golden/sppm_beams_small.npz (tests/golden/make_golden.py; self-generated, guards the restatement against
    drift): accepted-beam lists and counts bit-exact, radiance to 1e-6."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sppm_beams_small.npz"))
    c = H.make_case(n_photons=64, w=int(z["w"]), h=int(z["h"]), scale=float(z["scale"]), rng_seed=99, path_set=False,
                    max_depth=7, min_depth=2)
    c.beams, _ = R.synth_beams(int(z["n_beams"]), c.medium, seed=int(z["seed"]), threads=2)
    for tech in sorted(ob.BEAM_TECHNIQUES):
        r = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, threads=3, neighbours=True)
        np.testing.assert_array_equal(r.counts, z["counts_" + tech])
        np.testing.assert_array_equal(r.offsets, z["offsets_" + tech])
        np.testing.assert_array_equal(r.idx, z["idx_" + tech])
        H.assert_radiance_close(r.out, z["out_" + tech], 1e-6, tech)
        assert r.counts[:, 1].sum() > 50
