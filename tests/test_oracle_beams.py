"""CPU self-consistency checks of the G-Beams oracle restatement (beam3d and beam1d kernels): the parts of it the
reference pin cannot reach (tests/test_oracle_ref_pin.py pins cylinderIntersection and rayIntersectInternal1D
bit-exactly; the functor arithmetic is checked here through identities that hold whatever the inputs)."""
import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import records as R
from oracle import binding as ob


def _case(n_beams=1500, w=24, h=16, scale=4.0, seed=5, **kw):
    kw.setdefault("rng_seed", 99)
    c = H.make_case(n_photons=64, w=w, h=h, scale=scale, **kw)
    c.beams, c.n_beam_paths = R.synth_beams(n_beams, c.medium, seed=seed, threads=4)
    return c


def _identical_offsets(c):
    r = c.rays
    r.off_o[:] = np.repeat(r.view("o"), 4, axis=0).reshape(-1)
    r.off_d[:] = np.repeat(r.view("d"), 4, axis=0).reshape(-1)
    r.off_len[:] = np.repeat(r.edge_len, 4)
    r.off_eye[:] = np.repeat(r.view("eye_contrib"), 4, axis=0).reshape(-1)
    r.off_sensor[:] = 1.0
    r.off_valid[:] = 1


@pytest.mark.parametrize("k1d", [False, True])
def test_identical_offset_ray_reproduces_base(built, k1d):
    """Offset ray == base ray.  beam3d: the null shift re-evaluates the same kernel record.  beam1d: getShiftPos1D
    must land on the base intersection point beam(v), so the diffuse reconnection rebuilds the base path: the
    shifted contribution equals the base one and the balance-heuristic weight is 1/(1+1)."""
    c = _case(beam_kernel_1d=k1d, perturb=False)
    _identical_offsets(c)
    c.tri = np.zeros((0, 9), np.float32)     # no occluders: the reconnection shadow ray must not interfere
    res = ob.beams_gather(c.beams, c.rays, c.medium, c.config, c.tri, c.radius, threads=4)
    out = res.out.reshape(-1, 9, 3)
    primal = out[:, 0]
    assert res.counts[:, 1].sum() > 500 and primal.sum() > 0
    r = c.rays
    for k in range(4):
        w = np.full(r.n, 0.5)
        if k == 1:
            w[r.px == c.w - 1] = 1.0
        if k == 2:
            w[r.py == c.h - 1] = 1.0
        # beam1d: phi = pi/2 - asin(u / |localA.y|) (shift_volume_beams.cpp:68) is ill-conditioned in fp32 when the beam
        # origin lies close to the camera ray (u / |localA.y| -> 1), which moves the reconnected point by ~1e-3 r.
        # beam3d: the offset segment runs to edge_len while the base one stops at edge_len - Epsilon (gvpm.cpp:931-936
        # vs shift_volume_beams.cpp:238-240), so a cylinder clipped by the ray end has a slightly different pdf on
        # the few rays where that happens; everything else must match tightly
        want = primal * w[:, None]
        tol = np.maximum(np.abs(want) * 2e-3, primal.max() * 2e-5)
        for got in (out[:, 5 + k], out[:, 1 + k]):
            bad = np.abs(got - want) > tol
            assert bad.mean() < 0.03, (k, bad.sum())
            assert np.abs(got - want).max() <= np.abs(want).max() * 0.03


@pytest.mark.parametrize("k1d", [False, True])
def test_invalid_offsets_keep_full_weight(built, k1d):
    c = _case(beam_kernel_1d=k1d)
    c.rays.off_valid[:] = 0
    res = ob.beams_gather(c.beams, c.rays, c.medium, c.config, c.tri, c.radius, threads=4)
    out = res.out.reshape(-1, 9, 3)
    assert out[:, 0].sum() > 0 and not out[:, 1:5].any()
    for k in range(4):
        np.testing.assert_allclose(out[:, 5 + k], out[:, 0], rtol=1e-6)


@pytest.mark.parametrize("k1d", [False, True])
def test_fp64_error_budget(built, k1d):
    """fp32 vs fp64 instantiation on rays whose index sets agree (a predicate can flip in the last bit)."""
    c = _case(beam_kernel_1d=k1d)
    a = ob.beams_gather(c.beams, c.rays, c.medium, c.config, c.tri, c.radius, threads=4, neighbours=True)
    b = ob.beams_gather(c.beams, c.rays, c.medium, c.config, c.tri, c.radius, threads=4, neighbours=True, double=True)
    same = np.array([np.array_equal(a.idx[a.offsets[i]:a.offsets[i + 1]], b.idx[b.offsets[i]:b.offsets[i + 1]])
                     for i in range(c.rays.n)])
    assert same.mean() > 0.95
    err = H.rel_err(a.out[same], b.out[same])
    assert err.max() < 2e-4, err.max()


def test_beam1d_kernel_weight_and_sets(built):
    """beam1d accepts (ray, beam) iff the lines pass within r inside both segments: a subset relation ties it to
    beam3d's cylinder test at the same radius for pairs well inside the beam, and the primal is
    sum flux*T*T*sigma_s*phase / pdfFail / sin(theta) / (2r): positive and finite."""
    c1 = _case(beam_kernel_1d=True, path_set=False)
    r1 = ob.beams_gather(c1.beams, c1.rays, c1.medium, c1.config, c1.tri, c1.radius, threads=4, neighbours=True)
    assert np.isfinite(r1.out).all() and (r1.out >= 0).all()
    assert r1.counts[:, 0].sum() > 1000
    # every contributing pair is geometric, contributing <= geometric
    assert (r1.counts[:, 1] <= r1.counts[:, 0]).all()
    # halving the radius halves (roughly) the number of accepted pairs and keeps them a subset
    r2 = ob.beams_gather(c1.beams, c1.rays, c1.medium, c1.config, c1.tri, c1.radius * 0.5, threads=4, neighbours=True)
    for i in range(0, c1.rays.n, 7):
        s_big = set((r1.idx[r1.offsets[i]:r1.offsets[i + 1]] & 0x7FFFFFFF).tolist())
        s_small = set((r2.idx[r2.offsets[i]:r2.offsets[i + 1]] & 0x7FFFFFFF).tolist())
        assert s_small <= s_big
    ratio = r2.counts[:, 0].sum() / r1.counts[:, 0].sum()
    assert 0.4 < ratio < 0.6, ratio
