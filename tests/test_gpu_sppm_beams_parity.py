"""GPU parity of the sppm primal photon-beam entry (gvpm_gather_sppm_beams; beams.h:29-223,
sppm.cpp:823-860) against the CPU oracle, through the C ABI, for the four beam x beam techniques.  Bar: per-ray
counts and accepted-beam lists bit-exact, radiance within 1e-4 relative (fp32)."""
import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import records as R

pytestmark = pytest.mark.gpu
TECHNIQUES = ["beam1d", "beam3d_naive", "beam3d_egsr", "beam3d"]


def _case(n_beams=6000, w=40, h=24, scale=3.0, seed=5, **kw):
    kw.setdefault("rng_seed", 4321)
    kw.setdefault("max_depth", -1)
    c = H.make_case(n_photons=64, w=w, h=h, scale=scale, **kw)
    c.beams, c.n_beam_paths = R.synth_beams(n_beams, c.medium, seed=seed, threads=4)
    return c


def _ctx(c):
    from gvpm_b200.api import Context
    ctx = Context(0)
    ctx.set_medium(c.medium)
    ctx.set_config(c.config)
    ctx.set_occluders(c.tri)
    ctx.upload_beams(c.beams)
    ctx.build_beams(c.radius)
    ctx.upload_rays(c.rays)
    return ctx


def _check(c, tech, what, ctx=None):
    from oracle import binding as ob
    ref = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, neighbours=True)
    own = ctx is None
    ctx = ctx or _ctx(c)
    out, counts = ctx.gather_sppm_beams(tech)
    out_fast, _ = ctx.gather_sppm_beams(tech, counts=False)
    offsets, idx = ctx.dump_neighbours_sppm_beams(tech)
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    H.assert_radiance_close(out_fast, ref.out, 1e-4, what + " (no counts)")
    if own:
        ctx.close()
    return ref


@pytest.mark.parametrize("tech", TECHNIQUES)
@pytest.mark.parametrize("kw", [{}, {"max_depth": 5, "min_depth": 3}, {"long_beams": True}])
def test_sppm_beams_match_oracle(built, tech, kw):
    c = _case(**kw)
    ref = _check(c, tech, f"sppm beams {tech} {kw}")
    assert ref.counts[:, 0].sum() > 2000
    if kw.get("max_depth"):
        assert ref.counts[:, 1].sum() < ref.counts[:, 0].sum()


def test_sppm_beams_hg_small_radius_one_context(built):
    """All four techniques from one context / one hierarchy (HG phase, many more beams than pixels)."""
    c = _case(n_beams=60000, w=32, h=24, scale=0.7, phase="hg", hg_g=0.5)
    ctx = _ctx(c)
    for tech in TECHNIQUES:
        _check(c, tech, f"sppm beams hg {tech}", ctx)
    ctx.close()


def test_sppm_beams_subbeam_split_matches_reference_rule(built):
    """The sub-beam table of gvpm_build_beams is the SubBeamBVH constructor's (beams_accel.h:98-124): the naive
    technique, whose estimate depends on the cuts, agrees with the oracle only if it is."""
    c = _case(n_beams=500, scale=8.0)
    ref = _check(c, "beam3d_naive", "naive wide kernel")
    assert ref.counts[:, 0].max() > 5


def test_sppm_beams_errors_and_empty(built):
    from gvpm_b200.api import Context, GvpmError
    c = _case(n_beams=16)
    ctx = _ctx(c)
    with pytest.raises(GvpmError):
        ctx.gather_sppm_beams(7)
    ctx.upload_beams(c.beams.take(np.zeros(0, dtype=np.int64)))
    ctx.build_beams(c.radius)
    out, counts = ctx.gather_sppm_beams("beam3d")
    assert not out.any() and not counts.any()
    ctx.close()
