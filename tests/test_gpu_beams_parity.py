"""GPU parity of the G-Beams 3D gather against the CPU oracle, through the C ABI.
Bar: per-ray counts and beam index sets bit-exact, radiance within 1e-4 relative (fp32).  The two uniforms per
(ray, beam) come from the counter-based hash both sides share."""
import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import records as R

pytestmark = pytest.mark.gpu


def _case(n_beams=6000, w=40, h=24, scale=3.0, seed=5, **kw):
    kw.setdefault("rng_seed", 1234)
    c = H.make_case(n_photons=64, w=w, h=h, scale=scale, **kw)
    c.beams, c.n_beam_paths = R.synth_beams(n_beams, c.medium, seed=seed, threads=4)
    return c


def _check(c, what):
    from oracle import binding as ob
    from gvpm_b200.api import Context
    ref = ob.beams_gather(c.beams, c.rays, c.medium, c.config, c.tri, c.radius, neighbours=True)
    ctx = Context(0)
    ctx.set_medium(c.medium)
    ctx.set_config(c.config)
    ctx.set_occluders(c.tri)
    ctx.upload_beams(c.beams)
    ctx.build_beams(c.radius)
    ctx.upload_rays(c.rays)
    out, counts = ctx.gather_beams()
    out_fast, _ = ctx.gather_beams(counts=False)  # filters applied in the traversal
    offsets, idx = ctx.dump_neighbours_beams()
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    H.assert_radiance_close(out_fast, ref.out, 1e-4, what + " (prefiltered)")
    ctx.close()
    return ref


@pytest.mark.parametrize("kw", [
    {},
    {"use_shift_null": False},
    {"use_mis": False, "max_depth": 6},
    {"power_heuristic": True, "path_set": False},
    {"lighting_mode": 1 << 4},
    {"long_beams": True},
])
def test_beams3d_matches_oracle(built, kw):
    c = _case(**kw)
    ref = _check(c, f"beams {kw}")
    assert ref.counts[:, 0].sum() > 3000


@pytest.mark.parametrize("kw", [
    {},
    {"use_mis": False, "max_depth": 6},
    {"power_heuristic": True, "path_set": False},
    {"lighting_mode": 1 << 2, "use_shift_null": False},
    {"long_beams": True},
])
def test_beams1d_matches_oracle(built, kw):
    """beam1d kernel (EBeamBeam1D, newShiftBeam): line-line closest approach, 1/(2r)/sin(theta), getShiftPos1D."""
    c = _case(beam_kernel_1d=True, **kw)
    ref = _check(c, f"beams1d {kw}")
    assert ref.counts[:, 0].sum() > 3000


def test_beams1d_hg_small_radius(built):
    _check(_case(scale=1.0, n_beams=20000, phase="hg", hg_g=0.5, beam_kernel_1d=True), "beams1d hg")


def test_beams3d_hg_small_radius(built):
    _check(_case(scale=1.0, n_beams=20000, phase="hg", hg_g=0.5), "beams hg")


def test_beams_single_and_ragged(built):
    for n in (1, 3, 33):
        _check(_case(n_beams=n, scale=8.0), f"beams n={n}")
        _check(_case(n_beams=n, scale=8.0, beam_kernel_1d=True), f"beams1d n={n}")
