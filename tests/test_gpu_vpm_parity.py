"""GPU parity of the G-VPM point gather against the CPU oracle, through the
C ABI.  Bar: per-sample neighbour counts and index sets bit-exact, MVol exact, radiance within 1e-4
relative (fp32)."""
import numpy as np
import pytest

import gvpm_b200 as g
import gvpm_testlib as H

pytestmark = pytest.mark.gpu


def _case(scale=3.0, nb=8, n_photons=60000, w=40, h=24, vary_radius=True, **kw):
    c = H.make_case(n_photons=n_photons, w=w, h=h, scale=scale, **kw)
    rad = np.full(c.rays.n, c.radius, dtype=np.float32)
    if vary_radius:  # per-pixel SPPM radii (gp.scaleVol, gvpm.cpp:1131,1191-1195)
        rad *= np.random.default_rng(2).uniform(0.4, 1.0, c.rays.n).astype(np.float32)
    c.samples = g.synth_vpm_samples(c.rays, c.medium, rad, nb_camera_samples=nb, seed=99)
    c.nb = nb
    return c


def _check(c, what):
    from oracle import binding as ob
    ref = ob.vpm_gather(c.photons, c.rays, c.samples, c.medium, c.config, c.tri, c.nb, mode="brute", neighbours=True)
    ctx = H.gpu_context(c)
    ctx.upload_vpm_samples(c.samples)
    out, mvol, sc = ctx.gather_vpm(c.nb)
    offsets, idx = ctx.dump_neighbours_vpm(c.nb)
    np.testing.assert_array_equal(sc, ref.sample_counts)
    np.testing.assert_array_equal(mvol.astype(np.float32), ref.mvol)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    ctx.close()
    return ref


@pytest.mark.parametrize("kw", [
    {},
    {"use_shift_null": False},
    {"use_mis": False, "max_depth": 5},
    {"power_heuristic": True, "shadow_maxt_scale": 0.999},
    {"lighting_mode": 1 << 2},
])
def test_vpm_matches_oracle(built, kw):
    c = _case(**kw)
    ref = _check(c, f"vpm {kw}")
    assert ref.sample_counts[:, 0].sum() > 2000


def test_vpm_hg_and_uniform_radius(built):
    _check(_case(phase="hg", hg_g=0.4, vary_radius=False), "vpm hg")


def test_vpm_radius_larger_than_build_radius_is_rejected(built):
    from gvpm_b200.api import GvpmError
    c = _case(vary_radius=False)
    c.samples.radius[:] *= np.float32(1.5)
    ctx = H.gpu_context(c)
    ctx.upload_vpm_samples(c.samples)
    with pytest.raises(GvpmError, match="smaller than a sample radius"):
        ctx.gather_vpm(c.nb)
    ctx.close()


def test_sppm_shaped_vpm_pass(built):
    """sppm's point-photon volume pass (sppm.cpp:1095-1112 -> PhotonMap::estimateVolumeRadiance, photonmap.cpp:324-330): a
    range query per distance sample with the depth bound maxDepth - beam.depth, scaled by transmittance / (pdfSuccess *
    selBeam) - the gvpm entry with no valid offset path and edge_id = beam.depth.  MVol, index sets and the primal
    match the oracle; the primal equals the gradient-domain run's."""
    from oracle import binding as ob
    c = _case(max_depth=6)
    ref_grad = ob.vpm_gather(c.photons, c.rays, c.samples, c.medium, c.config, c.tri, c.nb, mode="brute")
    c.rays = c.rays.copy()
    c.rays.off_valid[:] = 0
    ref = _check(c, "sppm-shaped vpm")
    o = ref.out.reshape(-1, 9, 3)
    H.assert_radiance_close(o[:, 0], ref_grad.out.reshape(-1, 9, 3)[:, 0], 1e-6, "primal does not depend on the offsets")
    assert not o[:, 1:5].any()
    assert ref.sample_counts[:, 0].sum() > ref.sample_counts[:, 1].sum() > 1000     # the depth bound filters some photons
