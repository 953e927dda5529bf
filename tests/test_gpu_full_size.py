"""BASELINE.json configs[1..3] and the north-star's G-Beams-3D target at FULL size, through size-independent properties
(the oracle cannot run them whole): a sample of rays spread over the image is gathered by the CPU oracle against ALL
primitives and must match the GPU gather of the whole image - counts bit-exact, radiance within 1e-4 per ray."""
import os

import numpy as np
import pytest

import gvpm_b200 as g
import gvpm_testlib as H
from gvpm_b200 import records as R

pytestmark = pytest.mark.gpu


def _ctx(medium, config, tri=None):
    from gvpm_b200.api import Context
    ctx = Context(0)
    ctx.set_medium(medium)
    ctx.set_config(config)
    if tri is not None:
        ctx.set_occluders(tri)
    return ctx


def test_cfg2_vpm_full_size(built):
    """512x512 pixels x 40 distance samples against 1 M photons (scale 1.0): sampled rays vs the oracle's kd-tree range
    queries; MVol, per-sample counts and radiance."""
    from oracle import binding as ob
    w = h = 512
    nb = 40
    med = g.make_medium()
    ph, _ = g.synth_photons(1_000_000, med, seed=0xC0FFEE + 2, threads=os.cpu_count() or 8)
    rays = g.synth_rays(w, h, seed=0xC0FFEE + 3, block=-32)
    cfg, tri, radius = g.make_config(w, h), g.synth_occluders(), g.bre_radius(1.0)
    rad = np.full(rays.n, radius, dtype=np.float32)
    samples = g.synth_vpm_samples(rays, med, rad, nb_camera_samples=nb, seed=0xC0FFEE + 4)
    ctx = _ctx(med, cfg, tri)
    ctx.upload_photons(ph)
    ctx.build_points(radius)
    ctx.upload_rays(rays)
    ctx.upload_vpm_samples(samples)
    out, mvol, sc = ctx.gather_vpm(nb)
    assert sc[:, 0].sum() > 5_000_000
    sel = np.arange(0, rays.n, 257)
    sidx = np.nonzero(np.isin(samples.ray, sel))[0]
    sub_s = samples.take(sidx)
    sub_s.ray[:] = np.searchsorted(sel, sub_s.ray).astype(np.uint32)
    ref = ob.vpm_gather(ph, rays.take(sel), sub_s, med, cfg, tri, nb, mode="kdtree")
    np.testing.assert_array_equal(sc[sidx], ref.sample_counts)
    np.testing.assert_array_equal(mvol[sel].astype(np.float32), ref.mvol)
    H.assert_radiance_close(out[sel], ref.out, 1e-4, "cfg2 sampled rays vs oracle")
    ctx.close()


@pytest.mark.parametrize("name,w,h,n_beams,stride", [("cfg3", 1280, 720, 500_000, 1801), ("beams1080", 1920, 1080, 1_000_000, 6007)])
def test_beams_full_size(built, name, w, h, n_beams, stride):
    """G-Beams 3D at BASELINE size (scale 0.1): sampled rays vs the oracle's brute force over all beams."""
    from oracle import binding as ob
    med = g.make_medium()
    beams, _ = R.synth_beams(n_beams, med, seed=0xC0FFEE + 3, threads=os.cpu_count() or 8)
    rays = g.synth_rays(w, h, seed=0xC0FFEE + 4, block=-32)
    cfg, tri, radius = g.make_config(w, h, rng_seed=99), g.synth_occluders(), g.bre_radius(0.1)
    ctx = _ctx(med, cfg, tri)
    ctx.upload_beams(beams)
    ctx.build_beams(radius)
    ctx.upload_rays(rays)
    out, counts = ctx.gather_beams()
    out_fast, _ = ctx.gather_beams(counts=False)
    assert counts[:, 0].sum() > 10 * rays.n
    sel = np.arange(0, rays.n, stride)
    ref = ob.beams_gather(beams, rays.take(sel), med, cfg, tri, radius)
    np.testing.assert_array_equal(counts[sel], ref.counts)
    H.assert_radiance_close(out[sel], ref.out, 1e-4, f"{name} sampled rays vs oracle")
    H.assert_radiance_close(out_fast[sel], ref.out, 1e-4, f"{name} sampled rays vs oracle (prefiltered)")
    # the sub-beam cuts made on the device are the reference's (beams_accel.h:98-124)
    t12, bi = ob.subbeams(beams)
    assert ctx.n_subbeams() == len(bi)
    ctx.close()


def test_cfg4_planes_full_size(built):
    """1280x720 rays against 200 k planes of the LASER-style sheet, sensor inside the medium."""
    from oracle import binding as ob
    w, h = 1280, 720
    med = g.make_medium()
    beams, _ = R.synth_beams(200_000, med, seed=0xC0FFEE + 4, threads=os.cpu_count() or 8)
    planes = R.synth_planes(beams, med, seed=0xC0FFEE + 11)
    o = planes.view("origin")
    o[:, 0] = 0.5 + (o[:, 0] - 0.5) * 0.02
    planes.length1[:] *= 0.05
    rays = g.synth_rays(w, h, seed=0xC0FFEE + 5, block=-32, cam_dist=-0.05, cover=0.45)
    cfg = g.make_config(w, h)
    ctx = _ctx(med, cfg)
    ctx.upload_planes(planes)
    ctx.build_planes()
    ctx.upload_rays(rays)
    out, counts = ctx.gather_planes()
    assert counts[:, 0].sum() > rays.n
    sel = np.arange(0, rays.n, 1801)
    ref = ob.planes_gather(planes, rays.take(sel), med, cfg)
    np.testing.assert_array_equal(counts[sel], ref.counts)
    H.assert_radiance_close(out[sel], ref.out, 1e-4, "cfg4 sampled rays vs oracle")
    ctx.close()
