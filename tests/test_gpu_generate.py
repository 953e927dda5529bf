"""Rows f-1 / f-2 of SURVEY.md §8 on the device: gvpm_generate_rays against the host generator (bit-exact, every field)
and gvpm_trace_photons against its CPU restatement oracle/gvpm_oracle_trace.cpp (bit-exact, every field; number of
light paths equal), then the gather on the device-made inputs against the oracle."""
import numpy as np
import pytest

import gvpm_b200 as g
import gvpm_testlib as H

pytestmark = pytest.mark.gpu


def _ctx(medium=None, config=None):
    from gvpm_b200.api import Context
    ctx = Context(0)
    ctx.set_medium(medium or g.make_medium())
    ctx.set_config(config or g.make_config(64, 64))
    ctx.set_occluders(g.synth_occluders())
    return ctx


def _same_soa(a, b, what):
    assert a.n == b.n, what
    for name, dt, _ in a.FIELDS:
        x, y = getattr(a, name), getattr(b, name)
        if dt == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        bad = int(np.count_nonzero(x != y))
        assert bad == 0, f"{what}: {name}: {bad} of {x.size} values differ"


@pytest.mark.parametrize("w,h,block,cam_dist,cover,rows", [
    (96, 64, 32, 1.5, 0.96, None),
    (100, 70, -32, 1.5, 0.96, None),        # partial blocks on both borders, Z-order
    (64, 48, -16, 1.5, 1.3, None),          # wide: border pixels miss the open face
    (80, 40, 8, -0.05, 0.45, None),         # sensor inside the medium
    (128, 96, -32, 1.5, 0.96, (32, 64)),    # a band of rows (image sharding)
])
def test_generated_rays_equal_the_host_generator(built, w, h, block, cam_dist, cover, rows):
    y0, y1 = rows or (0, h)
    seed = 0xC0FFEE + w
    host = g.synth_rays(w, h, seed=seed, block=block, y0=y0, y1=y1, cam_dist=cam_dist, cover=cover)
    ctx = _ctx()
    n = ctx.generate_rays(g.box_scene_default(), g.pinhole_camera(w, h, cam_dist, cover), seed, block=block, y0=y0, y1=y1)
    assert n == host.n
    _same_soa(ctx.download_rays(), host, f"rays {w}x{h} block {block}")
    ctx.close()


@pytest.mark.parametrize("kw", [
    {},
    {"seed": 99, "n": 33},
    {"phase": "hg", "hg_g": 0.6, "n": 40000},
    {"max_depth": 5, "min_depth": 2, "n": 20000},
    {"rr_depth": 3, "n": 20000},
    {"n": 300000, "seed": 5},               # several blocks of the path scan, more than one batch is possible
])
def test_traced_photons_equal_the_restatement(built, kw):
    from oracle import binding as ob
    n, seed = kw.get("n", 60000), kw.get("seed", 1234)
    med = g.make_medium(phase=kw.get("phase", "isotropic"), g=kw.get("hg_g", 0.0))
    scene = g.box_scene_default()
    args = dict(max_depth=kw.get("max_depth", 12), rr_depth=kw.get("rr_depth", 1), min_depth=kw.get("min_depth", 0))
    ref, ref_paths = ob.trace_photons(scene, med, n, seed, **args)
    ctx = _ctx(med)
    paths = ctx.trace_photons(scene, n, seed, **args)
    assert paths == ref_paths
    _same_soa(ctx.download_photons(n), ref, f"photons {kw}")
    # a second call (the photons-per-path estimate now sizes one batch) gives the same set
    assert ctx.trace_photons(scene, n, seed, **args) == ref_paths
    _same_soa(ctx.download_photons(n), ref, f"photons {kw} (second call)")
    ctx.close()


def test_other_scene_parameters(built):
    from oracle import binding as ob
    scene = g.box_scene_default()
    scene.n_rects = 2
    scene.rect[1].y, scene.rect[1].x0, scene.rect[1].x1, scene.rect[1].z0, scene.rect[1].z1 = 0.25, 0.1, 0.5, 0.1, 0.6
    for c in range(3):
        scene.rect[1].albedo[c] = 0.3 + 0.2 * c
        scene.face_albedo[2][c] = 0.2
    scene.light_x0, scene.light_x1, scene.light_power = 0.2, 0.4, 37.0
    med = g.make_medium(sigma_t=3.0, albedo=0.6)
    ref, ref_paths = ob.trace_photons(scene, med, 30000, 3)
    ctx = _ctx(med)
    assert ctx.trace_photons(scene, 30000, 3) == ref_paths
    _same_soa(ctx.download_photons(30000), ref, "other scene")
    ctx.close()


def test_gather_on_device_made_inputs(built):
    """trace + generate + build (perspective grid) + gather, nothing uploaded: against the oracle's brute force on the
    downloaded copies of the same inputs"""
    from oracle import binding as ob
    w, h, n = 64, 48, 60000
    med, cfg = g.make_medium(), g.make_config(w, h)
    scene, cam = g.box_scene_default(), g.pinhole_camera(w, h)
    ctx = _ctx(med, cfg)
    paths = ctx.trace_photons(scene, n, 21)
    ctx.generate_rays(scene, cam, 22)
    radius = g.bre_radius(2.0)
    ctx.build_points_for_rays(radius, want_kept=False)
    assert ctx.accel_kind() == "frustum"
    out, counts = ctx.gather_bre()
    ph, rays = ctx.download_photons(n), ctx.download_rays()
    ref = ob.bre_gather(ph, rays, med, cfg, g.synth_occluders(), radius, mode="brute")
    np.testing.assert_array_equal(counts, ref.counts)
    assert counts[:, 0].sum() > 2000 and paths > 0
    H.assert_radiance_close(out, ref.out, 1e-4, "gather on device-made inputs")
    ctx.close()


@pytest.mark.parametrize("build", ["frustum", "bvh"])
def test_direct_records_give_the_same_gather(built, build):
    """gvpm_trace_photons_direct writes the gather's 128-byte records in place: same neighbour counts and the same
    radiance as the staged photons packed by the build"""
    w, h, n = 64, 48, 50000
    med, cfg = g.make_medium(), g.make_config(w, h)
    scene, cam = g.box_scene_default(), g.pinhole_camera(w, h)
    radius = g.bre_radius(2.0)
    res = []
    for direct in (False, True):
        ctx = _ctx(med, cfg)
        ctx.trace_photons(scene, n, 77, direct=direct)
        ctx.generate_rays(scene, cam, 78)
        if build == "frustum":
            ctx.build_points_for_rays(radius, want_kept=False)
        else:
            ctx.build_points(radius)
        assert ctx.accel_kind() == build
        out, counts = ctx.gather_bre()
        out2, _ = ctx.gather_bre(counts=False)
        res.append((out.copy(), counts.copy(), out2.copy()))
        ctx.close()
    np.testing.assert_array_equal(res[0][1], res[1][1])
    # (a ray's partial sums are added with float atomics in whatever order the warps retire: equal up to rounding)
    H.assert_radiance_close(res[1][0], res[0][0], 1e-5, "gather on direct records")
    H.assert_radiance_close(res[1][2], res[0][0], 1e-5, "prefiltered gather on direct records")
    assert res[0][1][:, 0].sum() > 2000
