"""gvpm_build_points_for_rays (ray-region pruning): the hierarchy over the photons the uploaded rays can
reach must give the same gathers as the full hierarchy - bit-exact neighbour index sets against the CPU oracle,
radiance within 1e-4 - while keeping only the part of the photon set the rays cross."""
import numpy as np
import pytest

import gvpm_b200 as g
import gvpm_testlib as H
from gvpm_b200 import shard

pytestmark = pytest.mark.gpu


def _oracle(case, rays):
    from oracle import binding as ob
    return ob.bre_gather(case.photons, rays, case.medium, case.config, case.tri, case.radius, mode="brute",
                         neighbours=True)


def _check_pruned(case, rays, what, expect_fraction=None):
    ref = _oracle(case, rays)
    ctx = H.gpu_context(case)
    ctx.upload_rays(rays)
    kept = ctx.build_points_for_rays(case.radius)
    out, counts = ctx.gather_bre()
    offsets, idx = ctx.dump_neighbours_bre()
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    # the neighbours themselves are of course among the kept photons
    assert kept >= len(np.unique(ref.idx & 0x7fffffff))
    assert kept <= case.photons.n
    if expect_fraction is not None:
        assert kept <= expect_fraction * case.photons.n, (kept, case.photons.n)
    ctx.close()
    return kept, ref


def test_full_image_pruned_equals_oracle(built):
    case = H.make_case(n_photons=60000, w=64, h=48, scale=2.0)
    kept, ref = _check_pruned(case, case.rays, "pruned, all rays")
    assert ref.counts[:, 0].sum() > 3000


@pytest.mark.parametrize("world,cycles", [(2, 2), (4, 1), (8, 2)])
def test_band_shards_keep_a_fraction_and_match(built, world, cycles):
    """Every rank of a column-band partition: same per-ray results as the oracle, a fraction of the photons."""
    case = H.make_case(n_photons=120000, w=256, h=64, scale=1.0)
    r = case.rays
    total_kept = 0
    for rank in (0, world - 1):
        idx = shard.band_indices(r.px, r.py, case.w, case.h, world, rank, cycles)
        assert len(idx) > 0
        kept, _ = _check_pruned(case, r.take(idx), f"band shard {rank}/{world}")
        total_kept += kept
    # two ranks' wedges + margins: well below two full photon sets
    assert total_kept < 2 * case.photons.n * min(1.0, 2.2 / world + 0.25)


def test_incoherent_ray_order_and_tiny_sets(built):
    """Shuffled rays (no warp is a pixel patch: the per-ray marking path), a single ray, and rays missing every
    photon."""
    case = H.make_case(n_photons=40000, w=48, h=32, scale=2.0)
    rng = np.random.default_rng(3)
    perm = rng.permutation(case.rays.n)
    _check_pruned(case, case.rays.take(perm), "shuffled")
    one = case.rays.take(np.array([case.rays.n // 2 + 7]))
    kept, _ = _check_pruned(case, one, "one ray", expect_fraction=0.2)
    assert kept > 0
    # rays translated far outside the photon cloud
    far = case.rays.take(np.arange(64))
    far.o[:] += 50.0
    far.off_o[:] += 50.0
    ctx = H.gpu_context(case)
    ctx.upload_rays(far)
    assert ctx.build_points_for_rays(case.radius) == 0
    out, counts = ctx.gather_bre()
    assert not out.any() and not counts.any()
    ctx.close()


def test_stale_hierarchy_is_refused_and_other_gathers_work(built):
    from gvpm_b200.api import GvpmError
    case = H.make_case(n_photons=50000, w=48, h=32, scale=2.0, sppm_primal=False)
    ctx = H.gpu_context(case)
    half = case.rays.take(np.arange(case.rays.n // 2))
    ctx.upload_rays(half)
    ctx.build_points_for_rays(case.radius)
    a, _ = ctx.gather_bre()
    ctx.upload_rays(case.rays)                      # other rays: the pruned hierarchy no longer covers them
    with pytest.raises(GvpmError):
        ctx.gather_bre()
    ctx.build_points(case.radius)                   # a full build serves any ray set
    b, _ = ctx.gather_bre()
    H.assert_radiance_close(a, b[:half.n], 1e-5, "pruned vs full")
    # G-VPM through a pruned hierarchy (the distance samples lie on the uploaded rays' segments)
    from oracle import binding as ob
    rad = np.full(case.rays.n, case.radius, dtype=np.float32)
    samples = g.synth_vpm_samples(case.rays, case.medium, rad, nb_camera_samples=6, seed=5)
    ref = ob.vpm_gather(case.photons, case.rays, samples, case.medium, case.config, case.tri, 6, mode="brute")
    ctx.build_points_for_rays(case.radius)
    ctx.upload_vpm_samples(samples)
    out, mvol, sc = ctx.gather_vpm(6)
    np.testing.assert_array_equal(sc, ref.sample_counts)
    H.assert_radiance_close(out, ref.out, 1e-4, "vpm pruned")
    ctx.close()


def test_pruned_needs_rays(built):
    from gvpm_b200.api import Context, GvpmError
    case = H.make_case(n_photons=2000, w=16, h=16)
    ctx = Context(0)
    ctx.set_medium(case.medium)
    ctx.set_config(case.config)
    ctx.set_occluders(case.tri)
    ctx.upload_photons(case.photons)
    with pytest.raises(GvpmError):
        ctx.build_points_for_rays(case.radius)
    ctx.close()


def test_cfg5_full_size_properties(built):
    """BASELINE.json configs[4] at full size (1920x1080 rays, 10 M photons, scale 0.1), through size-independent
    properties: (1) every 997th ray matches the oracle's reference-shaped kd-tree gather (counts bit-exact up to the
    reference tree's Epsilon sliver, radiance 1e-4); (2) two band shards gathered through pruned hierarchies return
    exactly the rows of the full gather's neighbour counts and the same radiance; (3) totals: the geometric
    neighbour count is the sum over the shards."""
    import os
    from oracle import binding as ob
    w, h, n = 1920, 1080, 10_000_000
    med = g.make_medium()
    ph, _ = g.synth_photons(n, med, seed=0xC0FFEE + 5, threads=os.cpu_count() or 8)
    rays = g.synth_rays(w, h, seed=0xC0FFEE + 6, block=-32)
    case = H.Case()
    case.medium, case.photons, case.rays, case.tri = med, ph, rays, g.synth_occluders()
    case.config, case.radius, case.w, case.h = g.make_config(w, h), g.bre_radius(0.1), w, h
    ctx = H.gpu_context(case)
    out, counts = ctx.gather_bre()
    assert counts[:, 0].sum() > 30_000_000
    # (1)
    sel = np.arange(0, rays.n, 997)
    ref = ob.bre_gather(ph, rays.take(sel), med, case.config, case.tri, case.radius, mode="kdtree")
    same = (ref.counts == counts[sel]).all(axis=1)
    assert same.mean() > 0.999
    H.assert_radiance_close(out[sel][same], ref.out[same], 1e-4, "cfg5 sampled rays vs oracle kd-tree")
    # (2) + (3)
    total = 0
    for rank in (1, 6):
        idx = shard.band_indices(rays.px, rays.py, w, h, 8, rank, 2)
        ctx.upload_rays(rays.take(idx))
        kept = ctx.build_points_for_rays(case.radius)
        assert kept < 0.3 * n
        o, c = ctx.gather_bre()
        np.testing.assert_array_equal(c, counts[idx])
        H.assert_radiance_close(o, out[idx], 1e-5, f"cfg5 band shard {rank}")
        total += int(c[:, 0].sum())
    assert 0 < total < counts[:, 0].sum()
    ctx.close()
