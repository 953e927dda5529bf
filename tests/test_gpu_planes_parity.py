"""GPU parity of the G-Planes 0D gather against the CPU oracle, through the C ABI.
Bar: per-ray counts and plane index sets bit-exact, radiance within 1e-4 relative (fp32)."""
import numpy as np
import pytest

import gvpm_testlib as H

pytestmark = pytest.mark.gpu


def _check(c, what, min_hits=0):
    from oracle import binding as ob
    from gvpm_b200.api import Context
    ref = ob.planes_gather(c.planes, c.rays, c.medium, c.config, neighbours=True)
    ctx = Context(0)
    ctx.set_medium(c.medium)
    ctx.set_config(c.config)
    ctx.upload_planes(c.planes)
    ctx.build_planes()
    ctx.upload_rays(c.rays)
    out, counts = ctx.gather_planes()
    out2, none = ctx.gather_planes(counts=False)
    offsets, idx = ctx.dump_neighbours_planes()
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    np.testing.assert_array_equal(out, out2)
    assert none is None
    assert int(ref.counts[:, 0].sum()) >= min_hits
    ctx.close()
    return ref


@pytest.mark.parametrize("kw", [
    {},
    {"use_mis": False},
    {"phase": "hg", "hg_g": 0.5},
    {"phase": "hg", "hg_g": -0.3, "seed": 77},
])
def test_planes0d_matches_oracle(built, kw):
    c = H.make_plane_case(**kw)
    _check(c, f"planes {kw}", min_hits=30000)


def test_planes_more_than_one_chunk(built):
    # > 512 planes per staged chunk, > 128 rays per block, ragged tails on both
    _check(H.make_plane_case(n_planes=3001, w=37, h=29), "planes ragged", min_hits=100000)


def test_planes_single_and_empty(built):
    for n in (1, 31, 33):
        _check(H.make_plane_case(n_planes=n, w=16, h=16), f"planes n={n}")
    c = H.make_plane_case(n_planes=64, w=16, h=16)
    c.rays.maxt[:] = 0.0  # empty medium segments: nothing may be gathered
    ref = _check(c, "planes empty rays")
    assert ref.counts.sum() == 0


def test_planes_concentrated_sheet(built):
    # "LASER-style": planes concentrated in a thin sheet, most leaf boxes miss most ray blocks
    c = H.make_plane_case(n_planes=4000, w=64, h=48, sheet=True)
    _check(c, "planes sheet", min_hits=1000)


def test_planes_outside_camera_edge_two(built):
    # edge_id != 1 exercises the t0 Jacobian factor for every plane (shift_volume_planes.h:357-359); the gather
    # itself does not care where the sensor is (the reference driver refuses an outside sensor, gvpm.cpp:785-787)
    c = H.make_plane_case(n_planes=1500, w=32, h=24, inside=False)
    _check(c, "planes outside", min_hits=20000)


def test_sppm_shaped_plane_pass(built):
    """sppm's primal photon planes (PhotonPlaneQuery, plane_struct.h:238-256): intersectPlane0D + getContrib0D and nothing
    else - the gvpm entry with no valid offset path.  Same index sets, the primal equals the oracle's and equals the
    primal of the gradient-domain run (it never depends on the offsets), shifted terms stay zero, every weight is 1."""
    from oracle import binding as ob
    from gvpm_b200.api import Context
    c = H.make_plane_case(n_planes=1500, w=40, h=24)
    ref_grad = ob.planes_gather(c.planes, c.rays, c.medium, c.config)
    rays = c.rays.copy()
    rays.off_valid[:] = 0
    ref = ob.planes_gather(c.planes, rays, c.medium, c.config, neighbours=True)
    ctx = Context(0)
    ctx.set_medium(c.medium)
    ctx.set_config(c.config)
    ctx.upload_planes(c.planes)
    ctx.build_planes()
    ctx.upload_rays(rays)
    out, counts = ctx.gather_planes()
    offsets, idx = ctx.dump_neighbours_planes()
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, "sppm-shaped planes")
    o = out.reshape(-1, 9, 3)
    H.assert_radiance_close(o[:, 0], ref_grad.out.reshape(-1, 9, 3)[:, 0], 1e-4, "primal does not depend on the offsets")
    assert not o[:, 1:5].any()                                  # no shifted contribution
    for k in range(4):                                          # weight 1: the weighted base equals the primal
        H.assert_radiance_close(o[:, 5 + k], o[:, 0], 1e-6, "weighted base with weight 1")
    assert int(ref.counts[:, 0].sum()) > 10000
    ctx.close()
