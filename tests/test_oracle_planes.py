"""CPU self-consistency tests of the G-Planes 0D oracle (oracle/ = test infrastructure; the pin against the reference's
compiled PlaneGradRadianceQuery is tests/test_oracle_functor_pin.py): the
reference-shaped balanced kd-tree + AABB hierarchy + DFS against brute force, a closed-form contribution, the
identity of the specular shift for coincident offset rays, the host mirror of transformBeam, and a committed
regression fixture."""
import os

import numpy as np
import pytest

import gvpm_b200 as g
import gvpm_testlib as H
from gvpm_b200 import records as R
from oracle import binding as ob

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "planes_small.npz")


@pytest.fixture(scope="module")
def case(built):
    return H.make_plane_case(n_planes=1500, w=32, h=20)


def test_plane_tree_equals_bruteforce(case):
    bf = ob.planes_gather(case.planes, case.rays, case.medium, case.config, mode="brute", neighbours=True, threads=4)
    kd = ob.planes_gather(case.planes, case.rays, case.medium, case.config, mode="kdtree", neighbours=True, threads=4)
    assert bf.counts[:, 0].sum() > 20000
    np.testing.assert_array_equal(kd.counts, bf.counts)
    np.testing.assert_array_equal(kd.idx, bf.idx)
    H.assert_radiance_close(kd.out, bf.out, 1e-5, "plane kd vs brute")


def test_plane_closed_form(built):
    """One axis-aligned plane hit head-on: contribution = T_cam * sigma_s^2 * flux / (4 pi) because the
    transmittances along the plane edges cancel against pdfFailure (plane_struct.h:150-192) and |w0.(w1 x d)| = 1."""
    med = g.make_medium(sigma_t=2.0, albedo=0.8)
    pl = R.PlaneSet(1, origin=[0.2, 0.2, 0.5], w0=[1, 0, 0], length0=[0.6], w1=[0, 1, 0], length1=[0.6],
                    flux=[1.0, 2.0, 3.0], edge_id=[1])
    rays = g.synth_rays(1, 1, cam_dist=-0.05)
    rays.o[:] = [0.5, 0.5, 0.0]
    rays.d[:] = [0, 0, 1]
    rays.mint[:] = 1e-4
    rays.maxt[:] = 1.0 - 1e-4
    rays.edge_len[:] = 1.0
    rays.off_valid[:] = 0
    cfg = g.make_config(1, 1)
    res = ob.planes_gather(pl, rays, med, cfg, threads=1)
    assert res.counts[0, 0] == 1
    sig_s = 2.0 * 0.8
    expect = np.exp(-2.0 * 0.5) * sig_s * sig_s * np.array([1.0, 2.0, 3.0]) / (4 * np.pi)
    np.testing.assert_allclose(res.out[0, :3], expect, rtol=2e-6)
    # invalid offsets: weight 1, no shifted flux (GradientSamplingResult defaults, shift_utilities.h:17-23)
    np.testing.assert_array_equal(res.out[0, 3:15], 0)
    np.testing.assert_allclose(res.out[0, 15:].reshape(4, 3), np.tile(res.out[0, :3], (4, 1)), rtol=1e-6)
    # a ray that starts behind the plane misses it (tCam <= mint)
    rays.o[:] = [0.5, 0.5, 0.6]
    assert ob.planes_gather(pl, rays, med, cfg, threads=1).counts[0, 0] == 0
    # outside the parallelogram (t0 > 1)
    rays.o[:] = [0.85, 0.5, 0.0]
    assert ob.planes_gather(pl, rays, med, cfg, threads=1).counts[0, 0] == 0


def test_specular_shift_identity(built):
    """Offset ray == base ray: the rotated w1 equals w1, the Jacobian is 1, S = base and the balance weight is
    1/(1 + sensorPart) (shift_volume_planes.h:263-416)."""
    c = H.make_plane_case(n_planes=800, w=24, h=16)
    r = c.rays
    r.view("off_o")[:] = np.tile(r.view("o"), (1, 4))
    r.view("off_d")[:] = np.tile(r.view("d"), (1, 4))
    r.view("off_len")[:] = r.view("edge_len")
    r.off_valid[:] = 1
    r.off_sensor[:] = 1.0
    res = ob.planes_gather(c.planes, r, c.medium, c.config, double=True, threads=4)
    primal = res.out[:, :3]
    for k in range(4):
        np.testing.assert_allclose(res.out[:, 3 + 3 * k:6 + 3 * k], 0.5 * primal, rtol=1e-5, atol=1e-6 * primal.max())
        np.testing.assert_allclose(res.out[:, 15 + 3 * k:18 + 3 * k], 0.5 * primal, rtol=1e-5, atol=1e-6 * primal.max())
    c.config.use_mis = 0
    res = ob.planes_gather(c.planes, r, c.medium, c.config, double=True, threads=4)
    np.testing.assert_allclose(res.out[:, 15:18], 0.5 * res.out[:, :3], rtol=1e-6)


def test_transform_beam_mirror(built):
    """Host mirror of LTPhotonPlane::transformBeam (gvpm_plane.h:53-73): unit w1, exponential length1 with mean
    1/sigma_t (+ Epsilon), isotropic directions, HG directions with mean cosine g about the beam direction."""
    med = g.make_medium(sigma_t=2.0)
    beams, _ = R.synth_beams(20000, med, seed=3, threads=4)
    pl = R.synth_planes(beams, med, seed=11)
    w0, w1 = pl.view("w0"), pl.view("w1")
    np.testing.assert_allclose(np.linalg.norm(w1, axis=1), 1.0, atol=2e-6)
    np.testing.assert_allclose(np.linalg.norm(w0, axis=1), 1.0, atol=2e-6)
    d = beams.view("end") - beams.view("origin")
    np.testing.assert_allclose(pl.length0, np.linalg.norm(d, axis=1), rtol=1e-6)
    assert abs(pl.length1.mean() - 0.5) < 0.02 and pl.length1.min() >= 1e-4
    assert abs((w0 * w1).sum(axis=1).mean()) < 0.02
    assert np.abs(w1.mean(axis=0)).max() < 0.02
    np.testing.assert_array_equal(pl.edge_id, beams.depth.astype(np.int32))
    hg = g.make_medium(sigma_t=2.0, phase="hg", g=0.6)
    plh = R.synth_planes(beams, hg, seed=11)
    cosm = (plh.view("w0") * plh.view("w1")).sum(axis=1).mean()
    assert abs(cosm - 0.6) < 0.02, cosm
    # same seed, same planes
    np.testing.assert_array_equal(R.synth_planes(beams, med, seed=11).w1, pl.w1)


def test_planes_golden_fixture(built):
    """Regression pin of the plane oracle (fixture made by tests/golden/make_golden.py)."""
    z = np.load(GOLDEN)
    c = H.make_plane_case(**{k: z[k].item() for k in ("n_planes", "w", "h", "seed")})
    res = ob.planes_gather(c.planes, c.rays, c.medium, c.config, neighbours=True, threads=4)
    np.testing.assert_array_equal(res.counts, z["counts"])
    np.testing.assert_array_equal(res.idx, z["idx"])
    H.assert_radiance_close(res.out, z["out"], 1e-5, "planes golden")
