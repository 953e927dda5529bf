"""GPU parity of the G-BRE gather against the CPU oracle, through the C ABI (ctypes).

Bar (BASELINE.json north_star): per-ray neighbour counts and index sets bit-exact; primal and the
four gradient contributions within 1e-4 relative (fp32)."""
import numpy as np
import pytest

import gvpm_testlib as H

pytestmark = pytest.mark.gpu


def _oracle(case, **kw):
    from oracle import binding as ob
    return ob.bre_gather(case.photons, case.rays, case.medium, case.config, case.tri, case.radius,
                         mode="brute", neighbours=True, **kw)


def _check(case, what):
    ref = _oracle(case)
    ctx = H.gpu_context(case)
    out, counts = ctx.gather_bre()
    offsets, idx = ctx.dump_neighbours_bre()
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)  # index sets + contributes bit
    worst = H.assert_radiance_close(out, ref.out, 1e-4, what)
    ctx.close()
    return ref, worst


@pytest.mark.parametrize("scale,kw", [
    (1.0, {}),
    (3.0, {}),
    (1.0, {"use_shift_null": False}),
    (1.0, {"path_set": False, "use_mis": False}),
    (2.0, {"power_heuristic": True, "max_depth": 6, "min_depth": 3}),
    (1.0, {"shadow_maxt_scale": 0.999}),
    (2.0, {"lighting_mode": 1 << 4}),
    (2.0, {"lighting_mode": 1 << 2}),
])
def test_bre3d_matches_oracle(built, scale, kw):
    case = H.make_case(n_photons=30000, w=48, h=32, scale=scale, **kw)
    ref, _ = _check(case, f"scale={scale} {kw}")
    assert ref.counts[:, 0].sum() > 1000, "test case too sparse to be meaningful"


def test_bre3d_hg_phase(built):
    case = H.make_case(n_photons=30000, w=48, h=32, scale=2.0, phase="hg", hg_g=0.5)
    _check(case, "hg")


def test_bre2d(built):
    case = H.make_case(n_photons=30000, w=48, h=32, scale=2.0, kernel_3d=False, use_shift_null=False)
    _check(case, "bre2d")


def test_ragged_and_empty(built):
    """1 photon, 33 photons (one full + one ragged leaf), zero rays, zero photons."""
    from gvpm_b200.api import Context
    for n in (1, 33, 1025):
        case = H.make_case(n_photons=n, w=16, h=16, scale=30.0)
        _check(case, f"n={n}")
    case = H.make_case(n_photons=64, w=16, h=16, scale=5.0)
    ctx = Context(0)
    ctx.set_medium(case.medium)
    ctx.set_config(case.config)
    ctx.set_occluders(case.tri)
    ctx.upload_photons(case.photons.take(np.arange(0)))
    ctx.build_points(case.radius)
    ctx.upload_rays(case.rays)
    out, counts = ctx.gather_bre()
    assert not out.any() and not counts.any()
    ctx.upload_rays(case.rays.take(np.arange(0)))
    out, counts = ctx.gather_bre()
    assert out.shape == (0, 27)
    ctx.close()


@pytest.mark.parametrize("order", ["random", "zorder", "strided"])
def test_ray_order_does_not_matter(built, order):
    """The tile traversal groups 32 consecutive rays; incoherent tiles fall back to quads of 4 lanes and to single
    rays.  Whatever the order of the ray list, every ray must gather exactly its own neighbour set."""
    import gvpm_b200 as g
    case = H.make_case(n_photons=30000, w=48, h=32, scale=2.0)
    if order == "zorder":
        case.rays = g.synth_rays(48, 32, seed=0xC0FFEE + 1, block=-16)
    elif order == "random":
        case.rays = case.rays.take(np.random.default_rng(3).permutation(case.rays.n))
    else:  # neighbours in a quad are coherent, quads are far apart
        idx = np.arange(case.rays.n).reshape(-1, 4)
        case.rays = case.rays.take(idx[np.random.default_rng(4).permutation(len(idx))].reshape(-1))
    ref, _ = _check(case, f"ray order {order}")
    assert ref.counts[:, 0].sum() > 1000
