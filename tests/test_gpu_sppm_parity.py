"""GPU parity of the sppm primal BRE entry (gvpm_gather_sppm_bre) against the CPU oracle,
through the C ABI: counts and photon index sets bit-exact, radiance within 1e-4 relative (fp32)."""
import numpy as np
import pytest

import gvpm_testlib as H

pytestmark = pytest.mark.gpu


def _check(c, what):
    from oracle import binding as ob
    ref = ob.sppm_bre_gather(c.photons, c.rays, c.medium, c.config, c.radius, mode="brute", neighbours=True)
    ctx = H.gpu_context(c)
    out, counts = ctx.gather_sppm_bre()
    out_fast, _ = ctx.gather_sppm_bre(counts=False)
    offsets, idx = ctx.dump_neighbours_bre()
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    H.assert_radiance_close(out_fast, ref.out, 1e-4, what + " (prefiltered)")
    ctx.close()
    return ref


@pytest.mark.parametrize("kw", [
    {},
    {"kernel_3d": False, "use_shift_null": False},
    {"max_depth": 4},
    {"max_depth": -1, "rng_seed": 77},
])
def test_sppm_bre_matches_oracle(built, kw):
    c = H.make_case(n_photons=30000, w=48, h=32, scale=2.0, sppm_primal=True, **kw)
    ref = _check(c, f"sppm bre {kw}")
    assert ref.counts[:, 0].sum() > 3000


def test_sppm_bre_hg_small_radius(built):
    _check(H.make_case(n_photons=200000, w=64, h=48, scale=0.5, phase="hg", hg_g=0.5, sppm_primal=True), "sppm hg")


def test_sppm_requires_flag(built):
    c = H.make_case(n_photons=1000, w=8, h=8, scale=2.0)
    ctx = H.gpu_context(c)
    with pytest.raises(RuntimeError):
        ctx.gather_sppm_bre()
    ctx.close()
