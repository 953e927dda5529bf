"""CPU checks of the restated sppm primal BRE (bre.cpp:167-259, sppm.cpp:926-981)."""
import numpy as np
import pytest

import gvpm_testlib as H
from oracle import binding as ob


@pytest.mark.parametrize("k3d", [True, False])
def test_sppm_tree_equals_bruteforce(built, k3d):
    """The reference-shaped kd/AABB traversal on the re-based ray selects exactly the brute-force set."""
    c = H.make_case(n_photons=20000, w=32, h=24, scale=2.5, kernel_3d=k3d, use_shift_null=False, sppm_primal=True,
                    rng_seed=3)
    a = ob.sppm_bre_gather(c.photons, c.rays, c.medium, c.config, c.radius, mode="kdtree", threads=4, neighbours=True)
    b = ob.sppm_bre_gather(c.photons, c.rays, c.medium, c.config, c.radius, mode="brute", threads=4, neighbours=True)
    assert a.counts[:, 0].sum() > 2000
    np.testing.assert_array_equal(a.counts, b.counts)
    np.testing.assert_array_equal(a.idx, b.idx)
    H.assert_radiance_close(a.out, b.out, 1e-5, "sppm tree vs brute")


def test_sppm_2d_matches_gvpm_primal_up_to_sigma_s(built):
    """Same estimator, two code paths: gvpm's BRE-2D primal (no pathSet) = sigma_s x sppm's BRE-2D result
    (gvpm folds sigma_s into the contribution, shift_volume_photon.h:79-86; sppm's photon power already carries it),
    except for photons in the last Epsilon of the segment (maxt vs edgeLen bound)."""
    c = H.make_case(n_photons=20000, w=32, h=24, scale=2.5, kernel_3d=False, use_shift_null=False, path_set=False,
                    max_depth=-1)
    g = ob.bre_gather(c.photons, c.rays, c.medium, c.config, c.tri, c.radius, mode="brute", threads=4)
    c.config.sppm_primal = 1
    s = ob.sppm_bre_gather(c.photons, c.rays, c.medium, c.config, c.radius, mode="brute", threads=4)
    same = g.counts[:, 1] == s.counts[:, 1]
    assert same.mean() > 0.98
    sigma_s = np.array(list(c.medium.sigma_s), dtype=np.float64)
    want = s.out[same].astype(np.float64) * sigma_s[None, :]
    got = g.out[same, :3].astype(np.float64)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
    assert err.max() < 5e-4, err.max()


def test_sppm_depth_filter_and_counts(built):
    c = H.make_case(n_photons=20000, w=24, h=16, scale=2.5, sppm_primal=True, max_depth=3, rng_seed=11)
    r = ob.sppm_bre_gather(c.photons, c.rays, c.medium, c.config, c.radius, mode="brute", threads=4, neighbours=True)
    assert (r.counts[:, 1] <= r.counts[:, 0]).all() and r.counts[:, 1].sum() < r.counts[:, 0].sum()
    contributes = (r.idx >> np.uint32(31)).astype(bool)
    depth = c.photons.depth[r.idx & np.uint32(0x7FFFFFFF)].astype(np.int64)
    ray_of = np.repeat(np.arange(c.rays.n), np.diff(r.offsets.astype(np.int64)))
    np.testing.assert_array_equal(contributes, depth <= 3 - c.rays.edge_id[ray_of])
