"""Pins the oracle restatement against the REFERENCE'S OWN code.

oracle/_ref/libgvpm_ref.so is built from /root/reference (oracle/Makefile, target `ref`): the reference's
PointKDTree (build + range query), AABB slab test, GPhotonMap + GradientBeamRadianceEstimator (hierarchy +
traversal), SubBeamBVH, PhotonPlaneBVH, cylinderIntersection, PhotonBeam::rayIntersectInternal1D,
PhotonPlane::intersectPlane0D, Triangle::rayIntersect, coordinateSystem(Coherent), solveQuadraticDouble.
tests/golden/ref_pins.npz holds its outputs on seeded inputs (tests/golden/make_ref_golden.py), so the pin
also holds where the reference tree is absent.  Everything is compared BIT-EXACTLY (floats as integer bits).
"""
import os

import numpy as np
import pytest

import pin_cases
from oracle import ref_binding as rb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_pins.npz")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def small():
    return pin_cases.inputs("small")


@pytest.fixture(scope="module")
def oracle_small(small):
    return pin_cases.run(rb.Side("oracle"), small)


def _same(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = int(np.count_nonzero(a != b))
    assert bad == 0, f"{what}: {bad} of {a.size} entries differ from the reference"


def test_golden_is_not_trivial(golden):
    # the vectors exercise every routine: hits and misses on both sides of each predicate
    assert golden["bre_idx"].size > 1000 and golden["range_idx"].size > 30
    for k in ("cyl_hit", "pl_hit", "b1d_hit", "tri_hit", "quad_ok"):
        assert 0 < golden[k].sum() < golden[k].size, k
    assert golden["kd_leaf"].sum() > 100 and golden["kd_depth"] > 10


@pytest.mark.parametrize("key", [
    "kd_depth", "kd_orig", "kd_right", "kd_leaf", "kd_axis",                 # PointKDTree::build (sliding midpoint)
    "bre_depth", "bre_off", "bre_idx", "bre_tdisk_bits",                     # BRE hierarchy + traversal + predicate
    "range_off", "range_idx",                                                # PointKDTree::executeQuery
    "cyl_hit", "cyl_tnear_bits", "cyl_tfar_bits",                            # cylinderIntersection
    "pl_hit", "pl_out_bits",                                                 # PhotonPlane::intersectPlane0D
    "b1d_hit", "b1d_out_bits",                                               # PhotonBeam::rayIntersectInternal1D
    "tri_hit",                                                               # Triangle::rayIntersect + interval
    "cs0_b_bits", "cs0_c_bits", "cs1_b_bits", "cs1_c_bits",                  # coordinateSystem / Coherent
    "quad_ok", "quad_x0_bits", "quad_x1_bits"])                              # solveQuadraticDouble
def test_oracle_equals_reference_golden(golden, oracle_small, key):
    _same(np.asarray(oracle_small[key]), np.asarray(golden[key]), key)


def _pairs_all(inp, n_prim):
    n_rays = len(inp["ray_o"])
    r = np.repeat(np.arange(n_rays), n_prim)
    p = np.tile(np.arange(n_prim), n_rays)
    return r, p


def test_reference_plane_bvh_offers_every_intersected_plane(golden, small):
    """PhotonPlaneBVH::query hands the functor every plane whose subtree box the ray hits; the set of planes
    that then pass intersectPlane0D must equal the brute-force set the oracle's plane gather is defined on."""
    inp = small
    nb = len(inp["pl_ori"])
    r, p = _pairs_all(inp, nb)
    hit, _ = rb.Side("oracle").plane0d(inp["pl_ori"][p], inp["pl_w0"][p], inp["pl_len0"][p], inp["pl_w1"][p],
                                       inp["pl_len1"][p], inp["ray_o"][r], inp["ray_d"][r], inp["ray_mint"][r],
                                       inp["ray_maxt"][r])
    brute = set(zip(r[hit].tolist(), p[hit].tolist()))
    off, idx = golden["plv_off"], golden["plv_idx"]
    offered = set()
    for i in range(len(off) - 1):
        offered.update((i, int(j)) for j in idx[off[i]:off[i + 1]])
    assert len(brute) > 50
    assert brute <= offered, f"{len(brute - offered)} intersected planes never reached the reference's functor"


def test_reference_subbeam_bvh_offers_every_intersected_beam(golden, small):
    """SubBeamBVH::query offers (beam, t1, t2) for every node whose subtree box is hit; every (ray, beam) pair
    whose cylinder test succeeds with tNear inside the beam must be offered with a sub-beam owning tNear
    (shift_volume_beams.h:214-220: tNear < 0 is owned by the sub-beam with t1 == 0, otherwise t1 < tNear < t2)."""
    inp = small
    nb = len(inp["beam_o"])
    r, p = _pairs_all(inp, nb)
    bo, be = inp["beam_o"][p], inp["beam_e"][p]
    bd = be - bo
    bl = np.sqrt((bd * bd).sum(1)).astype(np.float32)
    bd = (bd / bl[:, None]).astype(np.float32)
    ro, rd, mint, maxt = inp["ray_o"][r], inp["ray_d"][r], inp["ray_mint"][r], inp["ray_maxt"][r]
    hit, tn, _ = rb.Side("oracle").cylinder((ro + rd * mint[:, None]).astype(np.float32), rd,
                                            (maxt - mint).astype(np.float32), bo, bd, bl,
                                            np.full(len(r), inp["radius"], np.float32))
    off, idx = golden["sub_off"], golden["sub_idx"]
    t1, t2 = golden["sub_t1_bits"].view(np.float32), golden["sub_t2_bits"].view(np.float32)
    offered = {}
    for i in range(len(off) - 1):
        for k in range(int(off[i]), int(off[i + 1])):
            offered.setdefault((i, int(idx[k])), []).append((t1[k], t2[k]))
    # sub-beam cut: ceil(len / (avgLen / 10)) pieces of equal length (beams_accel.h:99-129, beams_struct.h:316-318)
    all_len = np.sqrt(((inp["beam_e"] - inp["beam_o"]) ** 2).sum(1))
    assert t2.max() <= all_len.max() * 1.001
    n_checked = 0
    for k in np.nonzero(hit)[0]:
        key = (int(r[k]), int(p[k]))
        near = tn[k]
        if not (near < 0 or 0 < near < bl[k]):
            continue
        # keep clear of sub-beam boundaries and of the beam end, where the owner is decided by the last bit
        segs = offered.get(key, [])
        own = [(a, b) for a, b in segs if (near < 0 and a == 0) or (a < near < b)]
        margin = min([abs(near - a) for a, _ in segs] + [abs(near - b) for _, b in segs] + [1.0]) if segs else 1.0
        if not own and margin < 1e-5:
            continue
        assert own, f"ray {key[0]} beam {key[1]} tNear {near}: no owning sub-beam offered by the reference BVH"
        n_checked += 1
    assert n_checked > 20


@pytest.mark.skipif(not rb.have_ref(), reason="oracle/_ref/libgvpm_ref.so not built (reference tree absent)")
def test_oracle_equals_live_reference_large():
    """Same comparison against the reference code itself on 60 k photons / 3 k rays / 200 k element queries."""
    inp = pin_cases.inputs("large", seed=7)
    a = pin_cases.run(rb.Side("ref"), inp)
    b = pin_cases.run(rb.Side("oracle"), inp)
    assert a["bre_idx"].size > 20000
    for k in a:
        _same(np.asarray(b[k]), np.asarray(a[k]), k)


@pytest.mark.skipif(not rb.have_ref(), reason="oracle/_ref/libgvpm_ref.so not built (reference tree absent)")
def test_golden_file_is_current(golden, small):
    """The committed vectors are what the reference code produces today."""
    a = pin_cases.run(rb.Side("ref"), small)
    for k in a:
        _same(np.asarray(golden[k]), np.asarray(a[k]), k)
