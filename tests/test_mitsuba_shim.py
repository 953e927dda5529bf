"""Rows a1 / a10 of SURVEY.md §8: the reference-side flattening shim (gvpm_b200/host/gvpm_mitsuba_shim.hpp) is real code
against the reference's own types (Path, PathVertex, GPhotonMap, LTBeamMap's beams, GatherPoint / ShiftGatherPoint,
Medium, Scene / TriMesh).  Mitsuba as a whole cannot be linked here: every entry point is instantiated in
tests/mitsuba_shim_check.cpp and type-checked against the reference tree, and the photon / beam / gather-point / medium
shims are also EXECUTED on reference objects (round trips below, through oracle/_ref/libgvpm_functor_ref.so)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GVPM_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not present on this machine")
def test_shim_compiles_against_the_reference_headers():
    pm = os.path.join(REF, "src", "integrators", "photonmapper")
    cmd = ["g++", "-std=gnu++14", "-fsyntax-only", "-w", "-include", "unistd.h", "-include", "math.h", "-DSINGLE_PRECISION",
           "-DSPECTRUM_SAMPLES=3", "-DMTS_SSE", "-I" + os.path.join(ROOT, "oracle", "ref_shim"),
           "-I" + os.path.join(REF, "include"), "-I" + pm, "-I" + os.path.join(REF, "src", "integrators"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "gvpm_b200", "host"),
           os.path.join(ROOT, "tests", "mitsuba_shim_check.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_shim_is_not_part_of_the_product_library():
    """it needs Mitsuba's headers: nothing under gvpm_b200/csrc or the host C API includes it"""
    for d in ("csrc", "host"):
        for f in os.listdir(os.path.join(ROOT, "gvpm_b200", d)):
            if f == "gvpm_mitsuba_shim.hpp":
                continue
            with open(os.path.join(ROOT, "gvpm_b200", d, f), errors="ignore") as fh:
                assert "gvpm_mitsuba_shim" not in fh.read(), f


# ---- executed, not only compiled: round trips through the shims on the reference's own objects ----------------------------------
# oracle/ref_functor.cpp (TEST INFRASTRUCTURE) rebuilds the reference's Path / GPhotonMap / LTPhotonBeam / GatherPoint objects
# from flattened records for the functor pins; the shim header is compiled into that library too and run on those objects.
# records -> reference objects -> shim -> records must give the input back.
def _fields(a, b, names):
    for name in names:
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"{name} does not survive the round trip"


def _need_ref():
    from oracle import functor_binding as fb
    if not fb.have_ref():
        pytest.skip("prebuilt reference library absent")
    return fb


@pytest.mark.parametrize("case", ["default", "hg_forward_0.7", "big"])
def test_flatten_photon_map_round_trip(built, case):
    """gvpm_shim::flattenPhotonMap (row a1) on a GPhotonMap of rebuilt light paths: every array comes back bit for bit -
    positions, running weight, parent / predecessor data, prefix product, albedo read from the BSDF, pdfs, the parent
    classification by VertexClassifier (forward HG media: glossy medium vertices still map to GVPM_PARENT_MEDIUM)."""
    import functor_pin_cases as cases
    fb = _need_ref()
    c = cases.bre_case(case)
    out = fb.shim_roundtrip("photons", c.photons, c.medium, c.config)
    _fields(c.photons, out, [f[0] for f in c.photons.FIELDS])


def test_append_light_path_beams_round_trip(built):
    """gvpm_shim::appendLightPathBeams (row a13) on the rebuilt light path of every beam.  `flux` is left out: the harness'
    light paths carry unit vertex weights, the flattened flux includes vertex(i).weight."""
    import functor_pin_cases as cases
    fb = _need_ref()
    c = cases.beams_case("default")
    out = fb.shim_roundtrip("beams", c.beams, c.medium, c.config, c.radius)
    _fields(c.beams, out, [f[0] for f in c.beams.FIELDS if f[0] != "flux"])
    rr = np.repeat(c.beams.rr_weight, 3)
    np.testing.assert_array_equal(out.flux, c.beams.prefix_flux * rr)       # prefix * vertex(i).weight (= 1) * rrWeight


@pytest.mark.parametrize("edge", [1, 2])
def test_append_gather_point_round_trip(built, edge):
    """gvpm_shim::appendGatherPoint (row a10) on rebuilt GatherPoint / ShiftGatherPoint objects: the ray is re-derived from
    the path's vertex positions as gvpm.cpp:1027-1038 does (direction and maxt within rounding), everything else - eye
    weights, pixel, sampler draw, offset edges, validVolumeEdge, sensorMIS - comes back bit for bit (sensorMIS at later
    edges within the rounding of its cancelling geometry terms)."""
    import functor_pin_cases as cases
    fb = _need_ref()
    c = cases.bre_case("invalid_offsets")
    c.rays.edge_id[:] = edge
    out = fb.shim_roundtrip("rays", c.rays, c.medium, c.config)
    _fields(c.rays, out, ["o", "mint", "edge_len", "eye_contrib", "xi", "px", "py", "edge_id", "off_valid"])
    np.testing.assert_allclose(out.d, c.rays.d, rtol=0, atol=3e-7)
    np.testing.assert_allclose(out.maxt, c.rays.maxt, rtol=3e-7, atol=0)
    v = c.rays.off_valid.astype(bool)
    assert (~v).any() and v.any()
    v3 = np.repeat(v, 3)
    for name, m in (("off_o", v3), ("off_d", v3), ("off_len", v), ("off_eye", v3)):
        assert np.array_equal(getattr(out, name)[m], getattr(c.rays, name)[m]), name
    assert not out.off_len[~v].any() and not out.off_sensor[~v].any()         # invalid offsets: zeroed records
    if edge == 1:
        assert np.array_equal(out.off_sensor[v], c.rays.off_sensor[v])
    else:
        np.testing.assert_allclose(out.off_sensor[v], c.rays.off_sensor[v], rtol=1e-6, atol=0)


def test_flatten_medium_round_trip(built):
    import gvpm_b200 as g
    fb = _need_ref()
    for kw in (dict(), dict(phase="hg", g=0.7), dict(phase="hg", g=-0.3, albedo=0.5, sigma_t=3.0)):
        m = g.make_medium(**kw)
        out = fb.shim_medium(m)
        assert bytes(out) == bytes(m), kw
