"""Rows a1 / a10 of SURVEY.md §8: the reference-side flattening shim (gvpm_b200/host/gvpm_mitsuba_shim.hpp) is real code
against the reference's own types (Path, PathVertex, GPhotonMap, LTBeamMap's beams, GatherPoint / ShiftGatherPoint,
Medium, Scene / TriMesh).  Mitsuba cannot be linked here, so the check is the compiler's: every entry
point is instantiated in tests/mitsuba_shim_check.cpp and type-checked against the reference tree."""
import os
import subprocess

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
