"""The on-disk fixture format: the header-only C++ writer a Mitsuba-side dump hook would use
(gvpm_b200/host/gvpm_fixture.hpp, reached through libgvpm_host.so) and the Python reader / writer agree byte for
byte, a fixture round-trips every input exactly, and the committed fixture still reproduces its recorded results."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import _native as N
from gvpm_b200 import fixture as F
from gvpm_b200 import records as R
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "bre_tiny.gvpmfix")


def _case():
    return H.make_case(n_photons=3000, w=16, h=12, scale=4.0, power_heuristic=True, max_depth=9, rng_seed=77)


def _expected(c):
    return ob.bre_gather(c.photons, c.rays, c.medium, c.config, c.tri, c.radius, mode="brute", neighbours=True, threads=4)


def _same_inputs(fx, c):
    for name, _, _ in R._PHOTON_FIELDS:
        assert np.array_equal(getattr(fx.photons, name), getattr(c.photons, name)), name
    for name, _, _ in R._RAY_FIELDS:
        assert np.array_equal(getattr(fx.rays, name), getattr(c.rays, name)), name
    assert bytes(fx.config) == bytes(c.config) and bytes(fx.medium) == bytes(c.medium)
    assert fx.radius == np.float32(c.radius) and np.array_equal(fx.tri.reshape(-1), np.asarray(c.tri, np.float32).reshape(-1))


def test_cpp_writer_and_python_writer_agree(built, tmp_path):
    c = _case()
    ref = _expected(c)
    p_py, p_cc = str(tmp_path / "py.gvpmfix"), str(tmp_path / "cc.gvpmfix")
    F.save(p_py, c.medium, c.config, c.radius, c.tri, c.photons, c.rays, ref.out, ref.offsets, ref.idx, producer="unit test")
    h = C.CDLL(os.path.join(ROOT, "gvpm_b200", "host", "libgvpm_host.so"))
    err = C.create_string_buffer(256)
    tri = np.ascontiguousarray(c.tri, dtype=np.float32)
    cph, cr = c.photons.as_c(), c.rays.as_c()
    out = np.ascontiguousarray(ref.out, dtype=np.float32)
    rc = h.gvpm_host_write_bre_fixture(p_cc.encode(), C.byref(c.medium), C.byref(c.config), C.c_float(c.radius),
                                       tri.ctypes.data_as(N.f32p), C.c_size_t(tri.size // 9), C.byref(cph),
                                       C.c_size_t(c.photons.n), C.byref(cr), C.c_size_t(c.rays.n),
                                       out.ctypes.data_as(N.f32p), ref.offsets.ctypes.data_as(N.u64p),
                                       ref.idx.ctypes.data_as(N.u32p), b"unit test", err, C.c_size_t(256))
    assert rc == 0, err.value
    assert open(p_py, "rb").read() == open(p_cc, "rb").read()
    fx = F.load(p_cc)
    _same_inputs(fx, c)
    assert fx.producer == "unit test"
    np.testing.assert_array_equal(fx.expected_out, ref.out)
    np.testing.assert_array_equal(fx.expected_idx, ref.idx)
    # the oracle on the loaded fixture reproduces the recorded results bit for bit
    again = ob.bre_gather(fx.photons, fx.rays, fx.medium, fx.config, fx.tri, fx.radius, mode="brute", neighbours=True,
                          threads=4)
    np.testing.assert_array_equal(again.out, ref.out)
    np.testing.assert_array_equal(again.idx, ref.idx)


def test_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.gvpmfix"
    p.write_bytes(b"NOTAFIXTURE" * 4)
    with pytest.raises(ValueError):
        F.read_sections(str(p))
    c = _case()
    good = tmp_path / "good.gvpmfix"
    F.save(str(good), c.medium, c.config, c.radius, c.tri, c.photons, c.rays)
    p.write_bytes(good.read_bytes()[:-100])
    with pytest.raises(ValueError):
        F.read_sections(str(p))
    fx = F.load(str(good))
    assert fx.expected_out is None and fx.expected_offsets is None


def test_committed_fixture_and_checker_tool(built):
    """tests/golden/bre_tiny.gvpmfix (written by tests/golden/make_fixture.py): the oracle still reproduces the recorded
    radiance and neighbour sets; tools/check_fixture.py --no-gpu agrees."""
    fx = F.load(GOLDEN)
    assert fx.expected_out is not None and fx.photons.n > 0
    r = ob.bre_gather(fx.photons, fx.rays, fx.medium, fx.config, fx.tri, fx.radius, mode="brute", neighbours=True, threads=4)
    np.testing.assert_array_equal(r.offsets, fx.expected_offsets)
    np.testing.assert_array_equal(r.idx, fx.expected_idx)
    H.assert_radiance_close(r.out, fx.expected_out, 1e-6, "committed fixture")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_fixture.py"), GOLDEN, "--no-gpu"],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "PASS" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_gpu_against_committed_fixture(built):
    fx = F.load(GOLDEN)
    from gvpm_b200.api import Context
    ctx = Context(0)
    ctx.set_medium(fx.medium)
    ctx.set_config(fx.config)
    ctx.set_occluders(fx.tri)
    ctx.upload_photons(fx.photons)
    ctx.build_points(fx.radius)
    ctx.upload_rays(fx.rays)
    out, _ = ctx.gather_bre()
    offsets, idx = ctx.dump_neighbours_bre()
    ctx.close()
    np.testing.assert_array_equal(offsets, fx.expected_offsets)
    np.testing.assert_array_equal(idx, fx.expected_idx)
    H.assert_radiance_close(out, fx.expected_out, 1e-4, "GPU vs committed fixture")
