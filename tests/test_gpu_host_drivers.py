"""The other gather drivers of the host-side mirror (gvpm_host::VolumeGatherB200::computeVolumeGradientBeams / Planes /
Photon, gvpm.cpp:880-986, 782-878, 1081-1203) through libgvpm_host.so: oracle gather + the reference's per-iteration
normalisation, APA running mean and radius reduction (beams: cube root, planes: linear, VPM: no APA, divided by the
total number of emitted paths)."""
import ctypes as C
import os

import numpy as np
import pytest

import gvpm_b200 as g
import gvpm_testlib as H
from gvpm_b200 import _native as N
from gvpm_b200 import records as R
from gvpm_b200 import shard

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_lib():
    import __graft_entry__ as ge
    from test_abi_and_host import HostParams
    ge.build()
    hl = C.CDLL(os.path.join(ROOT, "gvpm_b200", "host", "libgvpm_host.so"))
    hl.gvpm_host_create.restype = C.c_void_p
    hl.gvpm_host_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(HostParams), C.POINTER(N.Medium), C.c_float,
                                    N.f32p, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_beams_iteration.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.BeamSoA), C.c_size_t, C.POINTER(N.RaySoA),
                                             C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_planes_iteration.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.PlaneSoA), C.c_size_t, C.POINTER(N.RaySoA),
                                              C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_vpm_iteration.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.PhotonSoA), C.c_size_t, C.POINTER(N.RaySoA),
                                           C.c_size_t, C.POINTER(N.VpmSampleSoA), C.c_size_t, C.c_int, C.c_float, C.c_size_t,
                                           N.u32p, C.c_char_p, C.c_size_t]
    hl.gvpm_host_normalized_accumulators.argtypes = [C.c_void_p, N.f32p, C.c_size_t]
    hl.gvpm_host_scale.restype = C.c_double
    hl.gvpm_host_scale.argtypes = [C.c_void_p]
    hl.gvpm_host_radius.restype = C.c_float
    hl.gvpm_host_radius.argtypes = [C.c_void_p]
    hl.gvpm_host_accumulators.restype = N.f32p
    hl.gvpm_host_accumulators.argtypes = [C.c_void_p]
    hl.gvpm_host_destroy.argtypes = [C.c_void_p]
    return hl


def _create(hl, w, h, medium, tri, **params):
    from test_abi_and_host import host_params
    err = C.create_string_buffer(512)
    p = host_params(**params)
    hd = hl.gvpm_host_create(0, w, h, C.byref(p), C.byref(medium), R.SYNTH_BSPHERE_R, tri.ctypes.data_as(N.f32p),
                             tri.size // 9, err, 512)
    assert hd, err.value
    return hd, err


def test_host_beams_driver_two_iterations(built):
    from oracle import binding as ob
    hl = _host_lib()
    w, h, scale0 = 40, 24, 3.0
    c0 = H.make_case(n_photons=64, w=w, h=h, scale=scale0)
    hd, err = _create(hl, w, h, c0.medium, c0.tri, initialScaleVolume=scale0, volTechnique=3)   # EVolBeam3D
    acc_ref = np.zeros((h, w, 27), dtype=np.float32)
    scale = scale0
    for it in (1, 2):
        c = H.make_case(n_photons=64, w=w, h=h, scale=scale, seed=31 * it)   # rng_seed 0: what the host mirror configures
        beams, n_paths = R.synth_beams(5000, c.medium, seed=7 * it, threads=4)
        radius = hl.gvpm_host_radius(hd)
        assert abs(radius - g.bre_radius(scale)) <= 1e-9
        cb, cr = beams.as_c(), c.rays.as_c()
        assert hl.gvpm_host_beams_iteration(hd, it, C.byref(cb), beams.n, C.byref(cr), c.rays.n, n_paths, err, 512) == 0, err.value
        ref = ob.beams_gather(beams, c.rays, c.medium, c.config, c.tri, radius)
        img = shard.to_image(ref.out, c.rays.px, c.rays.py, w, h)
        acc_ref = (acc_ref * np.float32(it - 1) + img / np.float32(n_paths)) / np.float32(it)
        scale *= ((it - 1 + 0.7) / it) ** (1.0 / 3.0)    # beam3d is a 3-D kernel (volume_utils.h:35-41)
        assert abs(hl.gvpm_host_scale(hd) - scale) < 1e-12
    acc = np.ctypeslib.as_array(hl.gvpm_host_accumulators(hd), shape=(h, w, 27)).copy()
    assert np.abs(acc_ref).max() > 0
    H.assert_radiance_close(acc, acc_ref, 1e-4, "host beams driver accumulators")
    hl.gvpm_host_destroy(hd)


def test_host_planes_driver_two_iterations(built):
    from oracle import binding as ob
    hl = _host_lib()
    w, h, scale0 = 40, 24, 1.0
    c0 = H.make_plane_case(n_planes=64, w=w, h=h)
    tri = g.synth_occluders()
    hd, err = _create(hl, w, h, c0.medium, tri, initialScaleVolume=scale0, volTechnique=4)   # EVolPlane0D
    acc_ref = np.zeros((h, w, 27), dtype=np.float32)
    scale = scale0
    for it in (1, 2):
        c = H.make_plane_case(n_planes=1500, w=w, h=h, seed=0xC0FFEE + it)
        cp, cr = c.planes.as_c(), c.rays.as_c()
        assert hl.gvpm_host_planes_iteration(hd, it, C.byref(cp), c.planes.n, C.byref(cr), c.rays.n, c.n_paths, err, 512) == 0, err.value
        ref = ob.planes_gather(c.planes, c.rays, c.medium, c.config)
        img = shard.to_image(ref.out, c.rays.px, c.rays.py, w, h)
        acc_ref = (acc_ref * np.float32(it - 1) + img / np.float32(c.n_paths)) / np.float32(it)
        scale *= (it - 1 + 0.7) / it                      # plane0d reduces linearly
        assert abs(hl.gvpm_host_scale(hd) - scale) < 1e-12
    acc = np.ctypeslib.as_array(hl.gvpm_host_accumulators(hd), shape=(h, w, 27)).copy()
    assert np.abs(acc_ref).max() > 0
    H.assert_radiance_close(acc, acc_ref, 1e-4, "host planes driver accumulators")
    hl.gvpm_host_destroy(hd)


def test_host_vpm_driver_accumulates_and_normalises(built):
    from oracle import binding as ob
    hl = _host_lib()
    w, h, scale0, nb = 40, 24, 3.0, 8
    c0 = H.make_case(n_photons=64, w=w, h=h, scale=scale0)
    hd, err = _create(hl, w, h, c0.medium, c0.tri, initialScaleVolume=scale0, volTechnique=2)   # EVolVPM
    acc_ref = np.zeros((h, w, 27), dtype=np.float64)
    total = 0
    for it in (1, 2):
        c = H.make_case(n_photons=40000, w=w, h=h, scale=scale0, seed=17 * it)
        rad = np.full(c.rays.n, c.radius, dtype=np.float32) * np.random.default_rng(it).uniform(0.5, 1.0, c.rays.n).astype(np.float32)
        smp = g.synth_vpm_samples(c.rays, c.medium, rad, nb_camera_samples=nb, seed=99 + it)
        mvol = np.zeros(c.rays.n, dtype=np.uint32)
        cph, cr, cs = c.photons.as_c(), c.rays.as_c(), smp.as_c()
        rc = hl.gvpm_host_vpm_iteration(hd, it, C.byref(cph), c.photons.n, C.byref(cr), c.rays.n, C.byref(cs), smp.n, nb,
                                        C.c_float(c.radius), c.n_paths, mvol.ctypes.data_as(N.u32p), err, 512)
        assert rc == 0, err.value
        ref = ob.vpm_gather(c.photons, c.rays, smp, c.medium, c.config, c.tri, nb, mode="brute")
        np.testing.assert_array_equal(mvol.astype(np.float32), ref.mvol)
        acc_ref += shard.to_image(ref.out, c.rays.px, c.rays.py, w, h)
        total += c.n_paths
        assert abs(hl.gvpm_host_scale(hd) - scale0) < 1e-12     # no APA radius reduction on this path
    got = np.zeros((h, w, 27), dtype=np.float32)
    assert hl.gvpm_host_normalized_accumulators(hd, got.ctypes.data_as(N.f32p), got.size) == 0
    assert np.abs(acc_ref).max() > 0
    H.assert_radiance_close(got, (acc_ref / total).astype(np.float32), 1e-4, "host VPM driver, normalised accumulators")
    hl.gvpm_host_destroy(hd)
