"""ctypes binding of oracle/_ref/libgvpm_functor_ref.so: the REFERENCE'S OWN BRE shift functor
(VolumeGradientBREQuery::operator(), gvpm/shift/shift_volume_photon.cpp:658-856, with shiftNull / shiftPhotonDiffuse /
getShiftPos and everything they evaluate) built by `make -C oracle functor_ref` from /root/reference and driven on the
flattened C-ABI inputs by oracle/ref_functor.cpp.

TEST INFRASTRUCTURE: used by tests/test_oracle_functor_pin.py and tests/golden/make_functor_golden.py only."""
import ctypes as C
import os
import subprocess

import numpy as np

from gvpm_b200 import _native as N

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libgvpm_functor_ref.so")
REFERENCE_ROOT = os.environ.get("GVPM_REFERENCE_ROOT", "/root/reference")
_lib = None


def build_ref():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", _HERE, "functor_ref", f"REF={REFERENCE_ROOT}"])
    return os.path.exists(REF_LIB)


def have_ref():
    return os.path.exists(REF_LIB)


def load():
    global _lib
    if _lib is None:
        if not have_ref():
            raise FileNotFoundError(REF_LIB)
        _lib = C.CDLL(REF_LIB)
        _lib.ref_fn_bre_gather.restype = C.c_int
        _lib.ref_fn_bre_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                           N.f32p, C.c_size_t, C.c_float, N.f32p, N.u32p]
        _lib.ref_fn_vpm_gather.restype = C.c_int
        _lib.ref_fn_vpm_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_void_p, N.f32p, C.c_size_t, C.c_int, N.f32p, N.f32p]
        _lib.ref_fn_beams_gather.restype = C.c_int
        _lib.ref_fn_beams_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                             N.f32p, C.c_size_t, C.c_float, N.f32p, N.f32p, N.u32p]
        _lib.ref_fn_planes_gather.restype = C.c_int
        _lib.ref_fn_planes_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                              N.f32p, N.u32p]
        _lib.ref_fn_sppm_beams_gather.restype = C.c_int
        _lib.ref_fn_sppm_beams_gather.argtypes = [C.c_void_p, C.c_size_t, N.u32p, N.f32p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                  C.c_void_p, C.c_void_p, C.c_float, C.c_int, N.f32p, N.f32p, N.u32p]
        _lib.ref_fn_sppm_bre_gather.restype = C.c_int
        _lib.ref_fn_sppm_bre_gather.argtypes = [N.f32p, N.f32p, N.f32p, N.u8p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                                C.c_void_p, C.c_float, N.f32p, N.f32p]
        _lib.ref_fn_rgbe_roundtrip.restype = None
        _lib.ref_fn_rgbe_roundtrip.argtypes = [N.f32p, C.c_size_t, N.f32p]
        _lib.ref_fn_sppm_planes_gather.restype = C.c_int
        _lib.ref_fn_sppm_planes_gather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                                   N.f32p, N.u32p]
        _lib.ref_fn_bre_pass.restype = C.c_int
        _lib.ref_fn_bre_pass.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                         N.f32p, C.c_size_t, C.c_float, C.c_int, N.f32p, N.u32p, C.POINTER(C.c_double)]
        _lib.ref_fn_bre_open.restype = C.c_void_p
        _lib.ref_fn_bre_open.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, N.f32p, C.c_size_t, C.c_float,
                                         C.c_int, C.POINTER(C.c_double)]
        _lib.ref_fn_bre_run.restype = C.c_int
        _lib.ref_fn_bre_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, N.f32p, N.u32p,
                                        C.POINTER(C.c_double)]
        _lib.ref_fn_bre_close.restype = None
        _lib.ref_fn_bre_close.argtypes = [C.c_void_p]
        _lib.ref_fn_beams_pass.restype = C.c_int
        _lib.ref_fn_beams_pass.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, N.f32p,
                                           C.c_size_t, C.c_float, N.f32p, C.c_int, N.f32p, N.u32p, C.POINTER(C.c_double)]
        _lib.ref_fn_planes_pass.restype = C.c_int
        _lib.ref_fn_planes_pass.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int,
                                            N.f32p, C.POINTER(C.c_double)]
        _lib.ref_fn_vpm_pass.restype = C.c_int
        _lib.ref_fn_vpm_pass.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_void_p, N.f32p, C.c_size_t, C.c_int, C.c_int, N.f32p, N.f32p,
                                         C.POINTER(C.c_double)]
        dp = C.POINTER(C.c_double)
        _lib.ref_fn_tech_close.restype = None
        _lib.ref_fn_tech_close.argtypes = [C.c_void_p]
        _lib.ref_fn_beams_open.restype = C.c_void_p
        _lib.ref_fn_beams_open.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, N.f32p, C.c_size_t, C.c_float, dp]
        _lib.ref_fn_beams_run.restype = C.c_int
        _lib.ref_fn_beams_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, N.f32p, C.c_int, N.f32p, N.u32p, dp]
        _lib.ref_fn_planes_open.restype = C.c_void_p
        _lib.ref_fn_planes_open.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, dp]
        _lib.ref_fn_planes_run.restype = C.c_int
        _lib.ref_fn_planes_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, N.f32p, dp]
        _lib.ref_fn_vpm_open.restype = C.c_void_p
        _lib.ref_fn_vpm_open.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, N.f32p, C.c_size_t, C.c_int, dp]
        _lib.ref_fn_vpm_run.restype = C.c_int
        _lib.ref_fn_vpm_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                        N.f32p, N.f32p, dp]
        for fn in ("ref_fn_shim_photons", "ref_fn_shim_rays"):
            getattr(_lib, fn).restype = C.c_int
            getattr(_lib, fn).argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_fn_shim_beams.restype = C.c_int
        _lib.ref_fn_shim_beams.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        _lib.ref_fn_shim_medium.restype = C.c_int
        _lib.ref_fn_shim_medium.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def bre_gather(photons, rays, medium, config, tri, radius):
    """The reference functor over every (ray, photon) pair of the neighbour predicate (gvpm_accel.h:293-305), photons in
    index order.  Returns (out [n_rays, 27] = mediumFlux, shiftedMediumFlux[4], weightedMediumFlux[4]; counts [n_rays])."""
    lib = load()
    cph, cr = photons.as_c(), rays.as_c()
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    out = np.zeros(rays.n * 27, dtype=np.float32)
    counts = np.zeros(rays.n, dtype=np.uint32)
    rc = lib.ref_fn_bre_gather(C.byref(cph), photons.n, C.byref(cr), rays.n, C.byref(medium), C.byref(config),
                               tri.ctypes.data_as(N.f32p), tri.size // 9, radius, out.ctypes.data_as(N.f32p),
                               counts.ctypes.data_as(N.u32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_bre_gather refused the input: {rc}")
    return out.reshape(rays.n, 27), counts


def vpm_gather(photons, rays, samples, medium, config, tri, nb_camera_samples):
    """The reference's point-VPM functor (VolumeGradientPositionQuery::operator(), shift_volume_photon.cpp:489-655) over the
    host-drawn distance samples, folded per pixel as gvpm.cpp:1175-1182 does.  Returns (out [n_rays, 27], MVol [n_rays])."""
    lib = load()
    cph, cr, cs = photons.as_c(), rays.as_c(), samples.as_c()
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    out = np.zeros(rays.n * 27, dtype=np.float32)
    mvol = np.zeros(rays.n, dtype=np.float32)
    rc = lib.ref_fn_vpm_gather(C.byref(cph), photons.n, C.byref(cr), rays.n, C.byref(cs), samples.n, C.byref(medium),
                               C.byref(config), tri.ctypes.data_as(N.f32p), tri.size // 9, nb_camera_samples,
                               out.ctypes.data_as(N.f32p), mvol.ctypes.data_as(N.f32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_vpm_gather refused the input: {rc}")
    return out.reshape(rays.n, 27), mvol


def beam_uniforms(beams, rays, medium, config):
    """The two sampler draws per (ray, beam) pair as the C ABI derives them (oracle: Scene::beamUniform, dims 0 and 1)."""
    from oracle import binding as ob
    lib = ob.load()
    cr = rays.as_c()
    xi = np.zeros(rays.n * beams.n * 2, dtype=np.float32)
    lib.gvpm_oracle_beam_uniforms.restype = None
    lib.gvpm_oracle_beam_uniforms.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, N.f32p]
    lib.gvpm_oracle_beam_uniforms(C.byref(cr), rays.n, beams.n, C.byref(medium), C.byref(config), xi.ctypes.data_as(N.f32p))
    return xi


def beams_gather(beams, rays, medium, config, tri, radius):
    """The reference's photon-beam functor (BeamGradRadianceQuery::operator(), shift_volume_beams.cpp:139-353) over every
    (ray, beam) pair in beam order.  Returns (out [n_rays, 27], counts [n_rays, 2])."""
    lib = load()
    xi = beam_uniforms(beams, rays, medium, config)
    cb, cr = beams.as_c(), rays.as_c()
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    out = np.zeros(rays.n * 27, dtype=np.float32)
    counts = np.zeros(rays.n * 2, dtype=np.uint32)
    rc = lib.ref_fn_beams_gather(C.byref(cb), beams.n, C.byref(cr), rays.n, C.byref(medium), C.byref(config),
                                 tri.ctypes.data_as(N.f32p), tri.size // 9, radius, xi.ctypes.data_as(N.f32p),
                                 out.ctypes.data_as(N.f32p), counts.ctypes.data_as(N.u32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_beams_gather refused the input: {rc}")
    return out.reshape(rays.n, 27), counts.reshape(rays.n, 2)


def planes_gather(planes, rays, medium, config):
    """The reference's photon-plane functor (PlaneGradRadianceQuery::operator(), shift_volume_planes.h:56-101) over every
    (ray, plane) pair in plane order.  Returns (out [n_rays, 27], counts [n_rays, 2])."""
    lib = load()
    cp, cr = planes.as_c(), rays.as_c()
    out = np.zeros(rays.n * 27, dtype=np.float32)
    counts = np.zeros(rays.n * 2, dtype=np.uint32)
    rc = lib.ref_fn_planes_gather(C.byref(cp), planes.n, C.byref(cr), rays.n, C.byref(medium), C.byref(config),
                                  out.ctypes.data_as(N.f32p), counts.ctypes.data_as(N.u32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_planes_gather refused the input: {rc}")
    return out.reshape(rays.n, 27), counts.reshape(rays.n, 2)


def sppm_beams_gather(beams, rays, medium, config, radius, technique):
    """sppm's primal beam functor (BeamRadianceQuery<PhotonBeam>::operator(), photonmapper/beams.h:29-223) over every
    (camera beam, sub-beam) pair of the reference's sub-beam split, in table order.  technique: key of
    oracle.binding.BEAM_TECHNIQUES.  Returns (Li [n_rays, 3] WITHOUT the camera beam's weight, counts [n_rays, 2])."""
    from oracle import binding as ob
    lib, olib = load(), ob.load()
    tech = ob.BEAM_TECHNIQUES[technique]
    t12, bi = ob.subbeams(beams)
    n_sub = len(bi)
    # ordinal of each sub-beam within its beam (the table is in beam order)
    first = np.r_[True, bi[1:] != bi[:-1]]
    start = np.maximum.accumulate(np.where(first, np.arange(n_sub), 0))
    k = (np.arange(n_sub) - start).astype(np.uint32)
    naive = technique == "beam3d_naive"
    dim0 = (2 + 2 * k if naive else np.zeros(n_sub)).astype(np.uint32)
    dim1 = (3 + 2 * k if naive else np.ones(n_sub)).astype(np.uint32)
    cb, cr = beams.as_c(), rays.as_c()
    xi = np.zeros(rays.n * n_sub * 2, dtype=np.float32)
    olib.gvpm_oracle_beam_uniform_dims.restype = None
    olib.gvpm_oracle_beam_uniform_dims.argtypes = [C.c_void_p, C.c_size_t, N.u32p, N.u32p, N.u32p, C.c_size_t, C.c_void_p,
                                                   C.c_void_p, N.f32p]
    bi = np.ascontiguousarray(bi, dtype=np.uint32)
    olib.gvpm_oracle_beam_uniform_dims(C.byref(cr), rays.n, bi.ctypes.data_as(N.u32p), dim0.ctypes.data_as(N.u32p),
                                       dim1.ctypes.data_as(N.u32p), n_sub, C.byref(medium), C.byref(config),
                                       xi.ctypes.data_as(N.f32p))
    t12 = np.ascontiguousarray(t12, dtype=np.float32)
    out = np.zeros(rays.n * 3, dtype=np.float32)
    counts = np.zeros(rays.n * 2, dtype=np.uint32)
    rc = lib.ref_fn_sppm_beams_gather(C.byref(cb), beams.n, bi.ctypes.data_as(N.u32p), t12.ctypes.data_as(N.f32p), n_sub,
                                      C.byref(cr), rays.n, C.byref(medium), C.byref(config), radius, tech,
                                      xi.ctypes.data_as(N.f32p), out.ctypes.data_as(N.f32p), counts.ctypes.data_as(N.u32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_sppm_beams_gather refused the input: {rc}")
    return out.reshape(rays.n, 3), counts.reshape(rays.n, 2)


def rgbe_roundtrip(flux):
    """Spectrum::toRGBE + fromRGBE: what the stock Photon does to its power (photon.h:40-44,124-132)."""
    lib = load()
    a = np.ascontiguousarray(flux, dtype=np.float32).reshape(-1)
    out = np.zeros_like(a)
    lib.ref_fn_rgbe_roundtrip(a.ctypes.data_as(N.f32p), a.size // 3, out.ctypes.data_as(N.f32p))
    return out


def sppm_bre_gather(photons, direction, rays, medium, config, radius):
    """The loop body of sppm's BeamRadianceEstimator::query (photonmapper/bre.cpp:167-259) on every (ray, photon) pair in
    photon order.  `direction` [n, 3] is photon.getDirection(); photons.flux must be RGBE-representable.  Returns
    out [n_rays, 3] without beam.weight and m_scaleFactor."""
    from oracle import binding as ob
    lib, olib = load(), ob.load()
    cr = rays.as_c()
    xi = np.zeros(rays.n * photons.n, dtype=np.float32)
    olib.gvpm_oracle_sppm_uniforms.restype = None
    olib.gvpm_oracle_sppm_uniforms.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, N.f32p]
    olib.gvpm_oracle_sppm_uniforms(C.byref(cr), rays.n, photons.n, C.byref(medium), C.byref(config),
                                   xi.ctypes.data_as(N.f32p))
    pos = np.ascontiguousarray(photons.pos, dtype=np.float32)
    flux = np.ascontiguousarray(photons.flux, dtype=np.float32)
    d = np.ascontiguousarray(direction, dtype=np.float32).reshape(-1)
    depth = np.ascontiguousarray(photons.depth, dtype=np.uint8)
    out = np.zeros(rays.n * 3, dtype=np.float32)
    rc = lib.ref_fn_sppm_bre_gather(pos.ctypes.data_as(N.f32p), d.ctypes.data_as(N.f32p), flux.ctypes.data_as(N.f32p),
                                    depth.ctypes.data_as(N.u8p), photons.n, C.byref(cr), rays.n, C.byref(medium),
                                    C.byref(config), radius, xi.ctypes.data_as(N.f32p), out.ctypes.data_as(N.f32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_sppm_bre_gather refused the input: {rc}")
    return out.reshape(rays.n, 3)


def sppm_planes_gather(planes, rays, medium, config):
    """sppm's primal plane functor (PhotonPlaneQuery::operator(), photonmapper/plane_struct.h:238-256) over every (camera
    beam, plane) pair in plane order.  Returns (Li [n_rays, 3] without the camera beam's weight, counts [n_rays, 2])."""
    lib = load()
    cp, cr = planes.as_c(), rays.as_c()
    out = np.zeros(rays.n * 3, dtype=np.float32)
    counts = np.zeros(rays.n * 2, dtype=np.uint32)
    rc = lib.ref_fn_sppm_planes_gather(C.byref(cp), planes.n, C.byref(cr), rays.n, C.byref(medium), C.byref(config),
                                       out.ctypes.data_as(N.f32p), counts.ctypes.data_as(N.u32p))
    if rc != 0:
        raise RuntimeError(f"ref_fn_sppm_planes_gather refused the input: {rc}")
    return out.reshape(rays.n, 3), counts.reshape(rays.n, 2)


def bre_pass(photons, rays, medium, config, tri, radius, threads=1, begin=0, end=None):
    """The whole G-BRE gather pass on the reference's own code: GPhotonMap::build + GradientBeamRadianceEstimator +
    bre->query + VolumeGradientBREQuery per camera segment (gvpm.cpp:994-1042), rays [begin, end).  Returns
    (out [m, 27], functor calls [m], (kd build ms, hierarchy ms, gather ms))."""
    lib = load()
    end = rays.n if end is None else end
    cph, cr = photons.as_c(), rays.as_c()
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    m = end - begin
    out = np.zeros(m * 27, dtype=np.float32)
    counts = np.zeros(m, dtype=np.uint32)
    times = (C.c_double * 3)()
    rc = lib.ref_fn_bre_pass(C.byref(cph), photons.n, C.byref(cr), begin, end, C.byref(medium), C.byref(config),
                             tri.ctypes.data_as(N.f32p), tri.size // 9, radius, threads, out.ctypes.data_as(N.f32p),
                             counts.ctypes.data_as(N.u32p), times)
    if rc != 0:
        raise RuntimeError(f"ref_fn_bre_pass refused the input: {rc}")
    return out.reshape(m, 27), counts, tuple(times)


class BrePass:
    """The reference's own G-BRE structures over one photon set, kept between gathers (bench.py's reference arm):
    GPhotonMap::build + GradientBeamRadianceEstimator once, then bre->query + VolumeGradientBREQuery per sampled ray."""

    def __init__(self, photons, medium, config, tri, radius, threads=1):
        self.lib = load()
        cph = photons.as_c()
        tri = np.ascontiguousarray(tri, dtype=np.float32)
        times = (C.c_double * 3)()
        self.h = self.lib.ref_fn_bre_open(C.byref(cph), photons.n, C.byref(medium), C.byref(config),
                                          tri.ctypes.data_as(N.f32p), tri.size // 9, radius, threads, times)
        if not self.h:
            raise RuntimeError("ref_fn_bre_open refused the input")
        self.records_ms, self.kd_build_ms, self.hierarchy_ms = times[0], times[1], times[2]

    def run(self, rays, threads=1, begin=0, end=None, want_out=True):
        end = rays.n if end is None else end
        cr = rays.as_c()
        m = end - begin
        out = np.zeros(m * 27, dtype=np.float32) if want_out else None
        counts = np.zeros(m, dtype=np.uint32)
        ms = C.c_double(0)
        rc = self.lib.ref_fn_bre_run(self.h, C.byref(cr), begin, end, threads,
                                     out.ctypes.data_as(N.f32p) if want_out else None, counts.ctypes.data_as(N.u32p),
                                     C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"ref_fn_bre_run refused the input: {rc}")
        return (out.reshape(m, 27) if want_out else None), counts, ms.value

    def close(self):
        if self.h:
            self.lib.ref_fn_bre_close(self.h)
            self.h = None


def beams_pass(beams, rays, medium, config, tri, radius, threads=1):
    """The whole G-Beams pass on the reference's own code: SubBeamBVH<LTPhotonBeam> + BeamGradRadianceQuery per camera
    segment (gvpm.cpp:893-941).  Returns (out [n_rays, 27], accepted calls [n_rays], (build ms, gather ms))."""
    lib = load()
    xi = beam_uniforms(beams, rays, medium, config)
    cb, cr = beams.as_c(), rays.as_c()
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    out = np.zeros(rays.n * 27, dtype=np.float32)
    counts = np.zeros(rays.n * 2, dtype=np.uint32)
    times = (C.c_double * 2)()
    rc = lib.ref_fn_beams_pass(C.byref(cb), beams.n, C.byref(cr), rays.n, C.byref(medium), C.byref(config),
                               tri.ctypes.data_as(N.f32p), tri.size // 9, radius, xi.ctypes.data_as(N.f32p), threads,
                               out.ctypes.data_as(N.f32p), counts.ctypes.data_as(N.u32p), times)
    if rc != 0:
        raise RuntimeError(f"ref_fn_beams_pass refused the input: {rc}")
    return out.reshape(rays.n, 27), counts.reshape(rays.n, 2)[:, 0], tuple(times)


def planes_pass(planes, rays, medium, config, threads=1):
    """The whole G-Planes pass on the reference's own code: PhotonPlaneBVH<LTPhotonPlane> + PlaneGradRadianceQuery."""
    lib = load()
    cp, cr = planes.as_c(), rays.as_c()
    out = np.zeros(rays.n * 27, dtype=np.float32)
    times = (C.c_double * 2)()
    rc = lib.ref_fn_planes_pass(C.byref(cp), planes.n, C.byref(cr), rays.n, C.byref(medium), C.byref(config), threads,
                                out.ctypes.data_as(N.f32p), times)
    if rc != 0:
        raise RuntimeError(f"ref_fn_planes_pass refused the input: {rc}")
    return out.reshape(rays.n, 27), tuple(times)


def vpm_pass(photons, rays, samples, medium, config, tri, nb_camera_samples, threads=1):
    """The whole G-VPM pass on the reference's own code: GPhotonMap::build + evaluate (PointKDTree range query) +
    VolumeGradientDistanceQuery per distance sample, folded per pixel."""
    lib = load()
    cph, cr, cs = photons.as_c(), rays.as_c(), samples.as_c()
    tri = np.ascontiguousarray(tri, dtype=np.float32)
    out = np.zeros(rays.n * 27, dtype=np.float32)
    mvol = np.zeros(rays.n, dtype=np.float32)
    times = (C.c_double * 2)()
    rc = lib.ref_fn_vpm_pass(C.byref(cph), photons.n, C.byref(cr), rays.n, C.byref(cs), samples.n, C.byref(medium),
                             C.byref(config), tri.ctypes.data_as(N.f32p), tri.size // 9, nb_camera_samples, threads,
                             out.ctypes.data_as(N.f32p), mvol.ctypes.data_as(N.f32p), times)
    if rc != 0:
        raise RuntimeError(f"ref_fn_vpm_pass refused the input: {rc}")
    return out.reshape(rays.n, 27), mvol, tuple(times)


class TechniquePass:
    """The reference's own structures of one technique kept between gathers (bench.py's technique reference arms).
    kind "beams": SubBeamBVH<LTPhotonBeam> + BeamGradRadianceQuery (sampler draws from the C ABI's hash, evaluated in the
    harness); "planes": PhotonPlaneBVH<LTPhotonPlane> + PlaneGradRadianceQuery; "vpm": GPhotonMap (PointKDTree range query) +
    VolumeGradientDistanceQuery."""

    def __init__(self, kind, prims, medium, config, tri=None, radius=0.0, threads=1):
        self.lib, self.kind = load(), kind
        times = (C.c_double * 2)()
        cp = prims.as_c()
        tri = np.ascontiguousarray(tri if tri is not None else np.zeros(0), dtype=np.float32)
        if kind == "beams":
            self.h = self.lib.ref_fn_beams_open(C.byref(cp), prims.n, C.byref(medium), C.byref(config),
                                                tri.ctypes.data_as(N.f32p), tri.size // 9, radius, times)
        elif kind == "planes":
            self.h = self.lib.ref_fn_planes_open(C.byref(cp), prims.n, C.byref(medium), C.byref(config), times)
        elif kind == "vpm":
            self.h = self.lib.ref_fn_vpm_open(C.byref(cp), prims.n, C.byref(medium), C.byref(config),
                                              tri.ctypes.data_as(N.f32p), tri.size // 9, threads, times)
        else:
            raise ValueError(kind)
        if not self.h:
            raise RuntimeError(f"the reference harness refused the {kind} input")
        self.records_ms, self.build_ms = times[0], times[1]

    def run(self, rays, threads=1, samples=None, nb_camera_samples=0, want_out=True):
        """-> (out [n_rays, 27] or None, gather ms)"""
        cr = rays.as_c()
        out = np.zeros(rays.n * 27, dtype=np.float32) if want_out else None
        op = out.ctypes.data_as(N.f32p) if want_out else None
        ms = C.c_double(0)
        if self.kind == "beams":
            rc = self.lib.ref_fn_beams_run(self.h, C.byref(cr), rays.n, None, threads, op, None, C.byref(ms))
        elif self.kind == "planes":
            rc = self.lib.ref_fn_planes_run(self.h, C.byref(cr), rays.n, threads, op, C.byref(ms))
        else:
            cs = samples.as_c()
            rc = self.lib.ref_fn_vpm_run(self.h, C.byref(cr), rays.n, C.byref(cs), samples.n, nb_camera_samples, threads,
                                         op, None, C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"the reference harness refused the rays: {rc}")
        return (out.reshape(rays.n, 27) if want_out else None), ms.value

    def close(self):
        if self.h:
            self.lib.ref_fn_tech_close(self.h)
            self.h = None


def shim_roundtrip(kind, records, medium, config, radius=0.0):
    """Flattened records -> the reference objects the harness rebuilds -> the reference-side shim
    (gvpm_b200/host/gvpm_mitsuba_shim.hpp: flattenPhotonMap / appendLightPathBeams / appendGatherPoint) -> records again."""
    lib = load()
    out = type(records)(records.n)
    ci, co = records.as_c(), out.as_c()
    if kind == "photons":
        rc = lib.ref_fn_shim_photons(C.byref(ci), records.n, C.byref(medium), C.byref(config), C.byref(co))
    elif kind == "beams":
        rc = lib.ref_fn_shim_beams(C.byref(ci), records.n, C.byref(medium), C.byref(config), radius, C.byref(co))
    elif kind == "rays":
        rc = lib.ref_fn_shim_rays(C.byref(ci), records.n, C.byref(medium), C.byref(config), C.byref(co))
    else:
        raise ValueError(kind)
    if rc != 0:
        raise RuntimeError(f"shim round trip ({kind}) failed: {rc}")
    return out


def shim_medium(medium):
    """gvpm_shim::flattenMedium on the reference's HomogeneousMedium built from the record."""
    lib = load()
    out = type(medium)()
    if lib.ref_fn_shim_medium(C.byref(medium), C.byref(out)) != 0:
        raise RuntimeError("shim medium round trip failed")
    return out
