"""GPU parity of the sppm primal photon-beam entry (gvpm_gather_sppm_beams; beams.h:29-223,
sppm.cpp:823-860) against the CPU oracle, through the C ABI, for the four beam x beam techniques.  Bar: per-ray
counts and accepted-beam lists bit-exact, radiance within 1e-4 relative (fp32)."""
import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import records as R

pytestmark = pytest.mark.gpu
TECHNIQUES = ["beam1d", "beam3d_naive", "beam3d_egsr", "beam3d"]


def _case(n_beams=6000, w=40, h=24, scale=3.0, seed=5, **kw):
    kw.setdefault("rng_seed", 4321)
    kw.setdefault("max_depth", -1)
    c = H.make_case(n_photons=64, w=w, h=h, scale=scale, **kw)
    c.beams, c.n_beam_paths = R.synth_beams(n_beams, c.medium, seed=seed, threads=4)
    return c


def _ctx(c):
    from gvpm_b200.api import Context
    ctx = Context(0)
    ctx.set_medium(c.medium)
    ctx.set_config(c.config)
    ctx.set_occluders(c.tri)
    ctx.upload_beams(c.beams)
    ctx.build_beams(c.radius)
    ctx.upload_rays(c.rays)
    return ctx


def _check(c, tech, what, ctx=None):
    from oracle import binding as ob
    ref = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, c.radius, tech, neighbours=True)
    own = ctx is None
    ctx = ctx or _ctx(c)
    out, counts = ctx.gather_sppm_beams(tech)
    out_fast, _ = ctx.gather_sppm_beams(tech, counts=False)
    offsets, idx = ctx.dump_neighbours_sppm_beams(tech)
    np.testing.assert_array_equal(counts, ref.counts)
    np.testing.assert_array_equal(offsets, ref.offsets)
    np.testing.assert_array_equal(idx, ref.idx)
    H.assert_radiance_close(out, ref.out, 1e-4, what)
    H.assert_radiance_close(out_fast, ref.out, 1e-4, what + " (no counts)")
    if own:
        ctx.close()
    return ref


@pytest.mark.parametrize("tech", TECHNIQUES)
@pytest.mark.parametrize("kw", [{}, {"max_depth": 5, "min_depth": 3}, {"long_beams": True}])
def test_sppm_beams_match_oracle(built, tech, kw):
    c = _case(**kw)
    ref = _check(c, tech, f"sppm beams {tech} {kw}")
    assert ref.counts[:, 0].sum() > 2000
    if kw.get("max_depth"):
        assert ref.counts[:, 1].sum() < ref.counts[:, 0].sum()


def test_sppm_beams_hg_small_radius_one_context(built):
    """All four techniques from one context / one hierarchy (HG phase, many more beams than pixels)."""
    c = _case(n_beams=60000, w=32, h=24, scale=0.7, phase="hg", hg_g=0.5)
    ctx = _ctx(c)
    for tech in TECHNIQUES:
        _check(c, tech, f"sppm beams hg {tech}", ctx)
    ctx.close()


def test_sppm_beams_subbeam_split_matches_reference_rule(built):
    """The sub-beam table of gvpm_build_beams is the SubBeamBVH constructor's (beams_accel.h:98-124): the naive
    technique, whose estimate depends on the cuts, agrees with the oracle only if it is."""
    c = _case(n_beams=500, scale=8.0)
    ref = _check(c, "beam3d_naive", "naive wide kernel")
    assert ref.counts[:, 0].max() > 5


def test_sppm_beams_errors_and_empty(built):
    from gvpm_b200.api import Context, GvpmError
    c = _case(n_beams=16)
    ctx = _ctx(c)
    with pytest.raises(GvpmError):
        ctx.gather_sppm_beams(7)
    ctx.upload_beams(c.beams.take(np.zeros(0, dtype=np.int64)))
    ctx.build_beams(c.radius)
    out, counts = ctx.gather_sppm_beams("beam3d")
    assert not out.any() and not counts.any()
    ctx.close()


@pytest.mark.parametrize("tech_id,tech", [(5, "beam3d"), (3, "beam3d_naive"), (1, "bre3d")])
def test_sppm_host_mirror_two_iterations(built, tech_id, tech):
    """gvpm_host::SPPMVolumeGatherB200 (volumePhotonBeamPass / volumePhotonPassBRE, sppm.cpp:765-1001) == oracle gather
    + the reference's per-pixel sum over camera beams, 1 / shotParticles, APA running mean and radius reduction."""
    import ctypes as C
    import os
    import gvpm_b200 as g
    from gvpm_b200 import _native as N
    from oracle import binding as ob
    from test_abi_and_host import ROOT, SppmHostParams
    hl = C.CDLL(os.path.join(ROOT, "gvpm_b200", "host", "libgvpm_host.so"))
    hl.gvpm_host_sppm_create.restype = C.c_void_p
    hl.gvpm_host_sppm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(SppmHostParams), C.POINTER(N.Medium),
                                         C.c_float, C.c_char_p, C.c_size_t]
    hl.gvpm_host_sppm_beam_pass.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.BeamSoA), C.c_size_t, C.POINTER(N.RaySoA),
                                            C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_sppm_bre_pass.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.PhotonSoA), C.c_size_t, C.POINTER(N.RaySoA),
                                           C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_sppm_radius.restype = C.c_float
    hl.gvpm_host_sppm_radius.argtypes = [C.c_void_p]
    hl.gvpm_host_sppm_scale.restype = C.c_double
    hl.gvpm_host_sppm_scale.argtypes = [C.c_void_p]
    hl.gvpm_host_sppm_flux_vol.restype = N.f32p
    hl.gvpm_host_sppm_flux_vol.argtypes = [C.c_void_p]
    hl.gvpm_host_sppm_destroy.argtypes = [C.c_void_p]
    w, h, scale0, shot = 32, 24, 3.0, 5000
    err = C.create_string_buffer(512)
    p = SppmHostParams(maxDepth=8, minDepth=0, alpha=0.7, initialScaleVolume=scale0, volTechnique=tech_id, rngSeed=4321,
                       forceAPA=b"")
    med = g.make_medium()
    hd = hl.gvpm_host_sppm_create(0, w, h, C.byref(p), C.byref(med), g.records.SYNTH_BSPHERE_R, err, 512)
    assert hd, err.value
    flux_ref = np.zeros((h, w, 3), dtype=np.float32)
    scale = scale0
    for it in (1, 2):
        c = _case(n_beams=3000, w=w, h=h, scale=scale, seed=7 * it, max_depth=8, sppm_primal=True)
        radius = hl.gvpm_host_sppm_radius(hd)
        assert abs(radius - g.bre_radius(scale)) <= 1e-9
        cr = c.rays.as_c()
        if tech == "bre3d":
            c.photons, _ = g.synth_photons(20000, c.medium, seed=3 * it, threads=4)
            cph = c.photons.as_c()
            rc = hl.gvpm_host_sppm_bre_pass(hd, it, C.byref(cph), c.photons.n, C.byref(cr), c.rays.n, shot, err, 512)
            ref = ob.sppm_bre_gather(c.photons, c.rays, c.medium, c.config, radius, mode="brute")
        else:
            cb = c.beams.as_c()
            rc = hl.gvpm_host_sppm_beam_pass(hd, it, C.byref(cb), c.beams.n, C.byref(cr), c.rays.n, shot, err, 512)
            ref = ob.sppm_beams_gather(c.beams, c.rays, c.medium, c.config, radius, tech)
        assert rc == 0, err.value
        assert ref.counts[:, 1].sum() > 500
        pix = np.zeros((h, w, 3), dtype=np.float32)
        np.add.at(pix, (c.rays.py, c.rays.px), ref.out)
        flux_ref = (flux_ref * np.float32(it - 1) + pix / np.float32(shot)) / np.float32(it)
        scale *= ((it - 1 + 0.7) / it) ** (1 / 3)
        assert abs(hl.gvpm_host_sppm_scale(hd) - scale) < 1e-12
    got = np.ctypeslib.as_array(hl.gvpm_host_sppm_flux_vol(hd), shape=(h * w * 3,)).reshape(h, w, 3).copy()
    H.assert_radiance_close(got, flux_ref, 1e-4, f"sppm host mirror {tech}")
    hl.gvpm_host_sppm_destroy(hd)
