"""GPU tests (through the C ABI): golden fixture, the C++ host-side mirror of the reference drivers
(computeVolumeGradientPhotonBRE / scaleVolumeAPA / computeGradient), size-independent properties at
larger sizes, and the sharded path's equivalence with the single-GPU result."""
import ctypes as C
import os

import numpy as np
import pytest

import gvpm_testlib as H
from gvpm_b200 import _native as N
from gvpm_b200 import shard

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_matches_golden_fixture(built):
    z = np.load(os.path.join(ROOT, "tests", "golden", "bre_small.npz"))
    c = H.make_case(**{k: z[k].item() for k in ("n_photons", "w", "h", "scale", "seed")})
    ctx = H.gpu_context(c)
    out, counts = ctx.gather_bre()
    offsets, idx = ctx.dump_neighbours_bre()
    np.testing.assert_array_equal(counts, z["counts"])
    np.testing.assert_array_equal(offsets, z["offsets"])
    np.testing.assert_array_equal(idx, z["idx"])
    H.assert_radiance_close(out, z["out"], 1e-4, "golden")
    ctx.close()


def gradient_reference(acc, w, h, use_abs):
    """numpy restatement of computeGradient (gvpm.cpp:1205-1306), volume terms of an APA estimator."""
    a = acc.reshape(h, w, 9, 3)
    S, W = a[:, :, 1:5], a[:, :, 5:9]
    L, R, T, B = 0, 1, 2, 3
    gx = S[:, :, R] - W[:, :, R]
    gx[:, :-1] += W[:, 1:, L] - S[:, 1:, L]
    gy = S[:, :, T] - W[:, :, T]
    gy[:-1] += W[1:, :, B] - S[1:, :, B]
    if use_abs:
        gx, gy = np.abs(gx), np.abs(gy)
    return a[:, :, 0].copy(), gx, gy


@pytest.mark.parametrize("use_abs", [False, True])
def test_compute_gradient_kernel(built, use_abs):
    from gvpm_b200.api import Context
    rng = np.random.default_rng(3)
    w, h = 37, 23
    acc = rng.normal(size=(h * w * 27)).astype(np.float32)
    ctx = Context(0)
    thr, gx, gy = ctx.compute_gradient(acc, w, h, use_abs)
    rt, rx, ry = gradient_reference(acc.copy(), w, h, use_abs)
    np.testing.assert_array_equal(thr, rt)
    np.testing.assert_allclose(gx, rx, rtol=0, atol=1e-6)
    np.testing.assert_allclose(gy, ry, rtol=0, atol=1e-6)
    ctx.close()


def reuse_primal_reference(acc, w, h, inv_emitted):
    """numpy restatement of the reusePrimal throughput (gvpm.cpp:503-532), same order of additions"""
    a = acc.reshape(h, w, 9, 3)
    S, W = a[:, :, 1:5], a[:, :, 5:9]
    L, R, T, B = 0, 1, 2, 3
    t = np.zeros((h, w, 3), dtype=np.float32)
    t[:, :-1] += S[:, 1:, L]
    t[:, 1:] += S[:, :-1, R]
    t[:-1] += S[1:, :, B]
    t[1:] += S[:-1, :, T]
    t += ((W[:, :, B] + W[:, :, T]) + W[:, :, R]) + W[:, :, L]
    return (t * np.float32(0.25)) * np.float32(inv_emitted)


@pytest.mark.parametrize("inv_emitted", [1.0, 1.0 / 250000.0])
def test_reuse_primal_throughput(built, inv_emitted):
    from gvpm_b200.api import Context
    rng = np.random.default_rng(5)
    w, h = 29, 17
    acc = rng.uniform(0.0, 2.0, size=(h * w * 27)).astype(np.float32)
    ctx = Context(0)
    thr, gx, gy = ctx.compute_gradient(acc, w, h, False, reuse_primal=True, inv_emitted=inv_emitted)
    rt, rx, ry = gradient_reference(acc.copy(), w, h, False)
    np.testing.assert_array_equal(thr, reuse_primal_reference(acc.copy(), w, h, inv_emitted))
    np.testing.assert_allclose(gx, rx, rtol=0, atol=1e-6)     # the gradients do not change
    np.testing.assert_allclose(gy, ry, rtol=0, atol=1e-6)
    ctx.close()


def test_host_driver_two_iterations(built):
    """gvpm_host::VolumeGatherB200 == oracle gather + the reference's normalisation, APA running mean
    (gvpm.cpp:1054-1069) and radius reduction (:181-215) over two iterations."""
    import __graft_entry__ as ge
    from oracle import binding as ob
    from test_abi_and_host import HostParams, host_params
    ge.build()
    hl = C.CDLL(os.path.join(ROOT, "gvpm_b200", "host", "libgvpm_host.so"))
    hl.gvpm_host_create.restype = C.c_void_p
    hl.gvpm_host_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(HostParams), C.POINTER(N.Medium), C.c_float,
                                    N.f32p, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_bre_iteration.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.PhotonSoA), C.c_size_t,
                                           C.POINTER(N.RaySoA), C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    hl.gvpm_host_gradient.argtypes = [C.c_void_p, N.f32p, N.f32p, N.f32p, C.c_int, C.c_char_p, C.c_size_t]
    hl.gvpm_host_scale.restype = C.c_double
    hl.gvpm_host_scale.argtypes = [C.c_void_p]
    hl.gvpm_host_radius.restype = C.c_float
    hl.gvpm_host_radius.argtypes = [C.c_void_p]
    hl.gvpm_host_accumulators.restype = N.f32p
    hl.gvpm_host_accumulators.argtypes = [C.c_void_p]
    hl.gvpm_host_destroy.argtypes = [C.c_void_p]

    import gvpm_b200 as g
    w, h = 40, 24
    scale0 = 2.0
    err = C.create_string_buffer(512)
    p = host_params(initialScaleVolume=scale0)
    c0 = H.make_case(n_photons=20000, w=w, h=h, scale=scale0, seed=11)
    hd = hl.gvpm_host_create(0, w, h, C.byref(p), C.byref(c0.medium), g.records.SYNTH_BSPHERE_R,
                             c0.tri.ctypes.data_as(N.f32p), c0.tri.size // 9, err, 512)
    assert hd, err.value
    acc_ref = np.zeros((h, w, 27), dtype=np.float32)
    scale = scale0
    for it in (1, 2):
        c = H.make_case(n_photons=20000, w=w, h=h, scale=scale, seed=11 * it)
        radius = hl.gvpm_host_radius(hd)
        assert abs(radius - g.bre_radius(scale)) <= 1e-9
        cph, cr = c.photons.as_c(), c.rays.as_c()
        rc = hl.gvpm_host_bre_iteration(hd, it, C.byref(cph), c.photons.n, C.byref(cr), c.rays.n, c.n_paths, err, 512)
        assert rc == 0, err.value
        ref = ob.bre_gather(c.photons, c.rays, c.medium, c.config, c.tri, radius, mode="brute")
        img = shard.to_image(ref.out, c.rays.px, c.rays.py, w, h)
        acc_ref = (acc_ref * np.float32(it - 1) + img / np.float32(c.n_paths)) / np.float32(it)
        scale *= ((it - 1 + 0.7) / it) ** (1.0 / 3.0)
        assert abs(hl.gvpm_host_scale(hd) - scale) < 1e-12
    acc = np.ctypeslib.as_array(hl.gvpm_host_accumulators(hd), shape=(h, w, 27)).copy()
    H.assert_radiance_close(acc, acc_ref, 1e-4, "host driver accumulators")
    thr, gx, gy = (np.zeros((h, w, 3), dtype=np.float32) for _ in range(3))
    assert hl.gvpm_host_gradient(hd, thr.ctypes.data_as(N.f32p), gx.ctypes.data_as(N.f32p),
                                 gy.ctypes.data_as(N.f32p), 0, err, 512) == 0
    rt, rx, ry = gradient_reference(acc.copy(), w, h, False)
    np.testing.assert_array_equal(thr, rt)
    np.testing.assert_allclose(gx, rx, atol=1e-6 * max(1.0, np.abs(rx).max()))
    hl.gvpm_host_destroy(hd)


def test_properties_at_scale(built):
    """Larger than the oracle comfortably brute-forces: size-independent properties.
    (1) sharding invariance: gathering the image in 4 tile shards gives the bit-identical per-ray result;
    (2) linearity in the photon flux: scaling every flux by 2 scales all 27 outputs by exactly 2;
    (3) photon-order invariance of the neighbour sets: a random permutation of the photon array gives the
        same per-ray counts; (4) a sampled subset of rays matches the oracle (kd-tree mode)."""
    from oracle import binding as ob
    c = H.make_case(n_photons=400_000, w=256, h=160, scale=0.6)
    ctx = H.gpu_context(c)
    out, counts = ctx.gather_bre()
    assert counts[:, 0].sum() > 100_000
    # (1)
    parts, lists = [], []
    for r in range(4):
        idx = shard.local_indices(c.rays.px, c.rays.py, c.w, 4, r)
        ctx.upload_rays(c.rays.take(idx))
        o, _ = ctx.gather_bre()
        parts.append(o)
        lists.append(idx)
    # (per-ray sums are folded by a segmented warp scan + float atomics whose grouping depends on where
    #  a ray's pairs land in the pair list: equal up to fp32 summation order, not bitwise)
    H.assert_radiance_close(shard.assemble(parts, lists, c.rays.n), out, 1e-5, "sharding invariance")
    # (2)
    ph2 = c.photons.copy()
    ph2.flux[:] *= np.float32(2)
    ph2.prefix_flux[:] *= np.float32(2)
    ctx.upload_photons(ph2)
    ctx.build_points(c.radius)
    ctx.upload_rays(c.rays)
    out2, counts2 = ctx.gather_bre()
    np.testing.assert_array_equal(counts2, counts)
    H.assert_radiance_close(out2, out * np.float32(2), 1e-5, "linearity in the flux")
    # (3)
    perm = np.random.default_rng(5).permutation(c.photons.n)
    ctx.upload_photons(c.photons.take(perm))
    ctx.build_points(c.radius)
    out3, counts3 = ctx.gather_bre()
    np.testing.assert_array_equal(counts3, counts)
    H.assert_radiance_close(out3, out, 1e-4, "photon permutation")
    # (4)
    sel = np.arange(0, c.rays.n, 97)
    ref = ob.bre_gather(c.photons, c.rays.take(sel), c.medium, c.config, c.tri, c.radius, mode="kdtree")
    same = (ref.counts == counts[sel]).all(axis=1)
    assert same.mean() > 0.999  # the reference tree's Epsilon sliver at the ray end
    H.assert_radiance_close(out[sel][same], ref.out[same], 1e-4, "sampled rays vs oracle kd-tree")
    ctx.close()


@pytest.mark.parametrize("w,h", [(48, 32), (1024, 576)])
def test_pipelined_host_gather(built, w, h):
    """gvpm_gather_bre_host (chunked upload / gather / download on three streams) == upload_rays + gather_bre.
    1024x576 rays span several chunks."""
    c = H.make_case(n_photons=100_000, w=w, h=h, scale=1.0 if w > 100 else 3.0)
    ctx = H.gpu_context(c)
    out, counts = ctx.gather_bre()
    assert counts[:, 1].sum() > 1000
    piped = ctx.gather_bre_host(c.rays)
    H.assert_radiance_close(piped, out, 1e-5, "pipelined vs plain")
    # a second iteration through the same context (staging reuse across the copy / compute streams)
    piped2 = ctx.gather_bre_host(c.rays)
    H.assert_radiance_close(piped2, out, 1e-5, "pipelined, second call")
    ctx.close()
