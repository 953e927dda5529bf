"""Shared test scaffolding: seeded synthetic cases and comparison helpers."""
import numpy as np

import gvpm_b200 as g


class Case:
    pass


def make_case(n_photons=20000, w=48, h=32, scale=1.0, seed=0xC0FFEE, phase="isotropic", hg_g=0.0, block=32,
              perturb=True, **cfg_kw):
    c = Case()
    c.medium = g.make_medium(phase=phase, g=hg_g)
    c.photons, c.n_paths = g.synth_photons(n_photons, c.medium, seed=seed, threads=4)
    c.rays = g.synth_rays(w, h, seed=seed + 1, block=block)
    c.tri = g.synth_occluders()
    c.config = g.make_config(w, h, **cfg_kw)
    c.radius = g.bre_radius(scale)
    c.w, c.h = w, h
    if perturb:
        # non-trivial camera-side weights so swapped/ignored fields cannot cancel out
        rng = np.random.default_rng(seed)
        c.rays.eye_contrib[:] = rng.uniform(0.5, 1.5, c.rays.eye_contrib.shape).astype(np.float32)
        c.rays.off_eye[:] = rng.uniform(0.5, 1.5, c.rays.off_eye.shape).astype(np.float32)
        c.rays.off_sensor[:] = rng.uniform(0.7, 1.3, c.rays.off_sensor.shape).astype(np.float32)
    return c


def gpu_context(case, device=0):
    from gvpm_b200.api import Context
    ctx = Context(device)
    ctx.set_medium(case.medium)
    ctx.set_config(case.config)
    ctx.set_occluders(case.tri)
    ctx.upload_photons(case.photons)
    ctx.build_points(case.radius)
    ctx.upload_rays(case.rays)
    return ctx


def rel_err(got, ref):
    """Per-value relative error with an absolute floor of 1e-3 of the largest magnitude (values that
    are numerically zero).  All accumulated terms are non-negative, so there is no cancellation."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    floor = 1e-3 * np.abs(ref).max() if ref.size else 0.0
    return np.abs(got - ref) / np.maximum(np.abs(ref), floor if floor > 0 else 1.0)


def assert_radiance_close(got, ref, rtol=1e-4, what=""):
    """north_star tolerance: primal and gradient radiance within 1e-4 relative (fp32)."""
    err = rel_err(got, ref)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, (f"{what}: max relative error {worst:.3e} > {rtol} at "
                           f"{np.unravel_index(err.argmax(), err.shape)}")
    return worst
