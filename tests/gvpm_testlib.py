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


def make_plane_case(n_planes=1200, w=40, h=24, seed=0xC0FFEE, phase="isotropic", hg_g=0.0, inside=True, sheet=False,
                    **cfg_kw):
    """G-Planes 0D case: beams of the seeded light paths extended to planes by the host
    mirror of transformBeam; sensor inside the medium at (0.5, 0.5, 0.05) (gvpm.cpp:785-787)."""
    from gvpm_b200 import records as R
    c = Case()
    c.medium = g.make_medium(phase=phase, g=hg_g)
    c.beams, c.n_paths = R.synth_beams(n_planes, c.medium, seed=seed, threads=4)
    c.planes = R.synth_planes(c.beams, c.medium, seed=seed + 7)
    if sheet:
        # collimated emitter: origins squeezed into a 0.02-wide sheet around x = 0.5, first edges along -y
        o = c.planes.view("origin")
        o[:, 0] = 0.5 + (o[:, 0] - 0.5) * 0.02
        c.planes.length1[:] *= 0.05
    if inside:
        c.rays = g.synth_rays(w, h, seed=seed + 1, cam_dist=-0.05, cover=0.45)
    else:
        c.rays = g.synth_rays(w, h, seed=seed + 1)
    rng = np.random.default_rng(seed)
    c.rays.off_sensor[:] = rng.uniform(0.7, 1.3, c.rays.off_sensor.shape).astype(np.float32)
    c.config = g.make_config(w, h, **cfg_kw)
    c.w, c.h = w, h
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


def rel_err_per_ray(got, ref):
    """Error of every value relative to ITS OWN ray's scale: rows are rays (27 floats: primal, 4 shifted, 4 weighted
    spectra), the denominator is max(|value|, 1e-3 * that ray's largest |entry|).  A dim pixel is therefore held
    to the same 1e-4 as the brightest one (the global floor of rel_err would check it absolutely); the only values
    excused are those below 0.1 % of their own ray's primal scale, i.e. sums of non-negative terms that are
    numerically zero next to the ray's other entries.  Returns (err, fraction of non-zero values under that floor)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape and got.ndim == 2
    scale = np.abs(ref).max(axis=1, keepdims=True)
    floor = np.where(scale > 0, 1e-3 * scale, 1.0)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), floor)
    nz = np.abs(ref) > 0
    under = float((nz & (np.abs(ref) < floor)).sum()) / max(1, int(nz.sum()))
    return err, under


def assert_radiance_close(got, ref, rtol=1e-4, what=""):
    """north_star tolerance: primal and gradient radiance within 1e-4 relative (fp32): globally floored AND per ray."""
    err = rel_err(got, ref)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, (f"{what}: max relative error {worst:.3e} > {rtol} at "
                           f"{np.unravel_index(err.argmax(), err.shape)}")
    g2, r2 = np.asarray(got), np.asarray(ref)
    if g2.ndim == 2 and g2.size:
        e2, under = rel_err_per_ray(g2, r2)
        w2 = float(e2.max())
        assert w2 <= rtol, (f"{what}: per-ray relative error {w2:.3e} > {rtol} at "
                            f"{np.unravel_index(e2.argmax(), e2.shape)} ({under:.2%} of the non-zero values are under their ray's floor)")
        assert under < 0.05, f"{what}: {under:.2%} of the non-zero values sit under their ray's 1e-3 floor"
        worst = max(worst, w2)
    return worst
