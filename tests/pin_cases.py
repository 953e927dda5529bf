"""Seeded inputs + one `run(side, inputs)` for the reference pin (tests/test_oracle_ref_pin.py,
tests/golden/make_ref_golden.py).  `side` is oracle.ref_binding.Side("ref") — the reference's own code — or
Side("oracle") — the restatement; both produce a dict name -> ndarray that must be identical."""
import numpy as np


def _unit(v):
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def inputs(size="small", seed=20261017):
    rng = np.random.default_rng(seed)
    n, n_rays, m = {"small": (4000, 300, 4000), "large": (60000, 3000, 200000)}[size]
    inp = {}
    # photons: uniform + a dense cluster + a slab of points quantised to a coarse grid (ties in the kd build)
    a = rng.random((n // 2, 3), dtype=np.float32)
    b = (0.5 + 0.03 * rng.standard_normal((n // 4, 3))).astype(np.float32)
    c = (np.floor(rng.random((n - n // 2 - n // 4, 3)) * 16) / 16).astype(np.float32)
    inp["pos"] = np.concatenate([a, b, c]).astype(np.float32)
    inp["radius"] = np.float32(0.03 if size == "small" else 0.012)
    # camera rays: pinhole outside the box looking in + rays starting inside; some end inside the medium
    o = np.tile(np.array([0.5, 0.5, -1.5], np.float32), (n_rays, 1))
    inside = rng.random(n_rays) < 0.3
    o[inside] = rng.random((int(inside.sum()), 3), dtype=np.float32)
    tgt = rng.random((n_rays, 3), dtype=np.float32)
    inp["ray_o"] = o
    inp["ray_d"] = _unit(tgt - o)
    inp["ray_d"][0] = (0, 0, 1)            # axis-parallel rays: zero direction components in the slab test
    inp["ray_d"][1] = (1, 0, 0)
    inp["ray_mint"] = np.full(n_rays, 1e-4, np.float32)
    inp["ray_maxt"] = (0.2 + 2.5 * rng.random(n_rays)).astype(np.float32)
    # VPM query points with per-sample radii
    inp["q"] = rng.random((n_rays, 3), dtype=np.float32)
    inp["q_radius"] = (inp["radius"] * (0.3 + rng.random(n_rays))).astype(np.float32)
    # beams (short segments) and planes in the box
    nb = n // 20
    inp["beam_o"] = rng.random((nb, 3), dtype=np.float32)
    inp["beam_e"] = (inp["beam_o"] + _unit(rng.standard_normal((nb, 3))) *
                     (0.02 + 0.4 * rng.random((nb, 1)))).astype(np.float32)
    inp["pl_ori"] = rng.random((nb, 3), dtype=np.float32)
    inp["pl_w0"] = _unit(rng.standard_normal((nb, 3)))
    inp["pl_w1"] = _unit(rng.standard_normal((nb, 3)))
    inp["pl_len0"] = (0.02 + 0.4 * rng.random(nb)).astype(np.float32)
    inp["pl_len1"] = (0.02 + 0.4 * rng.random(nb)).astype(np.float32)
    # element-wise pair queries: camera segment (as the cylinder) x beam, ray x plane, ray x triangle
    pr = rng.integers(0, n_rays, m)
    pb = rng.integers(0, nb, m)
    inp["pair_ray"], inp["pair_prim"] = pr.astype(np.int64), pb.astype(np.int64)
    tri = rng.random((m, 3, 3), dtype=np.float32)
    tri[: m // 2, 1:] = tri[: m // 2, :1] + 0.3 * (tri[: m // 2, 1:] - 0.5)
    inp["tri"] = tri.reshape(m, 9)
    v = _unit(rng.standard_normal((m, 3)))
    v[:6] = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0, -1, 0], [-1, 0, 0]], np.float32)
    inp["vecs"] = v
    abc = rng.standard_normal((m, 3))
    abc[:10, 0] = 0
    abc[10:14, 1] = 0
    inp["abc"] = abc
    return inp


def run(side, inp):
    """Everything both sides implement, as name -> array."""
    out = {}
    depth, orig, right, leaf, axis = side.kd_layout(inp["pos"])
    out.update(kd_depth=np.int64(depth), kd_orig=orig, kd_right=right, kd_leaf=leaf, kd_axis=axis)
    off, idx, td, d2 = side.bre_visits(inp["pos"], float(inp["radius"]), inp["ray_o"], inp["ray_d"], inp["ray_mint"],
                                       inp["ray_maxt"])
    out.update(bre_off=off, bre_idx=idx, bre_tdisk_bits=td.view(np.uint32), bre_depth=np.int64(d2))
    off, idx = side.range_visits(inp["pos"], inp["q"], inp["q_radius"])
    out.update(range_off=off, range_idx=idx)
    pr, pp = inp["pair_ray"], inp["pair_prim"]
    ro, rd, mint, maxt = inp["ray_o"][pr], inp["ray_d"][pr], inp["ray_mint"][pr], inp["ray_maxt"][pr]
    # cylinderIntersection as BeamKernelRecord::eval calls it (shift_volume_beams.h:203-207): the cylinder is the
    # camera segment re-based at ray(mint), the "view" ray is the beam
    bo, be = inp["beam_o"][pp], inp["beam_e"][pp]
    bd = be - bo
    bl = np.sqrt((bd * bd).sum(1)).astype(np.float32)
    bd = (bd / bl[:, None]).astype(np.float32)
    hit, tn, tf = side.cylinder((ro + rd * mint[:, None]).astype(np.float32), rd, (maxt - mint).astype(np.float32), bo,
                                bd, bl, np.full(len(pr), inp["radius"] * 3, np.float32))
    out.update(cyl_hit=hit, cyl_tnear_bits=np.where(hit, tn, 0).view(np.uint64),
               cyl_tfar_bits=np.where(hit, tf, 0).view(np.uint64))
    hit, o4 = side.plane0d(inp["pl_ori"][pp], inp["pl_w0"][pp], inp["pl_len0"][pp], inp["pl_w1"][pp],
                           inp["pl_len1"][pp], ro, rd, mint, maxt)
    out.update(pl_hit=hit, pl_out_bits=np.where(hit[:, None], o4, 0).astype(np.float32).view(np.uint32))
    # rayIntersectInternal1D on (ray, sub-beam [tmin, tmax]) pairs: half of them with the whole beam as the range
    tmin = np.where(np.arange(len(pr)) % 2 == 0, 0, 0.3 * bl).astype(np.float32)
    tmax = np.where(np.arange(len(pr)) % 2 == 0, bl, 0.7 * bl).astype(np.float32)
    hit, o4 = side.beam1d(bo, be, np.full(len(pr), inp["radius"] * 3, np.float32), ro, rd, mint, maxt, tmin, tmax)
    out.update(b1d_hit=hit, b1d_out_bits=np.where(hit[:, None], o4, 0).astype(np.float32).view(np.uint32))
    out["tri_hit"] = side.triangle_any_hit(inp["tri"], ro, rd, mint, maxt)
    for coh in (0, 1):
        b, c = side.coordsys(inp["vecs"], coh)
        out[f"cs{coh}_b_bits"] = b.view(np.uint32)
        out[f"cs{coh}_c_bits"] = c.view(np.uint32)
    ok, x0, x1 = side.quadratic(inp["abc"])
    out.update(quad_ok=ok, quad_x0_bits=np.where(ok, x0, 0).view(np.uint64), quad_x1_bits=np.where(ok, x1, 0).view(np.uint64))
    return out
