"""Seeded inputs of the radiometric pins (tests/test_oracle_physics_pin.py, tests/golden/make_physics_golden.py) and the
runner that evaluates them on one side (reference code or oracle restatement)."""
import numpy as np


def _unit(v):
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def inputs(n=4000, seed=20260):
    rng = np.random.default_rng(seed)
    d = {}
    # medium: segment lengths from 1e-6 to 60 (the transmittance cut-off at 1e-20 is crossed), mint > 0 in half of them
    d["mint"] = np.where(rng.random(n) < 0.5, 0.0, rng.uniform(0, 0.3, n)).astype(np.float32)
    d["maxt"] = (d["mint"] + np.exp(rng.uniform(np.log(1e-6), np.log(60.0), n))).astype(np.float32)
    d["wi"], d["wo"] = _unit(rng.normal(size=(n, 3))), _unit(rng.normal(size=(n, 3)))
    # parents of the three in-scope types; normals axis-aligned (walls) and generic
    ptype = rng.integers(0, 3, n).astype(np.uint8)
    axis = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n)] * rng.choice([-1.0, 1.0], (n, 1)).astype(np.float32)
    normal = np.where(rng.random((n, 1)) < 0.7, axis, _unit(rng.normal(size=(n, 3)))).astype(np.float32)
    parent = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    pred = (parent + _unit(rng.normal(size=(n, 3))) * rng.uniform(0.05, 1.0, (n, 1))).astype(np.float32)
    new_d = _unit(rng.normal(size=(n, 3)))
    # most reconnections leave the surface on the lit side, some graze or point into it (side tests, zero cosines)
    flip = (np.einsum("ij,ij->i", normal, new_d) < 0) & (rng.random(n) < 0.8) & (ptype != 2)
    new_d[flip] *= -1
    graze = rng.random(n) < 0.03
    new_d[graze] = _unit(np.cross(normal[graze], _unit(rng.normal(size=(int(graze.sum()), 3)))) + 1e-12)
    ppdf = rng.uniform(0.01, 30.0, n).astype(np.float32)
    ppdf[rng.random(n) < 0.02] = 0.0
    d["rec"] = dict(parent_type=ptype, parent_pos=parent, pred_pos=pred, parent_n=normal,
                    albedo=rng.uniform(0.05, 0.9, (n, 3)).astype(np.float32), parent_pdf=ppdf,
                    edge_pdf=rng.uniform(0.05, 3.0, n).astype(np.float32),
                    rr_weight=np.where(rng.random(n) < 0.5, 1.0, rng.uniform(1.0, 4.0, n)).astype(np.float32),
                    new_d=new_d, new_len=np.exp(rng.uniform(np.log(1e-3), np.log(3.0), n)).astype(np.float32))
    d["normal"] = normal
    d["albedo1"] = np.array([0.63, 0.065, 0.05], np.float32)
    return d


MEDIA = [  # (sigma_s, sigma_a, mediumSamplingWeight)
    ((1.6, 1.6, 1.6), (0.4, 0.4, 0.4), 1.0),
    ((0.3, 0.3, 0.3), (0.05, 0.05, 0.05), 0.8),
    ((7.0, 7.0, 7.0), (1.0, 1.0, 1.0), 0.5),
]
PHASES = [(0, 0.0), (1, 0.4), (1, -0.7), (1, 0.05)]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).copy()


def run(side, d):
    """every pinned routine on the inputs `d`; floats returned as their bit patterns"""
    out = {}
    for k, (ss, sa, w) in enumerate(MEDIA):
        T, ps, pf = side.medium_eval(ss, sa, w, d["mint"], d["maxt"])
        out[f"med{k}_T_bits"], out[f"med{k}_ps_bits"], out[f"med{k}_pf_bits"] = bits(T), bits(ps), bits(pf)
    for k, (kind, g) in enumerate(PHASES):
        ev, pd = side.phase(kind, g, d["wi"], d["wo"])
        out[f"ph{k}_eval_bits"], out[f"ph{k}_pdf_bits"] = bits(ev), bits(pd)
    for k, ((ss, sa, w), (kind, g)) in enumerate([(MEDIA[0], PHASES[0]), (MEDIA[1], PHASES[1]), (MEDIA[2], PHASES[2])]):
        ok, thr, pdf = side.diffuse_reconnection(ss, sa, w, kind, g, d["rec"])
        out[f"rc{k}_ok"], out[f"rc{k}_thr_bits"], out[f"rc{k}_pdf_bits"] = ok, bits(thr), bits(pdf)
    return out
