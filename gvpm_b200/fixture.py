"""On-disk fixture of one gather iteration (gvpm_b200/host/gvpm_fixture.hpp documents the format).

A fixture holds the flattened inputs of one G-BRE iteration (medium, config, radius, occluders, photons, rays) and,
optionally, what its producer's own gather returned (27 floats per ray, per-ray neighbour lists).  The producer is
either this repository (tests) or a dump hook inside a real Mitsuba build of the reference, in
which case `tools/check_fixture.py` checks the CUDA path against the reference's own numbers.
"""
import struct

import numpy as np

from . import _native as N
from . import records as R

MAGIC = b"GVPMFIX1"
_DTYPES = {0: np.float32, 1: np.uint8, 2: np.uint32, 3: np.int32, 4: np.float64, 5: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}
_CONFIG_FIELDS = ["max_depth", "min_depth", "lighting_mode", "use_mis", "use_shift_null", "path_set",
                  "power_heuristic", "kernel_3d", "film_w", "film_h", "shadow_maxt_scale", "epsilon", "long_beams",
                  "rng_seed", "beam_kernel_1d", "sppm_primal"]


def read_sections(path):
    """-> {name: 1-D numpy array}"""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != MAGIC:
        raise ValueError(f"{path}: not a gvpm fixture")
    n_sections, = struct.unpack_from("<I", blob, 8)
    pos, out = 16, {}
    for _ in range(n_sections):
        name = blob[pos:pos + 32].split(b"\0", 1)[0].decode()
        dtype, _, count = struct.unpack_from("<IIQ", blob, pos + 32)
        if dtype not in _DTYPES:
            raise ValueError(f"{path}: section {name}: unknown dtype {dtype}")
        dt = np.dtype(_DTYPES[dtype])
        pos += 48
        nbytes = count * dt.itemsize
        if pos + nbytes > len(blob):
            raise ValueError(f"{path}: section {name} is truncated")
        out[name] = np.frombuffer(blob, dtype=dt, count=count, offset=pos).copy()
        pos += (nbytes + 7) & ~7
    return out


def write_sections(path, sections):
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", len(sections), 0))
        for name, a in sections.items():
            a = np.ascontiguousarray(a).reshape(-1)
            f.write(name.encode()[:31].ljust(32, b"\0") + struct.pack("<IIQ", _CODES[a.dtype], 0, a.size))
            f.write(a.tobytes())
            f.write(b"\0" * (-a.nbytes % 8))


class Fixture:
    """medium, config, radius, tri [n,9], photons (PhotonSet), rays (RaySet), expected_out [n_rays,27] or None,
    expected_offsets / expected_idx or None, producer (str)"""


def load(path):
    s = read_sections(path)
    fx = Fixture()
    m = s["medium"]
    fx.medium = N.Medium()
    for i in range(3):
        fx.medium.sigma_s[i], fx.medium.sigma_a[i] = float(m[i]), float(m[3 + i])
    fx.medium.phase_type, fx.medium.hg_g, fx.medium.sampling_weight = int(m[6]), float(m[7]), float(m[8])
    fx.config = N.Config()
    for name, v in zip(_CONFIG_FIELDS, s["config"]):
        cur = getattr(fx.config, name)
        setattr(fx.config, name, float(v) if isinstance(cur, float) else int(v))
    fx.radius = float(s["radius"][0])
    fx.tri = s["occluders"].reshape(-1, 9)
    n = s["photon.path_id"].size
    fx.photons = R.PhotonSet(n, **{name: s["photon." + name] for name, _, _ in R._PHOTON_FIELDS})
    q = s["ray.px"].size
    fx.rays = R.RaySet(q, **{name: s["ray." + name] for name, _, _ in R._RAY_FIELDS})
    fx.expected_out = s["expected.out"].reshape(q, N.GVPM_OUT_FLOATS) if "expected.out" in s else None
    fx.expected_offsets = s.get("expected.nbr_offsets")
    fx.expected_idx = s.get("expected.nbr_idx")
    fx.producer = s["meta.producer"].tobytes().decode() if "meta.producer" in s else ""
    return fx


def save(path, medium, config, radius, tri, photons, rays, expected_out=None, expected_offsets=None,
         expected_idx=None, producer="gvpm_b200"):
    s = {"medium": np.array(list(medium.sigma_s) + list(medium.sigma_a) +
                            [medium.phase_type, medium.hg_g, medium.sampling_weight], dtype=np.float32),
         "config": np.array([getattr(config, f) for f in _CONFIG_FIELDS], dtype=np.float64),
         "radius": np.array([radius], dtype=np.float32),
         "occluders": np.ascontiguousarray(tri, dtype=np.float32).reshape(-1)}
    for name, _, _ in R._PHOTON_FIELDS:
        s["photon." + name] = getattr(photons, name)
    for name, _, _ in R._RAY_FIELDS:
        s["ray." + name] = getattr(rays, name)
    if expected_out is not None:
        s["expected.out"] = np.ascontiguousarray(expected_out, dtype=np.float32).reshape(-1)
    if expected_offsets is not None:
        s["expected.nbr_offsets"] = np.ascontiguousarray(expected_offsets, dtype=np.uint64)
        s["expected.nbr_idx"] = np.ascontiguousarray(expected_idx, dtype=np.uint32)
    s["meta.producer"] = np.frombuffer(producer.encode(), dtype=np.uint8)
    write_sections(path, s)
