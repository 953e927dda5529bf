"""ctypes binding of oracle/_ref/libgvpm_ref.so (the REFERENCE'S OWN code, built by `make -C oracle ref`
from /root/reference) and of the matching `gvpm_oracle_pin_*` entry points of the oracle restatement.

TEST INFRASTRUCTURE: used by tests/test_oracle_ref_pin.py and tests/golden/make_ref_golden.py only.
Both sides take the same flat arrays, so `Side("ref")` and `Side("oracle")` are interchangeable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libgvpm_ref.so")
ORACLE_LIB = os.path.join(_HERE, "libgvpm_oracle.so")
REFERENCE_ROOT = os.environ.get("GVPM_REFERENCE_ROOT", "/root/reference")

f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)


def build_ref():
    """Compile the reference pieces when the reference tree is present (this container); on the GPU box
    the prebuilt .so travels with the snapshot.  Returns True when the library exists afterwards."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={REFERENCE_ROOT}"])
    return os.path.exists(REF_LIB)


def have_ref():
    return os.path.exists(REF_LIB)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Side:
    """One implementation of the pinned building blocks: kind "ref" (reference code) or "oracle"."""

    def __init__(self, kind):
        self.kind = kind
        if kind == "ref":
            if not have_ref():
                raise FileNotFoundError(REF_LIB)
            self.lib = C.CDLL(REF_LIB)
            self.pre = "ref_"
            assert self.lib.ref_float_bytes() == 4
        else:
            if not os.path.exists(ORACLE_LIB):
                subprocess.check_call(["make", "-s", "-C", _HERE])
            self.lib = C.CDLL(ORACLE_LIB)
            self.pre = "gvpm_oracle_pin_"

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.pre + name)
        f.restype = restype
        return f

    # ---- kd-tree -----------------------------------------------------------------------------------
    def kd_layout(self, pos):
        """(depth, orig, right, leaf, axis) of the sliding-midpoint kd-tree over pos [n,3]."""
        pos = _f32(pos)
        n = len(pos)
        orig, right = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        leaf, axis = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        f = self._fn("kd_layout", C.c_int)
        extra = [C.c_int(1)] if self.kind == "ref" else []
        depth = f(_p(pos, f32p), C.c_size_t(n), *extra, _p(orig, u32p), _p(right, u32p), _p(leaf, u8p), _p(axis, u8p))
        return depth, orig, right, leaf, axis

    def _csr(self, call, n_lists, extra_dtypes=()):
        off = np.zeros(n_lists + 1, np.uint64)
        total = call(off, None, [None] * len(extra_dtypes), 0)
        idx = np.zeros(max(total, 1), np.uint32)
        extra = [np.zeros(max(total, 1), dt) for dt in extra_dtypes]
        call(off, idx, extra, total)
        return (off, idx[:total]) + tuple(e[:total] for e in extra)

    def range_visits(self, pos, q, radius):
        """executeQuery visit order: (offsets, idx)."""
        pos, q, radius = _f32(pos), _f32(q), _f32(radius)
        f = self._fn("range_visits", C.c_longlong)

        def call(off, idx, extra, cap):
            return f(_p(pos, f32p), C.c_size_t(len(pos)), _p(q, f32p), _p(radius, f32p), C.c_size_t(len(q)),
                     _p(off, u64p), _p(idx, u32p), C.c_size_t(cap))
        return self._csr(call, len(q))

    def bre_visits(self, pos, radius, o, d, mint, maxt):
        """GradientBeamRadianceEstimator::query functor calls: (offsets, idx, tdisk, depth)."""
        pos, o, d, mint, maxt = _f32(pos), _f32(o), _f32(d), _f32(mint), _f32(maxt)
        f = self._fn("bre_visits", C.c_longlong)
        depth = C.c_int(0)

        def call(off, idx, extra, cap):
            return f(_p(pos, f32p), C.c_size_t(len(pos)), C.c_float(radius), _p(o, f32p), _p(d, f32p), _p(mint, f32p),
                     _p(maxt, f32p), C.c_size_t(len(o)), _p(off, u64p), _p(idx, u32p), _p(extra[0], f32p),
                     C.c_size_t(cap), C.byref(depth))
        off, idx, td = self._csr(call, len(o), (np.float32,))
        return off, idx, td, depth.value

    # ---- reference-only structures (no oracle twin: the oracle's beam / plane gathers are defined by brute force) --
    def subbeam_visits(self, origin, end, radius, o, d, mint, maxt):
        assert self.kind == "ref"
        origin, end, o, d, mint, maxt = _f32(origin), _f32(end), _f32(o), _f32(d), _f32(mint), _f32(maxt)
        f = self._fn("subbeam_visits", C.c_longlong)

        def call(off, idx, extra, cap):
            return f(_p(origin, f32p), _p(end, f32p), C.c_size_t(len(origin)), C.c_float(radius), _p(o, f32p),
                     _p(d, f32p), _p(mint, f32p), _p(maxt, f32p), C.c_size_t(len(o)), _p(off, u64p), _p(idx, u32p),
                     _p(extra[0], f32p), _p(extra[1], f32p), C.c_size_t(cap))
        return self._csr(call, len(o), (np.float32, np.float32))

    def plane_visits(self, ori, w0, len0, w1, len1, o, d, mint, maxt):
        assert self.kind == "ref"
        ori, w0, len0, w1, len1 = _f32(ori), _f32(w0), _f32(len0), _f32(w1), _f32(len1)
        o, d, mint, maxt = _f32(o), _f32(d), _f32(mint), _f32(maxt)
        f = self._fn("plane_visits", C.c_longlong)

        def call(off, idx, extra, cap):
            return f(_p(ori, f32p), _p(w0, f32p), _p(len0, f32p), _p(w1, f32p), _p(len1, f32p), C.c_size_t(len(ori)),
                     _p(o, f32p), _p(d, f32p), _p(mint, f32p), _p(maxt, f32p), C.c_size_t(len(o)), _p(off, u64p),
                     _p(idx, u32p), C.c_size_t(cap))
        return self._csr(call, len(o))

    # ---- intersection routines (element-wise over m queries) ------------------------------------------
    def cylinder(self, co, cd, cmaxt, vo, vd, vmaxt, radius):
        co, cd, cmaxt, vo, vd, vmaxt, radius = map(_f32, (co, cd, cmaxt, vo, vd, vmaxt, radius))
        m = len(co)
        hit, tn, tf = np.zeros(m, np.uint8), np.zeros(m, np.float64), np.zeros(m, np.float64)
        self._fn("cylinder")(_p(co, f32p), _p(cd, f32p), _p(cmaxt, f32p), _p(vo, f32p), _p(vd, f32p), _p(vmaxt, f32p),
                             _p(radius, f32p), C.c_size_t(m), _p(hit, u8p), _p(tn, f64p), _p(tf, f64p))
        return hit.astype(bool), tn, tf

    def plane0d(self, ori, w0, len0, w1, len1, o, d, mint, maxt):
        ori, w0, len0, w1, len1, o, d, mint, maxt = map(_f32, (ori, w0, len0, w1, len1, o, d, mint, maxt))
        m = len(ori)
        hit, out = np.zeros(m, np.uint8), np.zeros((m, 4), np.float32)
        self._fn("plane0d")(_p(ori, f32p), _p(w0, f32p), _p(len0, f32p), _p(w1, f32p), _p(len1, f32p), _p(o, f32p),
                            _p(d, f32p), _p(mint, f32p), _p(maxt, f32p), C.c_size_t(m), _p(hit, u8p), _p(out, f32p))
        return hit.astype(bool), out

    def beam1d(self, origin, end, radius, o, d, mint, maxt, tmin, tmax):
        origin, end, radius, o, d, mint, maxt, tmin, tmax = map(_f32, (origin, end, radius, o, d, mint, maxt, tmin, tmax))
        m = len(origin)
        hit, out = np.zeros(m, np.uint8), np.zeros((m, 4), np.float32)
        self._fn("beam1d")(_p(origin, f32p), _p(end, f32p), _p(radius, f32p), _p(o, f32p), _p(d, f32p), _p(mint, f32p),
                           _p(maxt, f32p), _p(tmin, f32p), _p(tmax, f32p), C.c_size_t(m), _p(hit, u8p), _p(out, f32p))
        return hit.astype(bool), out

    def triangle_any_hit(self, tri, o, d, mint, maxt):
        """hit of ONE triangle per query within [mint, maxt] (the Scene::rayIntersect(ray) any-hit semantics)."""
        tri, o, d, mint, maxt = map(_f32, (tri, o, d, mint, maxt))
        m = len(o)
        hit = np.zeros(m, np.uint8)
        if self.kind == "ref":
            uvt = np.zeros((m, 3), np.float32)
            self._fn("triangle")(_p(tri, f32p), _p(o, f32p), _p(d, f32p), C.c_size_t(m), _p(hit, u8p), _p(uvt, f32p))
            # the interval test of the caller: Triangle::rayIntersect returns the raw t (triangle.h:109-145);
            # TriMesh / kd-tree code accepts t in [mint, maxt]
            return hit.astype(bool) & (uvt[:, 2] >= mint) & (uvt[:, 2] <= maxt)
        self._fn("triangle")(_p(tri, f32p), _p(o, f32p), _p(d, f32p), _p(mint, f32p), _p(maxt, f32p), C.c_size_t(m),
                             _p(hit, u8p))
        return hit.astype(bool)

    def coordsys(self, a, coherent):
        a = _f32(a)
        b, c = np.zeros_like(a), np.zeros_like(a)
        self._fn("coordsys")(_p(a, f32p), C.c_size_t(len(a)), C.c_int(int(coherent)), _p(b, f32p), _p(c, f32p))
        return b, c

    def quadratic(self, abc):
        abc = np.ascontiguousarray(abc, dtype=np.float64)
        m = len(abc)
        ok, x0, x1 = np.zeros(m, np.uint8), np.zeros(m, np.float64), np.zeros(m, np.float64)
        self._fn("quadratic")(_p(abc, f64p), C.c_size_t(m), _p(ok, u8p), _p(x0, f64p), _p(x1, f64p))
        return ok.astype(bool), x0, x1
