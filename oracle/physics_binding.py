"""ctypes binding of oracle/_ref/libgvpm_physics_ref.so (the REFERENCE'S OWN medium / phase / BSDF / emitter plugins and
shift_diffuse.cpp, built by `make -C oracle physics_ref` from /root/reference) and of the matching
`gvpm_oracle_pin_*` entry points of the oracle restatement.

TEST INFRASTRUCTURE: used by tests/test_oracle_physics_pin.py and tests/golden/make_physics_golden.py only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libgvpm_physics_ref.so")
ORACLE_LIB = os.path.join(_HERE, "libgvpm_oracle.so")
REFERENCE_ROOT = os.environ.get("GVPM_REFERENCE_ROOT", "/root/reference")
f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)


def build_ref():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", _HERE, "physics_ref", f"REF={REFERENCE_ROOT}"])
    return os.path.exists(REF_LIB)


def have_ref():
    return os.path.exists(REF_LIB)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(f32p)


class Side:
    """kind "ref": the reference's code; "oracle": the restatement.  Same flat arrays in and out."""

    def __init__(self, kind):
        self.kind = kind
        if kind == "ref":
            if not have_ref():
                raise FileNotFoundError(REF_LIB)
            self.lib, self.pre = C.CDLL(REF_LIB), "ref_phys_"
        else:
            if not os.path.exists(ORACLE_LIB):
                subprocess.check_call(["make", "-s", "-C", _HERE])
            self.lib, self.pre = C.CDLL(ORACLE_LIB), "gvpm_oracle_pin_"

    def medium_eval(self, sig_s, sig_a, weight, mint, maxt):
        mint, maxt = _f(mint), _f(maxt)
        n = mint.size
        T, ps, pf = np.zeros(3 * n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        fn = getattr(self.lib, self.pre + "medium_eval")
        fn.argtypes = [f32p, f32p, C.c_float, C.c_size_t, f32p, f32p, f32p, f32p, f32p]
        fn.restype = None
        fn(_p(_f(sig_s)), _p(_f(sig_a)), weight, n, _p(mint), _p(maxt), _p(T), _p(ps), _p(pf))
        return T.reshape(n, 3), ps, pf

    def phase(self, kind, g, wi, wo):
        wi, wo = _f(wi).reshape(-1), _f(wo).reshape(-1)
        n = wi.size // 3
        ev, pd = np.zeros(n, np.float32), np.zeros(n, np.float32)
        fn = getattr(self.lib, self.pre + "phase")
        fn.argtypes = [C.c_int, C.c_float, C.c_size_t, f32p, f32p, f32p, f32p]
        fn.restype = None
        fn(kind, g, n, _p(wi), _p(wo), _p(ev), _p(pd))
        return ev, pd

    def diffuse_bsdf(self, albedo, normal, wi, wo):
        assert self.kind == "ref"
        normal, wi, wo = _f(normal).reshape(-1), _f(wi).reshape(-1), _f(wo).reshape(-1)
        n = wi.size // 3
        ev, pd = np.zeros(3 * n, np.float32), np.zeros(n, np.float32)
        fn = self.lib.ref_phys_diffuse_bsdf
        fn.argtypes = [f32p, C.c_size_t, f32p, f32p, f32p, f32p, f32p]
        fn.restype = None
        fn(_p(_f(albedo)), n, _p(normal), _p(wi), _p(wo), _p(ev), _p(pd))
        return ev.reshape(n, 3), pd

    def area_emitter(self, normal, d):
        assert self.kind == "ref"
        normal, d = _f(normal).reshape(-1), _f(d).reshape(-1)
        n = d.size // 3
        ev, pd = np.zeros(3 * n, np.float32), np.zeros(n, np.float32)
        fn = self.lib.ref_phys_area_emitter
        fn.argtypes = [C.c_size_t, f32p, f32p, f32p, f32p]
        fn.restype = None
        fn(n, _p(normal), _p(d), _p(ev), _p(pd))
        return ev.reshape(n, 3), pd

    def diffuse_reconnection(self, sig_s, sig_a, weight, phase, g, rec):
        """rec: dict of flat arrays parent_type (u8), parent_pos, pred_pos, parent_n, albedo, parent_pdf, edge_pdf,
        rr_weight, new_d, new_len -> (ok u8 [n], throughput [n,3], pdf [n])"""
        n = rec["parent_type"].size
        ok, thr, pdf = np.zeros(n, np.uint8), np.zeros(3 * n, np.float32), np.zeros(n, np.float32)
        fn = getattr(self.lib, self.pre + "diffuse_reconnection")
        fn.argtypes = [f32p, f32p, C.c_float, C.c_int, C.c_float, C.c_size_t, u8p] + [f32p] * 9 + [u8p, f32p, f32p]
        fn.restype = None
        a = {k: (np.ascontiguousarray(v, dtype=np.uint8) if k == "parent_type" else _f(v).reshape(-1)) for k, v in rec.items()}
        fn(_p(_f(sig_s)), _p(_f(sig_a)), weight, phase, g, n, a["parent_type"].ctypes.data_as(u8p), _p(a["parent_pos"]),
           _p(a["pred_pos"]), _p(a["parent_n"]), _p(a["albedo"]), _p(a["parent_pdf"]), _p(a["edge_pdf"]),
           _p(a["rr_weight"]), _p(a["new_d"]), _p(a["new_len"]), ok.ctypes.data_as(u8p), _p(thr), _p(pdf))
        return ok, thr.reshape(n, 3), pdf
