"""ctypes binding of oracle/_ref/libgvpm_integrator_ref.so: two member functions of the REFERENCE'S OWN integrator class,
GPMIntegrator::scaleVolumeAPA (gvpm/gvpm.cpp:181-215) and GPMIntegrator::computeGradient (:1205-1304), compiled from
/root/reference by `make -C oracle integrator_ref` (oracle/ref_integrator.cpp includes gvpm.cpp where it lies).

TEST INFRASTRUCTURE: used by tests/test_oracle_integrator_pin.py and tests/golden/make_integrator_golden.py only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libgvpm_integrator_ref.so")
REFERENCE_ROOT = os.environ.get("GVPM_REFERENCE_ROOT", "/root/reference")
f32p = C.POINTER(C.c_float)
# EVolumeTechnique, src/integrators/volume_utils.h:12-21
TECHNIQUES = {"bre2d": 0, "bre3d": 1, "distance": 2, "beam1d": 3, "beam3d_naive": 4, "beam3d_egsr": 5, "beam3d": 6, "plane0d": 7}
_lib = None


def build_ref():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", _HERE, "integrator_ref", f"REF={REFERENCE_ROOT}"])
    return os.path.exists(REF_LIB)


def have_ref():
    return os.path.exists(REF_LIB)


def load():
    global _lib
    if _lib is None:
        if not have_ref():
            raise FileNotFoundError(REF_LIB)
        _lib = C.CDLL(REF_LIB)
        _lib.ref_int_scale_volume_apa.restype = None
        _lib.ref_int_scale_volume_apa.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int, C.c_char_p, C.c_int, f32p]
        _lib.ref_int_compute_gradient.restype = None
        _lib.ref_int_compute_gradient.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, f32p, f32p]
        _lib.ref_int_sppm_scale_volume_apa.restype = None
        _lib.ref_int_sppm_scale_volume_apa.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int, C.c_char_p, f32p]
    return _lib


def scale_volume_apa(scale0, n, alpha, technique, force_apa="", use_3d_kernel_reduction=False):
    """globalScaleVolume after scaleVolumeAPA(it), it = 1 .. n (SINGLE_PRECISION build: a float)."""
    out = np.zeros(n, dtype=np.float32)
    load().ref_int_scale_volume_apa(scale0, n, alpha, TECHNIQUES[technique], force_apa.encode(),
                                    int(use_3d_kernel_reduction), out.ctypes.data_as(f32p))
    return out


def compute_gradient(acc, w, h, use_abs, technique="bre3d", total_emitted_volume=1):
    """-> (gx, gy) [h, w, 3] from the per-pixel accumulators acc [h * w * 27] (volume-only rendering)."""
    acc = np.ascontiguousarray(acc, dtype=np.float32).reshape(-1)
    gx = np.zeros(w * h * 3, dtype=np.float32)
    gy = np.zeros(w * h * 3, dtype=np.float32)
    load().ref_int_compute_gradient(acc.ctypes.data_as(f32p), w, h, int(use_abs), TECHNIQUES[technique],
                                    total_emitted_volume, gx.ctypes.data_as(f32p), gy.ctypes.data_as(f32p))
    return gx.reshape(h, w, 3), gy.reshape(h, w, 3)


def sppm_scale_volume_apa(scale0, n, alpha, technique, force_apa=""):
    """SPPMIntegrator::scaleVolumeAPA (sppm.cpp:255-290): globalScaleVolume after iteration it = 1 .. n."""
    out = np.zeros(n, dtype=np.float32)
    load().ref_int_sppm_scale_volume_apa(scale0, n, alpha, TECHNIQUES[technique], force_apa.encode(), out.ctypes.data_as(f32p))
    return out
