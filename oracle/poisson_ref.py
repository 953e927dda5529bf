"""ctypes binding of oracle/_ref/libgvpm_poisson_ref.so: the reference's own screened-Poisson solver
(src/integrators/poisson_solver, built by `make -C oracle poisson_ref`).  TEST INFRASTRUCTURE: it generates the golden
vectors of tests/golden/poisson_*.npz, checks the CUDA solver live when present, and is the CPU arm the
reconstruction is timed against (`kind = "reference"`)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libgvpm_poisson_ref.so")
# the same solver linked with the reference's own CUDA backend (`make -C oracle poisson_ref_cuda`): backend="CUDA"
LIB_CUDA = os.path.join(HERE, "_ref", "libgvpm_poisson_ref_cuda.so")
_lib = None
_lib_cuda = None
f32p = C.POINTER(C.c_float)


def available():
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        lib.gvpm_ref_poisson_solve.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, f32p, C.c_float, C.c_int, C.c_float,
                                               C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_char_p, f32p]
        lib.gvpm_ref_poisson_preset.argtypes = [C.c_char_p, C.POINTER(C.c_int), f32p, f32p, C.POINTER(C.c_int),
                                                C.POINTER(C.c_int), C.POINTER(C.c_int), f32p, f32p]
        _lib = lib
    return _lib


def cuda_available():
    return os.path.exists(LIB_CUDA)


def load_cuda():
    global _lib_cuda
    if _lib_cuda is None:
        lib = C.CDLL(LIB_CUDA)
        lib.gvpm_ref_poisson_solve.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, f32p, C.c_float, C.c_int, C.c_float,
                                               C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_char_p, f32p]
        _lib_cuda = lib
    return _lib_cuda


def preset(name):
    """Solver::Params::setConfigPreset (Solver.cpp:90-158) -> dict"""
    a, d, e, f = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    b, c, g, al = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    if load().gvpm_ref_poisson_preset(name.encode(), C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e),
                                      C.byref(f), C.byref(g), C.byref(al)) != 0:
        raise ValueError(name)
    return dict(irls_iter_max=a.value, irls_reg_init=b.value, irls_reg_iter=c.value, cg_iter_max=d.value,
                cg_iter_check=e.value, cg_precond=f.value, cg_tolerance=g.value, alpha=al.value)


def solve(throughput, dx, dy, direct=None, backend="Naive", **params):
    """throughput / dx / dy / direct: [h, w, 3] float32 (throughput and direct may be None) -> reconstruction.
    backend: "Naive" (Backend.cpp, one thread), "OpenMP" (BackendOpenMP.cpp) or "CUDA" (BackendCUDA.cu, needs a GPU)."""
    h, w, _ = dx.shape
    p = preset("L2D")
    p.update(params)
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (throughput, dx, dy, direct)]
    ptr = [None if a is None else a.ctypes.data_as(f32p) for a in arrs]
    out = np.zeros((h, w, 3), dtype=np.float32)
    lib = load_cuda() if backend == "CUDA" else load()
    rc = lib.gvpm_ref_poisson_solve(w, h, ptr[0], ptr[1], ptr[2], ptr[3], p["alpha"], p["irls_iter_max"],
                                       p["irls_reg_init"], p["irls_reg_iter"], p["cg_iter_max"], p["cg_iter_check"],
                                       p["cg_precond"], p["cg_tolerance"], backend.encode(), out.ctypes.data_as(f32p))
    if rc != 0:
        raise RuntimeError("gvpm_ref_poisson_solve failed")
    return out
