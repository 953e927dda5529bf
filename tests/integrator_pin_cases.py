"""Cases of the integrator pins (tests/test_oracle_integrator_pin.py, tests/golden/make_integrator_golden.py)."""
import numpy as np

N_ITER = 60
W, H = 37, 23
# key -> (reference technique name, forceAPA, use3DKernelReduction, alpha, initialScaleVolume)
SCHEDULES = {
    "bre3d": ("bre3d", "", False, 0.7, 1.0),
    "bre2d": ("bre2d", "", False, 0.7, 1.0),
    "distance": ("distance", "", False, 0.7, 0.35),
    "beam3d": ("beam3d", "", False, 0.5, 2.0),
    "beam1d": ("beam1d", "", False, 0.7, 1.0),
    "plane0d": ("plane0d", "", False, 0.9, 1.0),
    "beam1d_3d_reduction": ("beam1d", "", True, 0.7, 1.0),
    "bre3d_force_1d": ("bre3d", "1D", False, 0.7, 1.0),
    "bre2d_force_3d": ("bre2d", "3D", False, 0.7, 1.0),
    "beam1d_force_2d": ("beam1d", "2D", False, 0.7, 1.0),
}
# the host mirror's technique numbering (gvpm_b200/host/gvpm_host.hpp:34)
HOST_TECHNIQUE = {"bre2d": 0, "bre3d": 1, "distance": 2, "beam3d": 3, "plane0d": 4, "beam1d": 5}
# sppm: key -> (reference technique name, forceAPA, alpha, initialScaleVolume); the sppm mirror's technique numbering
# (gvpm_b200/host/gvpm_host.hpp:425)
SPPM_SCHEDULES = {
    "bre3d": ("bre3d", "", 0.7, 0.3),
    "bre2d": ("bre2d", "", 0.7, 1.0),
    "beam1d": ("beam1d", "", 0.6, 1.0),
    "beam3d_naive": ("beam3d_naive", "", 0.7, 1.0),
    "beam3d_egsr": ("beam3d_egsr", "", 0.7, 2.0),
    "beam3d": ("beam3d", "", 0.8, 1.0),
    "beam1d_force_3d": ("beam1d", "3D", 0.7, 1.0),
    "bre3d_force_2d": ("bre3d", "2D", 0.7, 1.0),
}
SPPM_HOST_TECHNIQUE = {"bre2d": 0, "bre3d": 1, "beam1d": 2, "beam3d_naive": 3, "beam3d_egsr": 4, "beam3d": 5}
# key -> (reference technique name, useAbs, m_totalEmittedVolume)
GRADIENTS = {
    "bre3d": ("bre3d", False, 250000),
    "bre3d_abs": ("bre3d", True, 250000),
    "beam3d_abs": ("beam3d", True, 1),
    "distance": ("distance", False, 250000),      # not an APA estimator: the gradients are divided by the emitted count
    "distance_abs": ("distance", True, 1000),
}


def accumulators():
    rng = np.random.default_rng(3)
    return rng.normal(size=(H * W * 27)).astype(np.float32)
