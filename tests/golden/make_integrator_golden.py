"""Regenerates tests/golden/integrator_pins.npz from two member functions of the REFERENCE'S OWN integrator class
(GPMIntegrator::scaleVolumeAPA and GPMIntegrator::computeGradient, gvpm/gvpm.cpp; SPPMIntegrator::scaleVolumeAPA, sppm.cpp; compiled from /root/reference by
`make -C oracle integrator_ref`).  Run in the container that holds the reference tree:
    python tests/golden/make_integrator_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import integrator_pin_cases as cases  # noqa: E402
from oracle import integrator_binding as ib  # noqa: E402

if __name__ == "__main__":
    assert ib.build_ref(), "the reference tree is needed to regenerate the vectors"
    out = {}
    for key, (tech, force, k3, alpha, s0) in cases.SCHEDULES.items():
        out[f"scale_{key}"] = ib.scale_volume_apa(s0, cases.N_ITER, alpha, tech, force, k3)
    for key, (tech, force, alpha, s0) in cases.SPPM_SCHEDULES.items():
        out[f"sppmscale_{key}"] = ib.sppm_scale_volume_apa(s0, cases.N_ITER, alpha, tech, force)
    acc = cases.accumulators()
    for key, (tech, use_abs, emitted) in cases.GRADIENTS.items():
        gx, gy = ib.compute_gradient(acc, cases.W, cases.H, use_abs, tech, emitted)
        out[f"grad_{key}_gx"], out[f"grad_{key}_gy"] = gx, gy
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "integrator_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
