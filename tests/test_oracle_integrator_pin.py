"""Rows a18 / a19 of SURVEY.md §8 pinned to the REFERENCE'S OWN compiled code: GPMIntegrator::scaleVolumeAPA
(gvpm/gvpm.cpp:181-215, the per-iteration kernel reduction) and GPMIntegrator::computeGradient (:1205-1304, gradient images
from the per-pixel accumulators), plus SPPMIntegrator::scaleVolumeAPA (photonmapper/sppm.cpp:255-290).  oracle/_ref/libgvpm_integrator_ref.so is built from /root/reference (oracle/Makefile,
target `integrator_ref`; oracle/ref_integrator.cpp includes gvpm.cpp where it lies and calls the two member functions on raw
storage); tests/golden/integrator_pins.npz holds its outputs (tests/golden/make_integrator_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import integrator_pin_cases as cases
from oracle import integrator_binding as ib
from test_abi_and_host import HostParams, SppmHostParams, _host, host_params
from test_gpu_host_and_gradient import gradient_reference

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "integrator_pins.npz")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN))


@pytest.mark.parametrize("key", list(cases.SCHEDULES))
def test_host_schedule_equals_reference(golden, key):
    """The host mirror's scaleVolumeAPA.  Evaluated in the reference build's Float (gvpm_host_scale_apa_f32: alpha and
    globalScaleVolume are Floats, so the ratio is formed in single precision and the product with cbrt / sqrt is rounded
    back to it) it reproduces the SINGLE_PRECISION reference bit for bit over 60 iterations; the double form the drivers
    keep (the reference's DOUBLE_PRECISION arithmetic) stays within 3e-6 of it."""
    tech, force, k3, alpha, s0 = cases.SCHEDULES[key]
    want = golden[f"scale_{key}"]
    assert len(np.unique(want)) == cases.N_ITER and want[-1] < 0.95 * s0       # a real schedule
    h = _host()
    h.gvpm_host_scale_apa_f32.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(HostParams), C.c_char_p, C.c_size_t]
    h.gvpm_host_scale_apa.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(HostParams), C.c_char_p, C.c_size_t]
    p = host_params(volTechnique=cases.HOST_TECHNIQUE[tech], forceAPA=force.encode(), alpha=alpha, initialScaleVolume=s0,
                    use3DKernelReduction=int(k3))
    err = C.create_string_buffer(256)
    sf, sd = C.c_float(s0), C.c_double(s0)
    got_f, got_d = [], []
    for it in range(1, cases.N_ITER + 1):
        assert h.gvpm_host_scale_apa_f32(C.byref(sf), it, C.byref(p), err, 256) == 0
        assert h.gvpm_host_scale_apa(C.byref(sd), it, C.byref(p), err, 256) == 0
        got_f.append(sf.value)
        got_d.append(sd.value)
    np.testing.assert_array_equal(np.array(got_f, np.float32), want)
    np.testing.assert_allclose(np.array(got_d), want.astype(np.float64), rtol=3e-6, atol=0)


@pytest.mark.parametrize("key", list(cases.SPPM_SCHEDULES))
def test_sppm_host_schedule_equals_reference(golden, key):
    """SPPMIntegrator::scaleVolumeAPA (sppm.cpp:255-290) against the sppm host mirror: float form bit for bit, double form
    within 3e-6."""
    tech, force, alpha, s0 = cases.SPPM_SCHEDULES[key]
    want = golden[f"sppmscale_{key}"]
    h = _host()
    h.gvpm_host_sppm_scale_apa_f32.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(SppmHostParams), C.c_char_p, C.c_size_t]
    h.gvpm_host_sppm_scale_apa.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(SppmHostParams), C.c_char_p, C.c_size_t]
    p = SppmHostParams(maxDepth=-1, minDepth=0, alpha=alpha, initialScaleVolume=s0,
                       volTechnique=cases.SPPM_HOST_TECHNIQUE[tech], rngSeed=0, forceAPA=force.encode())
    err = C.create_string_buffer(256)
    sf, sd = C.c_float(s0), C.c_double(s0)
    got_f, got_d = [], []
    for it in range(1, cases.N_ITER + 1):
        assert h.gvpm_host_sppm_scale_apa_f32(C.byref(sf), it, C.byref(p), err, 256) == 0
        assert h.gvpm_host_sppm_scale_apa(C.byref(sd), it, C.byref(p), err, 256) == 0
        got_f.append(sf.value)
        got_d.append(sd.value)
    np.testing.assert_array_equal(np.array(got_f, np.float32), want)
    np.testing.assert_allclose(np.array(got_d), want.astype(np.float64), rtol=3e-6, atol=0)


@pytest.mark.parametrize("key", list(cases.GRADIENTS))
def test_gradient_restatement_equals_reference(golden, key):
    """computeGradient, volume terms: the numpy restatement the CUDA kernel is tested against
    (tests/test_gpu_host_and_gradient.py: gradient_reference) is the reference's function bit for bit; estimators that are
    not APA ("distance") divide by the emitted count (Spectrum /= Float multiplies by the reciprocal)."""
    tech, use_abs, emitted = cases.GRADIENTS[key]
    _, gx, gy = gradient_reference(cases.accumulators(), cases.W, cases.H, use_abs)
    if tech == "distance":
        recip = np.float32(1) / np.float32(emitted)
        gx, gy = gx * recip, gy * recip
    np.testing.assert_array_equal(gx.view(np.uint32), golden[f"grad_{key}_gx"].view(np.uint32))
    np.testing.assert_array_equal(gy.view(np.uint32), golden[f"grad_{key}_gy"].view(np.uint32))
    assert np.count_nonzero(gx) > 0.99 * gx.size


@pytest.mark.skipif(not (ib.have_ref() or os.path.isdir(ib.REFERENCE_ROOT)), reason="reference tree / prebuilt library absent")
def test_golden_vectors_are_what_the_reference_computes_now(golden):
    if not ib.have_ref():
        assert ib.build_ref()
    for key, (tech, force, k3, alpha, s0) in cases.SCHEDULES.items():
        np.testing.assert_array_equal(ib.scale_volume_apa(s0, cases.N_ITER, alpha, tech, force, k3), golden[f"scale_{key}"])
    for key, (tech, force, alpha, s0) in cases.SPPM_SCHEDULES.items():
        np.testing.assert_array_equal(ib.sppm_scale_volume_apa(s0, cases.N_ITER, alpha, tech, force), golden[f"sppmscale_{key}"])
    acc = cases.accumulators()
    for key, (tech, use_abs, emitted) in cases.GRADIENTS.items():
        gx, gy = ib.compute_gradient(acc, cases.W, cases.H, use_abs, tech, emitted)
        np.testing.assert_array_equal(gx, golden[f"grad_{key}_gx"])
        np.testing.assert_array_equal(gy, golden[f"grad_{key}_gy"])
