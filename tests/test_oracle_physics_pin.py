"""Pins the RADIOMETRIC half of the oracle against the REFERENCE'S OWN code (rows a9 / a17 of SURVEY.md §8).

oracle/_ref/libgvpm_physics_ref.so is built from /root/reference (oracle/Makefile, target `physics_ref`): the
reference's HomogeneousMedium::eval (src/medium/homogeneous.cpp:432-513), the isotropic and Henyey-Greenstein phase
functions (src/phase/isotropic.cpp:76, src/phase/hg.cpp:107-110), the diffuse BSDF (src/bsdfs/diffuse.cpp:110-127), the
area emitter's directional term (src/emitters/area.cpp:132-150) and gvpm's diffuseReconnection
(gvpm/shift/operation/shift_diffuse.cpp:11-134) driven with PathVertex / PathEdge records of the three in-scope parent
types.  tests/golden/physics_pins.npz holds its outputs on seeded inputs (tests/golden/make_physics_golden.py), so the
pin also holds where the reference tree is absent.  Everything is compared BIT-EXACTLY (floats as integer bits)."""
import os

import numpy as np
import pytest

import physics_pin_cases as cases
from oracle import physics_binding as pb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "physics_pins.npz")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def data():
    return cases.inputs()


@pytest.fixture(scope="module")
def oracle_out(data):
    return cases.run(pb.Side("oracle"), data)


def _same(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = int(np.count_nonzero(a != b))
    assert bad == 0, f"{what}: {bad} of {a.size} entries differ from the reference"


def test_golden_is_not_trivial(golden):
    f = lambda k: golden[k].view(np.float32)
    assert (f("med2_T_bits") == 0).any() and (f("med0_T_bits") > 0.9).any()        # the 1e-20 cut-off and thin segments
    assert 0.3 < golden["rc0_ok"].mean() < 0.99                                      # failing and succeeding reconnections
    assert (f("rc0_thr_bits") == 0).any() and (f("rc0_pdf_bits") > 0).any()
    assert len(np.unique(golden["ph1_eval_bits"])) > 1000                            # HG really depends on the directions


KEYS = ([f"med{k}_{q}_bits" for k in range(3) for q in ("T", "ps", "pf")] +
        [f"ph{k}_{q}_bits" for k in range(4) for q in ("eval", "pdf")] +
        [f"rc{k}_{q}" for k in range(3) for q in ("ok", "thr_bits", "pdf_bits")])


@pytest.mark.parametrize("key", KEYS)
def test_oracle_equals_reference_golden(golden, oracle_out, key):
    _same(oracle_out[key], golden[key], key)


def test_bsdf_and_emitter_terms_as_the_oracle_folds_them(golden, data):
    """The oracle has no BSDF / emitter objects: diffuseReconnection uses albedo * (INV_PI * cosO) with cosines taken
    against the stored normal, and INV_PI * max(0, d . n) for the emitter.  Same bits as the reference plugins."""
    n, wi, wo = data["normal"], data["wi"], data["wo"]
    inv_pi = np.float32(0.31830988618379067154)
    dot = lambda a, b: ((a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]).astype(np.float32) + a[:, 2] * b[:, 2]).astype(np.float32)
    cos_i, cos_o = dot(n, wi), dot(n, wo)
    # Frame::toLocal: the z component is dot(n, v) as dot() evaluates it (x*x' + y*y' + z*z')
    live = (cos_i > 0) & (cos_o > 0)
    ev = np.where(live[:, None], data["albedo1"][None, :] * (inv_pi * cos_o)[:, None], np.float32(0)).astype(np.float32)
    pd = np.where(live, inv_pi * cos_o, np.float32(0)).astype(np.float32)
    _same(cases.bits(ev), golden["bsdf_eval_bits"], "diffuse BSDF eval")
    _same(cases.bits(pd), golden["bsdf_pdf_bits"], "diffuse BSDF pdf")
    dp = dot(wo, n)
    dp = np.where(dp < 0, np.float32(0), dp)
    e = (inv_pi * dp).astype(np.float32)
    _same(cases.bits(np.repeat(e[:, None], 3, axis=1)), golden["emit_eval_bits"], "area emitter evalDirection")
    _same(cases.bits(e), golden["emit_pdf_bits"], "area emitter pdfDirection")


@pytest.mark.skipif(not (pb.have_ref() or os.path.isdir(pb.REFERENCE_ROOT)), reason="reference tree / prebuilt library absent")
def test_oracle_equals_live_reference_on_fresh_inputs():
    """not only the committed vectors: other seeds, straight against the compiled reference"""
    assert pb.build_ref()
    ref, ora = pb.Side("ref"), pb.Side("oracle")
    for seed in (1, 2, 3):
        d = cases.inputs(n=3000, seed=seed)
        a, b = cases.run(ref, d), cases.run(ora, d)
        for k in a:
            _same(b[k], a[k], f"seed {seed} {k}")
