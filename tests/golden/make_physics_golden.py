"""Regenerates tests/golden/physics_pins.npz from the REFERENCE'S OWN radiometric code (oracle/_ref/libgvpm_physics_ref.so,
built from /root/reference by `make -C oracle physics_ref`).  Run in the container that holds the reference tree:
    python tests/golden/make_physics_golden.py
The vectors let tests/test_oracle_physics_pin.py hold the oracle to the reference where the tree is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import physics_pin_cases as cases  # noqa: E402
from oracle import physics_binding as pb  # noqa: E402

if __name__ == "__main__":
    assert pb.build_ref(), "the reference tree is needed to regenerate the vectors"
    d = cases.inputs()
    ref = pb.Side("ref")
    out = cases.run(ref, d)
    # the BSDF and emitter plugins on their own (the oracle folds them into diffuseReconnection; the test restates them)
    wi, wo = d["wi"], d["wo"]
    ev, pd = ref.diffuse_bsdf(d["albedo1"], d["normal"], wi, wo)
    out["bsdf_eval_bits"], out["bsdf_pdf_bits"] = cases.bits(ev), cases.bits(pd)
    ev, pd = ref.area_emitter(d["normal"], wo)
    out["emit_eval_bits"], out["emit_pdf_bits"] = cases.bits(ev), cases.bits(pd)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "physics_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
