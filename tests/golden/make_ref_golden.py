#!/usr/bin/env python
"""Generates tests/golden/ref_pins.npz: outputs of the REFERENCE'S OWN code (oracle/_ref/libgvpm_ref.so, built
by `make -C oracle ref` from /root/reference in the build container) on the seeded inputs of
tests/pin_cases.inputs("small").  The reference tree does not exist on the GPU box, so these vectors are what
pins the oracle restatement there; tests/test_oracle_ref_pin.py compares the oracle against them and, when the
library is present, against the live reference on larger inputs.

    python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import pin_cases  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

if __name__ == "__main__":
    if not rb.build_ref():
        raise SystemExit("the reference tree is not available here: cannot regenerate the golden vectors")
    ref = rb.Side("ref")
    inp = pin_cases.inputs("small")
    out = pin_cases.run(ref, inp)
    # structures only the reference has: which beams / planes its BVH traversals offer to the functor
    off, idx, t1, t2 = ref.subbeam_visits(inp["beam_o"], inp["beam_e"], float(inp["radius"]), inp["ray_o"], inp["ray_d"],
                                          inp["ray_mint"], inp["ray_maxt"])
    out.update(sub_off=off, sub_idx=idx, sub_t1_bits=t1.view(np.uint32), sub_t2_bits=t2.view(np.uint32))
    off, idx = ref.plane_visits(inp["pl_ori"], inp["pl_w0"], inp["pl_len0"], inp["pl_w1"], inp["pl_len1"], inp["ray_o"],
                                inp["ray_d"], inp["ray_mint"], inp["ray_maxt"])
    out.update(plv_off=off, plv_idx=idx)
    path = os.path.join(HERE, "ref_pins.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")
