#!/usr/bin/env python
"""Measured L2 bandwidth of this B200 (BASELINE.json's metric names "% of HBM/L2 peak"; BASELINE.md left the L2 peak to be
measured): a device-to-device copy whose source + destination (2 x 24 MB) stay resident in the 126 MB L2, repeated back
to back, timed with CUDA events; read + write bytes per second, best of 5 batches.  Also the same copy over 2 x 2 GB
(HBM) for comparison with MEASURED_PEAKS.json.  Writes one JSON line (-> profiles/r2_l2_peak.json)."""
import json
import sys

import torch

dev = torch.device("cuda", 0)
out = {"gpu": torch.cuda.get_device_name(0)}
for name, nbytes, reps in (("l2", 24 << 20, 400), ("l2_48mb", 48 << 20, 200), ("hbm", 2 << 30, 10)):
    a = torch.empty(nbytes // 4, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = max(best, 2.0 * nbytes * reps / (ms * 1e-3) / 1e9)
    out[name + "_gbs"] = best
    out[name + "_bytes_each"] = nbytes
    del a, b
out["how"] = "torch b.copy_(a), read+write bytes, back to back, CUDA events, best of 5 batches; l2: 2 x 24 MB (L2-resident), hbm: 2 x 2 GB"
print(json.dumps(out))
