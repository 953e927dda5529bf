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
# read-only sweep with the library's own 128-bit load kernel (gvpm_measure_read_bandwidth): 16 / 32 / 64 MB stay in L2
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from gvpm_b200.api import Context  # noqa: E402
ctx = Context(0)
for mb, reps in ((16, 400), (32, 200), (64, 100), (96, 60), (4096, 4)):
    out[f"read_{mb}mb_gbs"] = max(ctx.measure_read_bandwidth(mb << 20, reps) for _ in range(3))
ctx.close()
out["l2_read_peak_gbs"] = max(out["read_16mb_gbs"], out["read_32mb_gbs"], out["read_64mb_gbs"])
out["how"] = "torch b.copy_(a), read+write bytes, back to back, CUDA events, best of 5 batches; l2: 2 x 24 MB (L2-resident), hbm: 2 x 2 GB; read_*: k_l2_read (gvpm_measure_read_bandwidth), best of 3"
print(json.dumps(out))
