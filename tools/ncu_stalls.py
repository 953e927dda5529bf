#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an .ncu-rep (source page): where a stall reason concentrates.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep k_bre_shade stall_long_sb [top]
"""
import csv
import subprocess
import sys

rep, kern, col = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out, path, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        path = r[1]
    elif len(r) > 4 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 10 and r[0].isdigit():
        try:
            v = int(r[hdr.index(col) - len(hdr)] or 0)
            sm = int(r[hdr.index("# Samples") - len(hdr)] or 0)
        except ValueError:
            continue
        out.append((v, sm, path.split("/")[-1], r[0], r[1].strip()[:110]))
tot = sum(o[0] for o in out)
tsm = sum(o[1] for o in out)
print(f"{col}: {tot} samples of {tsm} total ({tot / max(tsm, 1) * 100:.1f}%)")
out.sort(reverse=True)
for v, sm, f, ln, src in out[:top]:
    print(f"{v:>8} {v / max(tot, 1) * 100:5.1f}%  {f}:{ln}  {src}")
