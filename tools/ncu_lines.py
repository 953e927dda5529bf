#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep (source page, cuda,sass view).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_bre_shade [top]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
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
        try:  # index from the right: a source line may contain quotes that confuse the CSV split
            ie = int(r[hdr.index("Instructions Executed") - len(hdr)])
            te = int(r[hdr.index("Thread Instructions Executed") - len(hdr)])
            sm = int(r[hdr.index("# Samples") - len(hdr)])
        except ValueError:
            continue
        out.append((ie, te, sm, path.split("/")[-1], r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out)
tsm = sum(o[2] for o in out)
print(f"total warp instructions {tot}, samples {tsm}")
out.sort(reverse=True)
for ie, te, sm, f, ln, src in out[:top]:
    print(f"{ie:>11} {ie / tot * 100:5.1f}%  thr/inst {te / max(ie, 1):5.1f}  samples {sm / max(tsm, 1) * 100:5.1f}%  {f}:{ln}  {src}")
