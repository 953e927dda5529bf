#!/usr/bin/env python
"""DRAM and L2 bytes per launch of the kernels in an `ncu --set full` capture -> an entry of profiles/traffic.json, plus
the raw-page CSV next to it (so that roofline.traffic can be re-derived from a committed artefact).

    python tools/ncu_traffic.py gpurun_out/r2_full_bre.ncu-rep cfg5 profiles/r2_full_bre_raw.csv.gz [--peak profiles/r2_l2_peak.json]
"""
import csv
import gzip
import json
import os
import subprocess
import sys

rep, workload, csv_out = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
with gzip.open(csv_out, "wt") as f:
    f.write(txt)
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]


def val(d, u, key):
    v = float(d[key].replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u[key], 1.0)


entry = {"kernels": {}, "source": f"{os.path.relpath(csv_out, ROOT)} (ncu --set full --clock-control none, raw page of {os.path.basename(rep)})"}
tot_dram = tot_l2 = 0.0
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    dram = val(d, u, "dram__bytes_read.sum") + val(d, u, "dram__bytes_write.sum")
    l2 = val(d, u, "lts__t_bytes.sum") if "lts__t_bytes.sum" in d else None
    if l2 is None and "lts__t_sectors.sum" in d:
        l2 = float(d["lts__t_sectors.sum"].replace(",", "")) * 32.0
    ms = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u["gpu__time_duration.sum"]]
    entry["kernels"][name] = {"dram_bytes": dram, "l2_bytes": l2, "ms_under_ncu": ms}
    tot_dram += dram
    tot_l2 += l2 or 0.0
entry["bytes_per_launch"] = tot_dram
entry["l2_bytes_per_launch"] = tot_l2
entry["kernel"] = " + ".join(entry["kernels"])
path = os.path.join(ROOT, "profiles", "traffic.json")
try:
    allj = json.load(open(path))
except Exception:
    allj = {}
allj["comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum (and lts__t_bytes.sum) per launch from the `ncu --set full` capture named in "
                   "`source`; bench.py copies the entry that matches its workload into roofline.traffic / roofline.l2 at N=1")
allj[workload] = entry
json.dump(allj, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))
