#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo): where the code size comes from.

    python tools/sass_lines.py build/obj/gather_beams.o k_beam_shadeILb0 [top]
"""
import collections
import re
import subprocess
import sys

obj, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
import glob
import os
import tempfile
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cnt, by_func = collections.Counter(), collections.Counter()
cur, inside, n = None, False, 0
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        inside = kern in m.group(1)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line) and cur:
        cnt[cur] += 1
        n += 1
print(f"{n} instructions ({n * 16 / 1024:.1f} KB)")
files = collections.Counter()
for (f, l), c in cnt.items():
    files[f] += c
print("by file:", dict(files))
for (f, l), c in cnt.most_common(top):
    print(f"{c:6d} {100 * c / n:5.1f}%  {f}:{l}")
