#!/usr/bin/env python3
"""Per-source-line stall samples of one kernel: joins the SASS source page of an ncu report (stall samples per
instruction) with the line table of the SAME build of the library (nvdisasm --print-line-info).
usage: line_profile.py report.ncu-rep mangled_kernel_name [lib.so]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kernel = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "connect4_b200/lib/libc4b200.so"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
sass = [(int(r[iS]), r[iSrc].strip()) for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
want = sys.argv[4] if len(sys.argv) > 4 else "c4_search"
cub = [f for f in os.listdir(tmp) if want in f][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        on = (".text." + kernel + " ") in l + " " or l.strip().endswith(kernel + " --------------------------")
        on = kernel in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
n = len(lines)
copies = max(1, len(sass) // n)
assert len(sass) % n == 0, (len(sass), n)
agg = collections.Counter()
tot = 0
for k in range(copies):                      # the report lists the kernel once per profiled launch
    for (smp, txt), (loc, dtxt) in zip(sass[k * n:(k + 1) * n], lines):
        agg[loc] += smp
        tot += smp
src = {}
for (f, ln), s in agg.most_common(45):
    if f not in src:
        p = [os.path.join(d, f) for d in ("connect4_b200/csrc", "include") if os.path.exists(os.path.join(d, f))]
        src[f] = open(p[0]).read().splitlines() if p else []
    text = src[f][ln - 1].strip()[:110] if 0 < ln <= len(src[f]) else ""
    print("%5.1f%%  %s:%d  %s" % (100.0 * s / tot, f, ln, text))
