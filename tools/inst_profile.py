#!/usr/bin/env python3
"""Executed warp instructions per SOURCE line of one kernel: joins the SASS source page of an ncu report (instructions
executed per SASS instruction) with the line table of the SAME build (nvdisasm --print-line-info).
usage: inst_profile.py report.ncu-rep kernel_substring cubin_substring [lib.so]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kernel, want = sys.argv[1], sys.argv[2], sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else "connect4_b200/lib/libc4b200.so"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(int(r[iE]), int(r[iS])) for r in rows[2:] if len(r) > iE and r[iE].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if want in f][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        on = kernel in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l):
        lines.append(cur)
n = len(lines)
assert len(sass) % n == 0, (len(sass), n)
agg, smp = collections.Counter(), collections.Counter()
for (e, s), loc in zip(sass[:n], lines):
    agg[loc] += e
    smp[loc] += s
tot = sum(agg.values())
src = {}
print("total executed warp instructions %.3g" % tot)
for (f, ln), e in agg.most_common(40):
    if f not in src:
        p = [os.path.join(d, f) for d in ("connect4_b200/csrc", "include") if os.path.exists(os.path.join(d, f))]
        src[f] = open(p[0]).read().splitlines() if p else []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100.0 * e / tot, 100.0 * smp[(f, ln)] / max(1, sum(smp.values())), f, ln, text))
