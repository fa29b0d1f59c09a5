#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line.

usage: ncu_line_summary.py <sass.csv> <object.o> <kernel-substring> <source.cu> [top]
Maps SASS offsets to lines with `nvdisasm --print-line-info` on the cubin extracted from the object
(the profiled binary must be the one built from the current source)."""
import csv, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict
csvf, obj, kern, src = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
base = os.path.basename(src)
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        # keep the innermost location that is in our source file; inlined-at chains mention it too
        if os.path.basename(m.group(1)) == base:
            cur = int(m.group(2))
        else:
            m2 = re.search(r'inlined at "[^"]*%s", line (\d+)' % re.escape(base), ln)
            cur = int(m2.group(1)) if m2 else cur
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvf)))
hdr = rows[1]
iA, iX, iN = 0, hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
addrs = []
for r in rows[2:]:
    try:
        addrs.append(int(r[iA], 16))
    except ValueError:
        addrs.append(None)
a0 = min(a for a in addrs if a is not None)
samp, execd = Counter(), Counter()
st = defaultdict(Counter)
for r, a in zip(rows[2:], addrs):
    if a is None:
        continue
    try:
        n, ns = int(r[iX]), int(r[iN])
    except ValueError:
        continue
    l = line_of.get(a - a0)
    samp[l] += ns
    execd[l] += n
    for i in stall_cols:
        st[l][hdr[i][6:]] += int(r[i] or 0)
tot = sum(samp.values())
srcl = open(src).read().splitlines()
print(f"total samples {tot}, warp instructions {sum(execd.values())}")
for l, ns in samp.most_common(top):
    txt = srcl[l - 1].strip()[:90] if l else "?"
    top2 = ", ".join(f"{k}:{v}" for k, v in st[l].most_common(3))
    print(f"{100*ns/tot:5.1f}%  exec {execd[l]:>9}  L{l}: {txt}   [{top2}]")
