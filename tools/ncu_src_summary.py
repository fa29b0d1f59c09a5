#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: opcode mix, hottest instructions by stall samples, local-memory ops."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iX, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
iW = hdr.index('L1 Wavefronts Shared'); iWi = hdr.index('L1 Wavefronts Shared Ideal')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ops, samp = Counter(), Counter()
tot = totS = 0
recs = []
stalls = Counter()
for r in rows[2:]:
    s = r[iS].strip()
    try: n = int(r[iX]); ns = int(r[iN])
    except ValueError: continue
    toks = s.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    ops[op] += n; samp[op] += ns; tot += n; totS += ns
    recs.append((ns, n, r[0][-5:], s, int(r[iW] or 0), int(r[iWi] or 0)))
    for i in stall_cols: stalls[hdr[i]] += int(r[i] or 0)
print("warp instructions executed:", tot, " stall samples:", totS)
print("opcode: executed%  samples%")
for k, v in ops.most_common(22): print(f"  {k:12s} {100*v/tot:5.1f}%  {100*samp[k]/totS:5.1f}%")
print("stall reasons:", {k: f"{100*v/totS:.1f}%" for k, v in stalls.most_common(8)})
print("hottest instructions (samples, executed, addr, sass, shared wavefronts/ideal):")
for x in sorted(recs, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]: print("  ", x)
