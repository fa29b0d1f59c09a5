#!/usr/bin/env python
"""Per-kernel totals of the LAST sparse solve in an `ncu --csv --metrics gpu__time_duration.sum` launch list of
`tools/prof_driver.py sparse` (forward leaves -> separators -> backward).  usage: ncu_last_solve.py <launches.csv> [-v]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, out = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            out.append((d["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", ""), d["Grid Size"],
                        float(d["Metric Value"].replace(",", "")) / 1e3))
idx = [i for i, o in enumerate(out) if o[0].startswith("mf_forward_leaf")]
start = idx[-1]
while start - 1 in idx:
    start -= 1
agg = {}
for o in out[start:]:
    agg[o[0]] = agg.get(o[0], 0) + o[2]
for k, v in agg.items():
    print(f"{k:28s} {v:8.1f} us")
print(f"{'solve total':28s} {sum(agg.values()):8.1f} us;  factorisation kernels: "
      f"{sum(o[2] for o in out if 'factor' in o[0] or 'netperm' in o[0]) / max(1, sum(1 for o in out if o[0] == 'mf_netperm_kernel')):.1f} us")
if "-v" in sys.argv:
    for o in out[start:]:
        print("  %-28s grid %-16s %8.1f us" % o)
