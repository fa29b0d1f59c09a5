#!/usr/bin/env python
"""Sum an `ncu --csv --metrics ...` launch list per kernel: launches, average of every metric.
usage: ncu_launch_summary.py <launches.csv>"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        k = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        agg[k][d["Metric Name"] + " [" + d["Metric Unit"] + "]"] += v
        if d["Metric Name"] == "gpu__time_duration.sum":
            cnt[k] += 1
tot = sum(m.get("gpu__time_duration.sum [ns]", m.get("gpu__time_duration.sum [us]", 0.0)) for m in agg.values())
for k, m in sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum [ns]", kv[1].get("gpu__time_duration.sum [us]", 0.0))):
    t = m.get("gpu__time_duration.sum [ns]", m.get("gpu__time_duration.sum [us]", 0.0))
    print(f"{k:34s} launches {cnt[k]:4d}  share {100 * t / tot:5.1f}%  " +
          "  ".join(f"{a}={b / max(cnt[k], 1):.4g}" for a, b in sorted(m.items())))
