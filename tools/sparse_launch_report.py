#!/usr/bin/env python
"""profiles/r02_sparse_launches.txt from an ncu launch list of `tools/prof_driver.py sparse` (config 3): per-kernel totals
and the per-launch device time / DRAM bytes / occupancy of the LAST solve in the list.
usage: sparse_launch_report.py <launches_raw.csv> [title]"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value")}
L = {}
for r in rows[1:]:
    if len(r) != len(h):
        continue
    k = int(r[ix["ID"]])
    L.setdefault(k, {"name": r[ix["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "").split("(")[0].split("<")[0], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    try:
        L[k][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        pass
ids = sorted(L)
# the last solve = the trailing run of forward/backward launches; the factorisation = the mf_factor launches right before the first sweep
sweeps = [k for k in ids if L[k]["name"].startswith(("mf_forward", "mf_backward"))]
start = 0
for n in range(1, len(sweeps)):   # a forward launch right after a backward launch opens a new solve
    if L[sweeps[n]]["name"].startswith("mf_forward") and L[sweeps[n - 1]]["name"].startswith("mf_backward"):
        start = n
solve = sweeps[start:]
nsolve = len(solve)
first_sweep = next(k for k in ids if L[k]["name"].startswith("mf_forward"))
fact = [k for k in ids if k < first_sweep and L[k]["name"].startswith(("mf_factor", "mf_netperm"))]
fact = fact[-(len(fact) // max(1, sum(1 for k in fact if L[k]["name"].startswith("mf_netperm")))):] if fact else fact
print("# " + (sys.argv[2] if len(sys.argv) > 2 else "config 3 (MPC QP, N = 240 000, 256 right-hand sides)") + ": ONE factorisation and ONE solve")
print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active...,smsp__issue_active... --clock-control none -c 400 --csv python tools/prof_driver.py sparse   (serialised, cold-cache launches)")
def line(k):
    d = L[k]
    t = d.get("gpu__time_duration.sum", 0.0) / 1e3
    rd, wr = d.get("dram__bytes_read.sum", 0.0) / 1e6, d.get("dram__bytes_write.sum", 0.0) / 1e6
    s = f"  {d['name']:28s} grid {d['grid']:16s} block {d['block']:14s} {t:8.1f} us"
    if rd or wr:
        s += f"   DRAM read {rd:7.1f} MB  write {wr:7.1f} MB  ({(rd + wr) / max(t, 1e-9):5.2f} TB/s)"
    if "sm__warps_active.avg.pct_of_peak_sustained_active" in d:
        s += f"   warps {d['sm__warps_active.avg.pct_of_peak_sustained_active']:4.1f} %  issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):4.1f} %"
    return s, t, rd + wr
for title, ks in (("factorisation", fact), ("solve (forward sweep, then backward sweep)", solve)):
    tot = byt = 0.0
    out = []
    per = {}
    for k in ks:
        s, t, b = line(k)
        out.append(s); tot += t; byt += b
        per[L[k]["name"]] = per.get(L[k]["name"], 0.0) + t
    print(f"{title}: {tot:.1f} us over {len(ks)} launches" + (f", {byt / 1e3:.2f} GB of DRAM traffic ({byt / max(tot, 1e-9):.2f} TB/s)" if byt else ""))
    for n, t in sorted(per.items(), key=lambda kv: -kv[1]):
        print(f"    {n:28s} {t:8.1f} us")
    print("\n".join(out))
