#!/usr/bin/env python
"""Print duration, issue utilisation and the largest warp-stall reasons per kernel from an ncu raw-page CSV (stdin)."""
import csv, sys
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[idx["Kernel Name"]][:44]
    get = lambda k: r[idx[k]] if k in idx else "-"
    stalls = sorted(((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                     and r[i] not in ("", "n/a")), reverse=True)[:6]
    print("%-44s %8s us  issue %5s%%  occ %5s%%  regs %s  inst %s" % (
        name, get("gpu__time_duration.sum"), get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        get("sm__warps_active.avg.pct_of_peak_sustained_active"), get("launch__registers_per_thread"), get("smsp__inst_executed.sum")))
    print("      stalls per issue: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls))
