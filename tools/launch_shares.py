#!/usr/bin/env python
"""Compact an `ncu --metrics gpu__time_duration.sum --csv` launch list: one row per launch (id, kernel, grid, block,
ns) and the time share per kernel name. usage: launch_shares.py launches.csv OUT.csv OUT_shares.json"""
import collections
import csv
import json
import re
import sys

src, out_csv, out_json = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(smpc::)?(smpc_\w+)(<[^>]*>)?", name)
    if m:
        return m.group(2) + (m.group(3) or "")
    return re.sub(r"\(.*", "", name)[:90]


agg = collections.OrderedDict()
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_ns"])
    for r in rows:
        k = short(r[4])
        ns = float(r[14].replace(",", ""))
        w.writerow([r[0], k, r[8], r[7], int(ns)])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
total = sum(a[1] for a in agg.values())
shares = {k: {"launches": a[0], "total_ms": round(a[1] / 1e6, 3), "share": round(a[1] / total, 5)}
          for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
json.dump({"command": "ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 "
                      "--no-cpu-baseline (all legs; cold-cache, serialised launches: shares, not absolutes)",
           "launches": len(rows), "total_ms": round(total / 1e6, 2), "by_kernel": shares}, open(out_json, "w"), indent=1)
for k, v in list(shares.items())[:14]:
    print(f"{v['share']:8.4f} {v['total_ms']:12.2f} ms {v['launches']:6d}  {k}")
