#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and count per kernel."""
import csv, re, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"pn::\(anonymous namespace\)::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    tot[name] += v_us; cnt[name] += 1
s = sum(tot.values())
print(f"{'kernel':72s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{k[:72]:72s} {cnt[k]:5d} {v:10.1f} {v/cnt[k]:9.1f} {v/s:6.1%}")
print(f"{'TOTAL':72s} {sum(cnt.values()):5d} {s:10.1f}")
