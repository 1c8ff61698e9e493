"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/summarize_launches.py launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ci = {c: i for i, c in enumerate(rows[h])}
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rows[h + 1:]:
    if len(r) <= ci["Metric Value"] or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])
    name = re.sub(r"^void ", "", name)[:90]
    agg[name][0] += 1
    agg[name][1] += us
    total += us
print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f}% |")
print(f"\ntotal {total / 1000:.2f} ms over {sum(v[0] for v in agg.values())} launches "
      "(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)")
