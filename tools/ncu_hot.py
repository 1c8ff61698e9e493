"""Prints the hottest SASS instructions (warp-stall samples) of an ncu report's source page.
    python tools/ncu_hot.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {c: i for i, c in enumerate(hdr)}
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[h + 1:]:
    try:
        data.append((int(r[ci["# Samples"]]), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(rows[0][1][:100], "| total samples", tot, "| instructions", len(data))
for k, (n, r) in enumerate(data):
    if n >= tot * minpct / 100:
        st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{k:5d} {100 * n / tot:5.1f}%  exec={r[ci['Instructions Executed']]:>8}  {r[ci['Source']].strip()[:70]:70s} {st}")
