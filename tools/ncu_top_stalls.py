#!/usr/bin/env python
"""Top stall locations of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: ncu -i rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {k: i for i, k in enumerate(h)}
key = ci["# Samples"]


def _num(x):
    try:
        float(x or 0)
        return True
    except ValueError:
        return False


body = [r for r in rows[hi + 1:] if len(r) >= len(h) - 1 and _num(r[key])]   # a second table (another view) may follow
tot = sum(float(r[key] or 0) for r in body)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(float(r[ci[k]] or 0) for r in body) for k in stalls}
print("total samples", tot)
print("by reason:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
order = sorted(range(len(body)), key=lambda i: -float(body[i][key] or 0))
for i in order[:n]:
    r = body[i]
    top = sorted(((float(r[ci[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{float(r[key]) / tot * 100:5.1f}%  #{i:4d} {r[ci['Source']].strip()[:90]:90s} {top[0][1]}={int(top[0][0])} {top[1][1]}={int(top[1][0])}")
