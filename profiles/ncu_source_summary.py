#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall samples by reason and the hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; ncu_source_summary.py src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = Counter()
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[ix['# Samples']] == '# Samples' or not r[ix['# Samples']].isdigit():
        continue
    s = int(r[ix["# Samples"]] or 0)
    for n in stalls:
        tot[n] += int(r[ix[n]] or 0)
    lines.append((s, r[ix["Address"]], r[ix["Source"]], {n: int(r[ix[n]] or 0) for n in stalls if int(r[ix[n]] or 0)},
                  int(r[ix["Instructions Executed"]] or 0)))
total = sum(l[0] for l in lines)
print("kernel:", rows[0][1][:100])
print("total samples", total, " instructions executed (warp)", sum(l[4] for l in lines))
for n, v in tot.most_common(10):
    print(f"  {n:28s} {v:8d} {100.0 * v / max(total, 1):5.1f}%")
print("hottest SASS:")
for s, a, src, st, ie in sorted(lines, reverse=True)[:top]:
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {s:6d} {100.0 * s / max(total, 1):5.1f}%  {src[:70]:70s} {main}")
