#!/usr/bin/env python
"""Join `ncu --page source --csv` (SASS view) with nvdisasm line info: executed warp instructions and
stall samples per CUDA source line.

usage: ncu_by_line.py src.csv file.cubin kernel_substring [top]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# offset -> (file, line) for the wanted kernel
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
off2line, infn, cur = {}, False, None
for l in dis:
    if l.startswith("//---") and ".text." in l:
        infn = kname in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
seen, recs = set(), []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["# Samples"]].isdigit():
        continue
    a = int(r[ix["Address"]], 16)
    if a in seen:
        continue
    seen.add(a)
    recs.append((a, int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
base = min(a for a, _, _ in recs)
per = defaultdict(lambda: [0, 0])
for a, ex, smp in recs:
    k = off2line.get(a - base, ("?", 0))
    per[k][0] += ex
    per[k][1] += smp
tot_ex = sum(v[0] for v in per.values())
tot_s = sum(v[1] for v in per.values())
print(f"total executed {tot_ex}  samples {tot_s}")
srcs = {}
for (f, ln), (ex, smp) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            path = [p for p in ("/root/repo/gpu-wah_b200/csrc/" + f,) ][0]
            srcs[f] = open(path).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:80] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100*ex/tot_ex:5.1f}% ex {100*smp/max(tot_s,1):5.1f}% st  {f}:{ln:4d}  {text}")
