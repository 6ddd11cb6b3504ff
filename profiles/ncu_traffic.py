#!/usr/bin/env python
"""DRAM traffic per launch out of `ncu --set full` captures, for bench.py's roofline.traffic.

    python profiles/ncu_traffic.py WORKLOAD DENSITY MODE report.ncu-rep [more.ncu-rep ...]

reads dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum of every wah_* kernel launch in the reports
and merges them into profiles/ncu_traffic.json as table[WORKLOAD][DENSITY][MODE][kernel].  The captures are made with
scripts/prof_kernels.py on the same generator, size and seed family bench.py uses for that workload."""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) == len(hdr):
            yield {h: (v, u) for h, v, u in zip(hdr, r, units)}


def num(cell):
    v, u = cell
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)


def main():
    workload, density, mode = sys.argv[1:4]
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    slot = table.setdefault(workload, {}).setdefault(f"{float(density):g}", {}).setdefault(mode, {})
    for rep in sys.argv[4:]:
        for r in rows_of(rep):
            name = r["Kernel Name"][0].split("(")[0].split("::")[-1].split("<")[0]
            if not name.startswith("wah_"):
                continue
            slot[name] = {"dram_read_bytes": num(r["dram__bytes_read.sum"]), "dram_write_bytes": num(r["dram__bytes_write.sum"]),
                          "duration_us_under_ncu": num(r["gpu__time_duration.sum"]), "report": os.path.basename(rep)}
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(slot, indent=1))


if __name__ == "__main__":
    main()
