#!/usr/bin/env python
"""Summarise an ncu report (ncu -i X.ncu-rep --page raw --csv) into a small markdown table.

usage: summarize.py report.ncu-rep [report2.ncu-rep ...] > summary.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"### {rep.split('/')[-1]}\n")
        names = [lbl for k, lbl in KEYS if k in ix]
        print("| kernel | " + " | ".join(names) + " |")
        print("|---|" + "---|" * len(names))
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            k = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
            vals = []
            for key, _ in KEYS:
                if key in ix:
                    v, u = r[ix[key]], units[ix[key]]
                    vals.append(f"{v} {u}".strip())
            print(f"| {k} | " + " | ".join(vals) + " |")
        print()


if __name__ == "__main__":
    main()
