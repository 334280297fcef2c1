#!/usr/bin/env python
"""Markdown table of the select kernels in one ncu --set full report (one column per captured launch).

    python tools/ncu_select_summary.py <report.ncu-rep> "<title>" > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].split("(")[0].split("::")[-1] for r in data]
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none, cold caches, serialised)\n")
    print("| metric | unit | " + " | ".join(f"{i}: `{n}`" for i, n in enumerate(names)) + " |")
    print("|---|---|" + "---|" * len(names))
    for k in KEYS:
        if k in col:
            print(f"| `{k}` | {units[col[k]]} | " + " | ".join(r[col[k]] for r in data) + " |")
    t, b = col["gpu__time_duration.sum"], col["dram__bytes_read.sum"]
    print("\n| launch | DRAM read GB/s |\n|---|---|")
    for i, r in enumerate(data):
        try:
            us = float(r[t].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[t], 1.0)
            mb = float(r[b].replace(",", "")) * {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}.get(units[b], 1.0)
            print(f"| {i}: `{names[i]}` | {mb / us * 1e3:.0f} |")
        except ValueError:
            pass


if __name__ == "__main__":
    main()
