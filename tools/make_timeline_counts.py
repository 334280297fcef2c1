#!/usr/bin/env python
"""profiles/timeline_counts.json from one ncu capture of the timeline kernel of THIS build:

    python tools/make_timeline_counts.py gpurun_out/prof_timeline.ncu-rep <executed_path_months> [label]

The file is keyed by monte_carlo_retirement_b200.build.source_hash(); bench.py ignores it (and says
so) when the hash differs from the sources it runs."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FP64 = ("DFMA", "DADD", "DMUL", "DSETP")


def main():
    from monte_carlo_retirement_b200.build import source_hash

    rep, pm = sys.argv[1], float(sys.argv[2])
    label = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hdr = srows[1]
    ie = hdr.index("Instructions Executed")
    per = {}
    for r in srows[2:]:
        try:
            n = int(r[ie])
        except (ValueError, IndexError):
            continue
        t = r[1].strip()
        op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
        per[op] = per.get(op, 0) + n
    wm = pm / 32.0
    total = sum(per.values())

    def f(name):
        return float(d[name].replace(",", ""))

    out = {
        "source_hash": source_hash(),
        "source": f"{label} (ncu --set full --clock-control none --import-source on, python tools/run_timeline.py, C3 shape)",
        "kernel": d.get("Kernel Name"),
        "executed_path_months": pm,
        "instr_per_path_month": total / wm,
        "fp64_instr_per_path_month": sum(per.get(k, 0) for k in FP64) / wm,
        "by_opcode_per_path_month": {k: round(v / wm, 2) for k, v in sorted(per.items(), key=lambda kv: -kv[1]) if v / wm >= 0.05},
        "ncu": {
            "gpu_time_ms": f("gpu__time_duration.sum") * (1e-6 if f("gpu__time_duration.sum") > 1e4 else 1.0),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "xu_pipe_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": int(f("launch__registers_per_thread")),
            "achieved_warps_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "dram_bytes_read": f("dram__bytes_read.sum") * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(
                dict(zip(rows[0], rows[1])).get("dram__bytes_read.sum", "byte"), 1.0),
            "dram_bytes_write": f("dram__bytes_write.sum") * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(
                dict(zip(rows[0], rows[1])).get("dram__bytes_write.sum", "byte"), 1.0),
        },
    }
    path = os.path.join(ROOT, "profiles", "timeline_counts.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: out[k] for k in ("source_hash", "instr_per_path_month", "fp64_instr_per_path_month")}), "->", path)


if __name__ == "__main__":
    main()
