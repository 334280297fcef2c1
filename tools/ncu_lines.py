#!/usr/bin/env python
"""Dynamic warp-instruction counts per CUDA source line from an ncu report (needs -lineinfo and
--import-source on), normalised per executed path-month.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <executed_path_months> [top_n] > profiles/<name>.md
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, pm = sys.argv[1], float(sys.argv[2]) / 32.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    per = collections.defaultdict(lambda: [0, 0, ""])
    cur = hdr = None
    for r in csv.reader(io.StringIO(txt)):
        if r and r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
        elif r and r[0] == "Line No":
            hdr, i_e, i_s = r, r.index("Instructions Executed"), r.index("# Samples")
        elif hdr and cur and len(r) > i_e and r[0].isdigit() and r[2] == "-":
            try:
                n = int(r[i_e] or 0)
            except ValueError:
                continue
            e = per[(cur, int(r[0]))]
            e[0] += n
            e[1] += int(r[i_s] or 0)
            e[2] = r[1].strip()[:110]
    total = sum(v[0] for v in per.values())
    files = collections.Counter()
    for (f, _), v in per.items():
        files[f] += v[0]
    print(f"# warp instructions per executed path-month by source line — `{rep}`\n")
    print(f"total {total / pm:.1f} per path-month; by file: " +
          ", ".join(f"`{f}` {n / pm:.1f}" for f, n in files.most_common() if n / pm >= 0.05) + "\n")
    print("| file:line | instr / path-month | stall samples | source |\n|---|---|---|---|")
    for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| `{f}:{ln}` | {v[0] / pm:.2f} | {v[1]} | `{v[2].replace('|', '¦')}` |")


if __name__ == "__main__":
    main()
