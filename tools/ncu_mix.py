#!/usr/bin/env python
"""Dynamic SASS instruction mix per executed path-month from an ncu report (source page).

    python tools/ncu_mix.py gpurun_out/prof.ncu-rep <executed_path_months> [--hot N]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, pm = sys.argv[1], float(sys.argv[2])
    hot = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if "Source" in r)
    i_s, i_e, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    mix, smp = collections.Counter(), collections.Counter()
    recs = []
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= i_e:
            continue
        src = r[i_s].strip()
        n = int(r[i_e] or 0)
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
        op = m.group(2) if m else src
        mix[op] += n
        smp[op] += int(r[i_smp] or 0)
        recs.append((n, int(r[i_smp] or 0), src))
    wpm = pm / 32.0
    tot = sum(mix.values())
    print(f"warp instructions {tot:.4e}; per executed path-month {tot / wpm:.1f}")
    fp64 = sum(n for op, n in mix.items() if op in ("DFMA", "DADD", "DMUL", "DSETP", "F2F", "DMNMX"))
    print(f"FP64-pipe instructions per path-month {fp64 / wpm:.1f}")
    for op, n in mix.most_common(28):
        print(f"  {op:10s} {n / wpm:7.1f}  samples {smp[op]}")
    if hot:
        mx = max(n for n, _, _ in recs)
        for i, (n, s, src) in enumerate(recs):
            if n > 0.5 * mx:
                print(f"{i:5d} {n:10d} {s:6d}  {src}")


if __name__ == "__main__":
    main()
