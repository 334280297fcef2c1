#!/bin/bash
# final 1-GPU evidence of the round: bench line, launch list (the ncu --set full capture of the select kernels,
# profiles/r02x_select_summary.md, was taken by an earlier version of this script: + "ncu --set full
# --clock-control none --import-source on -k regex:k_sel_ -s 12 -c 6 -o … python tools/run_step.py 3")
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02x_bench_n1.json 2> $O/r02x_bench_n1.err; python -c "
import json; d=json.load(open('$O/r02x_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02x_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > $O/r02x_ncu.log 2>&1; echo "ncu launch list rc=$?"
