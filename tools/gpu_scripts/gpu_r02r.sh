#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_native.py -x -q -k "quantiles or histograms or aggregates or large_batch or sweep_mode" 2>&1 | tail -15
timeout 300 python tools/time_aggregates.py 2>&1 | tail -6 | tee $O/r02r_time_aggregates.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-search > $O/r02r_bench.json 2> $O/r02r_bench.err; python -c "
import json; d=json.load(open('$O/r02r_bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > $O/r02r_ncu.log 2>&1; echo "ncu rc=$?"
