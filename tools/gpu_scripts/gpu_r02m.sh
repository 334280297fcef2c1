#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for p in 0 1; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-search --pipeline $p > $O/r02m_bench_p$p.json 2> $O/r02m_bench_p$p.err; python -c "
import json; d=json.load(open('$O/r02m_bench_p$p.json')); print('pipeline $p', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"; done
