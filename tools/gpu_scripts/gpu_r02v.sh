#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for i in 1 2 3 4 5; do for ms in 10 25 100; do
MCR_BENCH_SAMPLE_MS=$ms timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-search 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sample_ms=$ms', round(d['ms_per_step'],3), d['clocks']['samples'])"; done; done | tee $O/r02v_sampler.log
