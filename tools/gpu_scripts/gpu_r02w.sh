#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1100 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/time_aggregates.py | tail -1
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-search 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['gpu_launches'])"; done
