#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 --no-search > $O/r02z_bench_n4.json 2> $O/r02z_bench_n4.err
python -c "
import json
d=json.load(open('$O/r02z_bench_n4.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['ms_per_step'], d['config'].get('select_fallbacks'))"
