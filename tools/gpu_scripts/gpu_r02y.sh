#!/bin/bash
# 8 GPUs, final: bench line (with the search and c5 blocks), single-process MultiDevice timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02y_bench_n8.json 2> $O/r02y_bench_n8.err
python -c "
import json
d=json.load(open('$O/r02y_bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['ms_per_step'], d['config'].get('select_fallbacks'), d.get('c5'), {k:v.get('wall_s') for k,v in d.get('search',{}).items() if isinstance(v,dict)})"
tail -3 $O/r02y_bench_n8.err
timeout 200 python tools/multi_device_timing.py > $O/r02y_md_timing.json 2> $O/r02y_md_timing.err; cat $O/r02y_md_timing.json; tail -2 $O/r02y_md_timing.err
