#!/bin/bash
# 2 GPUs: sharded / multi-device tests, sharded check, N=2 bench, MultiDevice timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_sharded.py -q -m gpu --timeout 300 > $O/r02u_pytest.log 2>&1; echo "rc=$?" >> $O/r02u_pytest.log
tail -5 $O/r02u_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_check.py 2>&1 | tail -3
timeout 200 python tools/multi_device_timing.py > $O/r02u_md_timing.json 2> $O/r02u_md_timing.err; cat $O/r02u_md_timing.json; tail -3 $O/r02u_md_timing.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-search > $O/r02u_bench_n2.json 2> $O/r02u_bench_n2.err
python -c "
import json
d=json.load(open('$O/r02u_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['ms_per_step'], d['config'].get('select_fallbacks'))"
