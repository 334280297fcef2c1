#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r02d_env.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_search_vs_reference_draws.py tests/test_gpu_native.py -q -m gpu --timeout 600 -k "nccl or search_selects or fast_native or pooled" > $O/r02d_pytest.log 2>&1; echo "rc=$?" >> $O/r02d_pytest.log
tail -15 $O/r02d_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02d_bench_n2.json 2> $O/r02d_bench_n2.err
tail -3 $O/r02d_bench_n2.err
python -c "
import json
d=json.load(open('$O/r02d_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])"
