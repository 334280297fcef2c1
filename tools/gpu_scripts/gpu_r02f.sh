#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r02f_env.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02f_bench_n8.json 2> $O/r02f_bench_n8.err; echo "rc=$?"
tail -3 $O/r02f_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/sharded_check.py > $O/r02f_sharded_check.log 2>&1; tail -2 $O/r02f_sharded_check.log
