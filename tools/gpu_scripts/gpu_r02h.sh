#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02h_bench_n8.json 2> $O/r02h_bench_n8.err; echo "rc=$?"
tail -2 $O/r02h_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 10 --warmup 3 --no-search > $O/r02h_bench_n4.json 2> $O/r02h_bench_n4.err; echo "rc=$?"
timeout 300 python -m pytest tests/test_gpu_multi_device.py -q -m gpu --timeout 200 > $O/r02h_pytest.log 2>&1; echo "rc=$?" >> $O/r02h_pytest.log; tail -5 $O/r02h_pytest.log
timeout 200 python tools/multi_device_timing.py > $O/r02h_md_timing.json 2> $O/r02h_md_timing.err; cat $O/r02h_md_timing.json; tail -2 $O/r02h_md_timing.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/sharded_check.py > $O/r02h_sharded_check.log 2>&1; tail -1 $O/r02h_sharded_check.log
