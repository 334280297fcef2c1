#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > $O/r02e_pytest.log 2>&1; echo "rc=$?" >> $O/r02e_pytest.log
tail -8 $O/r02e_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02e_bench.json 2> $O/r02e_bench.err; echo "bench rc=$?"
tail -3 $O/r02e_bench.err
timeout 300 python tools/time_aggregates.py > $O/r02e_agg_time.log 2>&1
