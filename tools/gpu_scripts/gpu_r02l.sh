#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python tools/run_timeline.py --reps 7 > $O/r02l_timeline.log 2>&1; cat $O/r02l_timeline.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/r02l_pytest.log 2>&1; echo "rc=$?" >> $O/r02l_pytest.log
tail -6 $O/r02l_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_timeline -s 2 -c 1 -o $O/prof_timeline_r02l -f python tools/run_timeline.py --reps 4 > $O/r02l_ncu_tl.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02l_bench.json 2> $O/r02l_bench.err; python -c "
import json; d=json.load(open('$O/r02l_bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > $O/r02l_ncu_bench.log 2>&1
