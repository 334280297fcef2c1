#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_native.py -x -q -k "quantiles or histograms or aggregates or large_batch or sweep_mode" 2>&1 | tail -15
timeout 300 python tools/time_aggregates.py 2>&1 | tail -6 | tee $O/r02s_time_aggregates.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > $O/r02s_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sel_(hist|collect|tail)$' -s 14 -c 7 -o $O/prof_select_r02s python tools/time_aggregates.py > $O/r02s_ncu2.log 2>&1; echo "ncu rc=$?"
