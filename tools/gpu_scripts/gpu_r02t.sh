#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
L=$PWD/monte_carlo_retirement_b200/_lib
timeout 600 python -m pytest tests/test_gpu_native.py -x -q -k "quantiles or histograms or aggregates or large_batch or sweep_mode" 2>&1 | tail -15
for v in "" _u8; do echo "== variant '$v'"; MCR_LIB=$L/libmcr_b200$v.so timeout 300 python tools/time_aggregates.py 2>&1 | tail -6; done | tee $O/r02t_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > $O/r02t_ncu.log 2>&1; echo "ncu rc=$?"
