#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
L=$PWD/monte_carlo_retirement_b200/_lib
timeout 600 python -m pytest tests/test_gpu_native.py -x -q -k "quantiles or histograms or aggregates or large_batch or sweep_mode" 2>&1 | tail -3
for v in "" _u8 _u8c128 _u4c32; do echo "== variant '$v'"; MCR_LIB=$L/libmcr_b200$v.so timeout 300 python tools/time_aggregates.py 2>&1 | tail -6; done | tee $O/r02q_variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sel_(hist|collect)$' -c 8 -o $O/prof_select_r02q python tools/time_aggregates.py > $O/r02q_ncu.log 2>&1; echo "ncu rc=$?"
