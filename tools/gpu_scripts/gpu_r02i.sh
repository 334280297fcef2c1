#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "scenario_sweep or mass_of_zeros or pooled or quantiles or reference_test_suite or benchmarked" > $O/r02i_pytest.log 2>&1; echo "rc=$?" >> $O/r02i_pytest.log
tail -6 $O/r02i_pytest.log
