#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
tools/microbench/issue_model > $O/r02b_issue_model.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -k "benchmarked or concurrent or years_to_ruin or search_selects or fuzz" > $O/r02b_pytest.log 2>&1; echo "rc=$?" >> $O/r02b_pytest.log
tail -5 $O/r02b_pytest.log
