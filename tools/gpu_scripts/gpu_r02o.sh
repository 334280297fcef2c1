#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 300 python tools/diag_step.py 2>&1 | tee $O/r02o_diag.log
