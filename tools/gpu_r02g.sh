#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_sharded.py tests/test_gpu_native.py -q -m gpu --timeout 600 -k "multi_device or peer_all_reduce or dropin_uses or mass_of_zeros or nccl" > $O/r02g_pytest.log 2>&1; echo "rc=$?" >> $O/r02g_pytest.log
tail -30 $O/r02g_pytest.log
timeout 300 python tools/multi_device_timing.py > $O/r02g_md_timing.json 2> $O/r02g_md_timing.err; cat $O/r02g_md_timing.json; tail -3 $O/r02g_md_timing.err
