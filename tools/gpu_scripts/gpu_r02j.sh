#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "quantiles or pooled or distributed_select or histograms or run_aggregates or mass_of_zeros or logical_shards or payload or large_batch or series_sweep" > $O/r02j_pytest.log 2>&1; echo "rc=$?" >> $O/r02j_pytest.log
tail -6 $O/r02j_pytest.log
timeout 300 python tools/time_aggregates.py > $O/r02j_agg_time.log 2>&1; cat $O/r02j_agg_time.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-search > $O/r02j_bench.json 2> $O/r02j_bench.err; python -c "
import json; d=json.load(open('$O/r02j_bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > $O/r02j_ncu_bench.log 2>&1
