#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
L=monte_carlo_retirement_b200/_lib
echo "== variants" > $O/r02k_variants.log
for v in "" _mb5; do
  echo "variant '$v'" >> $O/r02k_variants.log
  MCR_LIB=$PWD/$L/libmcr_b200$v.so timeout 300 python tools/run_timeline.py --reps 7 >> $O/r02k_variants.log 2>&1
done
timeout 300 python tools/run_timeline.py --reps 5 --scenario SYNTH_C3_VOL >> $O/r02k_variants.log 2>&1
timeout 300 python tools/run_timeline.py --reps 5 --no-series --search 40 >> $O/r02k_variants.log 2>&1
cat $O/r02k_variants.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/r02k_pytest.log 2>&1; echo "rc=$?" >> $O/r02k_pytest.log
tail -6 $O/r02k_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_timeline -s 2 -c 1 -o $O/prof_timeline_r02k -f python tools/run_timeline.py --reps 4 > $O/r02k_ncu_tl.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-search > $O/r02k_bench.json 2> $O/r02k_bench.err; python -c "
import json; d=json.load(open('$O/r02k_bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
