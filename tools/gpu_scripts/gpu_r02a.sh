#!/bin/bash
# round-2 GPU call A: full GPU suite, timeline variants, bench, sanitizer runs on the select kernels, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02a_env.txt; nproc >> $O/r02a_env.txt
L=monte_carlo_retirement_b200/_lib
echo "== variants" > $O/r02a_variants.log
for v in "" _mb8 _mb7 _mb5 _mb4 _b256mb3 _b64mb12; do
  echo "variant '$v'" >> $O/r02a_variants.log
  MCR_LIB=$PWD/$L/libmcr_b200$v.so timeout 300 python tools/run_timeline.py --reps 7 >> $O/r02a_variants.log 2>&1
done
timeout 300 python tools/run_timeline.py --reps 5 --scenario SYNTH_C3_VOL >> $O/r02a_variants.log 2>&1
timeout 300 python tools/run_timeline.py --reps 5 --no-series --search 40 >> $O/r02a_variants.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > $O/r02a_pytest.log 2>&1; echo "rc=$?" >> $O/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02a_bench.json 2> $O/r02a_bench.err
K="quantiles_match_pandas or pooled_tail or distributed_select or histograms"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_native.py tests/test_gpu_sharded.py -x -q -m gpu -k "$K" > $O/r02a_memcheck.log 2>&1; echo "rc=$?" >> $O/r02a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_native.py tests/test_gpu_sharded.py -x -q -m gpu -k "$K" > $O/r02a_racecheck.log 2>&1; echo "rc=$?" >> $O/r02a_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_timeline -s 2 -c 1 -o $O/prof_timeline_r02a -f python tools/run_timeline.py --reps 4 > $O/r02a_ncu_tl.log 2>&1
ls -la $O | tail -20
