#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02n_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_gputests.log
tail -5 gpurun_out/r02n_gputests.log
timeout 900 python bench.py > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench exit $?"
head -c 300 gpurun_out/r02n_bench_n1.json; echo; tail -3 gpurun_out/r02n_bench_n1.err
timeout 300 python bench.py --steps 2 --warmup 1 --batch 20000 --no-cpu-baseline > gpurun_out/r02n_small_plain.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 20000 --no-cpu-baseline > gpurun_out/r02n_ncu_launches.log 2>&1
timeout 300 python scripts/mono_profile.py --batch 20000 --launches 1 > gpurun_out/r02n_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/r02n_mono \
    python scripts/mono_profile.py --batch 20000 --launches 1 > gpurun_out/r02n_ncu_full.log 2>&1
cat gpurun_out/r02n_profile_plain.log
