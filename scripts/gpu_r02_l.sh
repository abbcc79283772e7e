#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err; echo "bench exit $?"
head -c 500 gpurun_out/r02l_bench_n1.json; echo; tail -3 gpurun_out/r02l_bench_n1.err
timeout 300 python bench.py --steps 2 --warmup 1 --batch 20000 --no-cpu-baseline > gpurun_out/r02l_small_plain.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02l_launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 20000 --no-cpu-baseline > gpurun_out/r02l_ncu_launches.log 2>&1
timeout 300 python scripts/mono_profile.py --batch 20000 --launches 1 > gpurun_out/r02l_profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/r02l_mono \
    python scripts/mono_profile.py --batch 20000 --launches 1 > gpurun_out/r02l_ncu_full.log 2>&1
cat gpurun_out/r02l_profile_plain.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
