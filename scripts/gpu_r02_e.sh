#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02e_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_gputests.log
tail -12 gpurun_out/r02e_gputests.log
timeout 900 python bench.py > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench exit $?"
head -c 600 gpurun_out/r02e_bench_n1.json; echo; tail -3 gpurun_out/r02e_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02e_bench_ref.json 2> gpurun_out/r02e_bench_ref.err; echo "ref exit $?"
cat gpurun_out/r02e_bench_ref.json | head -c 900; echo
timeout 900 python bench.py --workload config3 --lattice 5 --chi 64 --batch 1184 --steps 1 --warmup 1 > gpurun_out/r02e_config3_d5_chi64.json 2> gpurun_out/r02e_c3a.err; echo "c3a exit $?"
head -c 700 gpurun_out/r02e_config3_d5_chi64.json; echo; tail -3 gpurun_out/r02e_c3a.err
timeout 900 python bench.py --workload config3 --lattice 9 --chi 64 --batch 296 --steps 1 --warmup 1 > gpurun_out/r02e_config3_d9_chi64.json 2> gpurun_out/r02e_c3b.err; echo "c3b exit $?"
head -c 700 gpurun_out/r02e_config3_d9_chi64.json; echo; tail -3 gpurun_out/r02e_c3b.err
timeout 300 python bench.py --steps 2 --warmup 1 --batch 20000 --no-cpu-baseline > gpurun_out/r02e_small_plain.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 20000 --no-cpu-baseline > gpurun_out/r02e_ncu_launches.log 2>&1
ls gpurun_out | grep r02e
