#!/bin/bash
# round 2, first GPU pass: the GPU test tier, a bench line, the launch list and one full ncu capture of the mono kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_gputests.log
tail -15 gpurun_out/r02a_gputests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
timeout 300 python scripts/mono_profile.py --batch 16384 --launches 2 > gpurun_out/r02a_profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02a_launches.csv \
    python scripts/mono_profile.py --batch 16384 --launches 2 > gpurun_out/r02a_ncu_launches.log 2>&1
cat gpurun_out/r02a_profile_plain.log
timeout 300 python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/r02a_profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/r02a_mono \
    python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/r02a_ncu_full.log 2>&1
ls -la gpurun_out | tail
