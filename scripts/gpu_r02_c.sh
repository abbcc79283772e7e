#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/mono_profile.py --batch 65536 --launches 2 > gpurun_out/r02c_profile_plain.log 2>&1
cat gpurun_out/r02c_profile_plain.log
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r02c_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_gputests.log
tail -25 gpurun_out/r02c_gputests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"
cat gpurun_out/r02c_bench.json | head -c 1500; tail -3 gpurun_out/r02c_bench.err
timeout 300 python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/r02c_profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/r02c_mono \
    python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/r02c_ncu_full.log 2>&1
