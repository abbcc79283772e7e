#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/mono_profile.py --batch 65536 --launches 2 > gpurun_out/r02b_profile_plain.log 2>&1
cat gpurun_out/r02b_profile_plain.log
timeout 1500 python -m pytest tests/test_gpu_mono.py -x -q -s > gpurun_out/r02b_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_gputests.log
tail -25 gpurun_out/r02b_gputests.log
timeout 300 python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/r02b_profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mono_decode -s 1 -c 1 -o gpurun_out/r02b_mono \
    python scripts/mono_profile.py --batch 8192 --launches 1 > gpurun_out/r02b_ncu_full.log 2>&1
