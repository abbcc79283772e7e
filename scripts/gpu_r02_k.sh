#!/bin/bash
set -x
mkdir -p gpurun_out
for v in main c24; do
  if [ $v = c24 ]; then export MDOPT_B200_LIB=$PWD/build/libmdopt_c24.so; else unset MDOPT_B200_LIB; fi
  timeout 300 python scripts/mono_profile.py --batch 100000 --launches 3 > gpurun_out/r02k_profile_$v.log 2>&1
  cat gpurun_out/r02k_profile_$v.log
done
unset MDOPT_B200_LIB
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02k_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_gputests.log
tail -6 gpurun_out/r02k_gputests.log
