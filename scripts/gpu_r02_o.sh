#!/bin/bash
set -x
mkdir -p gpurun_out
for v in main unroll; do
  if [ $v = unroll ]; then export MDOPT_B200_LIB=$PWD/build/libmdopt_unroll.so; else unset MDOPT_B200_LIB; fi
  timeout 300 python scripts/mono_profile.py --batch 100000 --launches 3 > gpurun_out/r02o_profile_$v.log 2>&1
  cat gpurun_out/r02o_profile_$v.log
done
