#!/bin/bash
set -x
mkdir -p gpurun_out
for v in main nopf; do
  if [ $v = nopf ]; then export MDOPT_B200_LIB=$PWD/build/libmdopt_nopf.so; else unset MDOPT_B200_LIB; fi
  timeout 300 python scripts/mono_profile.py --batch 65536 --launches 3 > gpurun_out/r02d_profile_$v.log 2>&1
  cat gpurun_out/r02d_profile_$v.log
done
unset MDOPT_B200_LIB
timeout 600 python -m pytest tests/test_gpu_mono.py tests/test_gpu_decode.py -q -x > gpurun_out/r02d_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_gputests.log
tail -5 gpurun_out/r02d_gputests.log
