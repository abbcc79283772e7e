#!/bin/bash
set -x
mkdir -p gpurun_out
for v in 0 28 24 16; do
  if [ $v = 0 ]; then unset MDOPT_MONO_LOCKSTEP; else export MDOPT_MONO_LOCKSTEP=$v; fi
  timeout 300 python scripts/mono_profile.py --batch 100000 --launches 3 > gpurun_out/r02m_profile_ls$v.log 2>&1
  cat gpurun_out/r02m_profile_ls$v.log
done
export MDOPT_MONO_LOCKSTEP=28
timeout 600 python -m pytest tests/test_gpu_mono.py -m gpu -q -x -k "headline or bench_batch or non_truncating" > gpurun_out/r02m_gputests_ls.log 2>&1; tail -4 gpurun_out/r02m_gputests_ls.log
