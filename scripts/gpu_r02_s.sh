#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py -m gpu -q -x > gpurun_out/r02s_gputests.log 2>&1; tail -4 gpurun_out/r02s_gputests.log
python scripts/gpu_depol_counters.py 5 64 1184 > gpurun_out/r02s_depol_counters.json 2>gpurun_out/r02s.err; cat gpurun_out/r02s_depol_counters.json; tail -3 gpurun_out/r02s.err
