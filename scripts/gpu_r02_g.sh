#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mono.py tests/test_gpu_generic.py tests/test_driver_and_backend.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/r02g_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_gputests.log
tail -8 gpurun_out/r02g_gputests.log
timeout 600 python scripts/gpu_config4_hgp.py 888 512 0.01 > gpurun_out/r02g_config4.json 2> gpurun_out/r02g_config4.err; echo "config4 exit $?"; cat gpurun_out/r02g_config4.json; tail -3 gpurun_out/r02g_config4.err
timeout 600 python scripts/sanitize_small.py > gpurun_out/r02g_sanitize_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02g_sanitize_memcheck.log 2>&1; echo "memcheck exit $?"
tail -6 gpurun_out/r02g_sanitize_memcheck.log
