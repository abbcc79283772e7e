#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r02i_gputests.log 2>&1; tail -5 gpurun_out/r02i_gputests.log
timeout 600 python scripts/bench_kernels.py > gpurun_out/r02i_kernels.jsonl 2> gpurun_out/r02i_kernels.err; cat gpurun_out/r02i_kernels.jsonl | cut -c1-420; tail -3 gpurun_out/r02i_kernels.err
