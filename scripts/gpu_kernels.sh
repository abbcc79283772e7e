timeout 600 python scripts/bench_kernels.py > gpurun_out/r01_kernels.jsonl 2> gpurun_out/r01_kernels.err
cat gpurun_out/r01_kernels.jsonl | cut -c1-500; tail -3 gpurun_out/r01_kernels.err
