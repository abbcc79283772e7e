#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02r_gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_gputests.log
tail -5 gpurun_out/r02r_gputests.log
python scripts/gpu_depol_counters.py 5 64 1184 > gpurun_out/r02r_depol_counters.json 2>gpurun_out/r02r.err; cat gpurun_out/r02r_depol_counters.json
timeout 900 python bench.py > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err; echo "bench exit $?"
head -c 300 gpurun_out/r02r_bench_n1.json; echo
timeout 900 python bench.py --workload config3 --lattice 5 --chi 64 --batch 1184 --steps 1 --warmup 1 > gpurun_out/r02r_config3_d5_chi64.json 2>/dev/null; head -c 200 gpurun_out/r02r_config3_d5_chi64.json; echo
timeout 900 python bench.py --workload config3 --lattice 9 --chi 64 --batch 296 --steps 1 --warmup 1 > gpurun_out/r02r_config3_d9_chi64.json 2>/dev/null; head -c 200 gpurun_out/r02r_config3_d9_chi64.json; echo
timeout 600 python scripts/bench_kernels.py > gpurun_out/r02r_kernels.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r02r_kernels.jsonl
