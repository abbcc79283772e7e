#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_gputests.log 2>&1; tail -4 gpurun_out/r02u_gputests.log
timeout 600 python bench.py --workload config3 --lattice 5 --chi 64 --batch 1184 --steps 2 --warmup 1 > gpurun_out/r02u_config3_d5_chi64.json 2>gpurun_out/r02u.err; cut -c1-200 gpurun_out/r02u_config3_d5_chi64.json
timeout 900 python bench.py --workload config3 --lattice 9 --chi 64 --batch 1184 --steps 1 --warmup 1 > gpurun_out/r02u_config3_d9_chi64.json 2>>gpurun_out/r02u.err; cut -c1-200 gpurun_out/r02u_config3_d9_chi64.json
timeout 900 python bench.py --workload config3 --lattice 7 --chi 128 --batch 148 --steps 1 --warmup 1 > gpurun_out/r02u_config3_d7_chi128.json 2>>gpurun_out/r02u.err; cut -c1-200 gpurun_out/r02u_config3_d7_chi128.json
tail -3 gpurun_out/r02u.err
