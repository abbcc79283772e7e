#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload config3 --lattice 9 --chi 128 --batch 148 --steps 1 --warmup 0 > gpurun_out/r02v_config3_d9_chi128.json 2>gpurun_out/r02v.err; cut -c1-200 gpurun_out/r02v_config3_d9_chi128.json
timeout 1500 python bench.py --workload config3 --lattice 9 --chi 256 --batch 148 --steps 1 --warmup 0 > gpurun_out/r02v_config3_d9_chi256.json 2>>gpurun_out/r02v.err; cut -c1-200 gpurun_out/r02v_config3_d9_chi256.json
tail -3 gpurun_out/r02v.err
