#!/bin/bash
# config 3 (d=9 depolarising) on the dense global-memory tier: chi 128, then chi 256 if the first is quick enough
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generic.py -m gpu -q > gpurun_out/r02j_gputests.log 2>&1; tail -3 gpurun_out/r02j_gputests.log
S=$(date +%s)
timeout 900 python bench.py --workload config3 --lattice 9 --chi 128 --batch 148 --steps 1 --warmup 0 > gpurun_out/r02j_config3_d9_chi128.json 2> gpurun_out/r02j_c3a.err; echo "chi128 exit $?"
E=$(date +%s); echo "chi128 took $((E-S)) s"
head -c 900 gpurun_out/r02j_config3_d9_chi128.json; echo; tail -2 gpurun_out/r02j_c3a.err
if [ $((E-S)) -lt 200 ]; then
timeout 1500 python bench.py --workload config3 --lattice 9 --chi 256 --batch 148 --steps 1 --warmup 0 > gpurun_out/r02j_config3_d9_chi256.json 2> gpurun_out/r02j_c3b.err; echo "chi256 exit $?"
head -c 900 gpurun_out/r02j_config3_d9_chi256.json; echo; tail -2 gpurun_out/r02j_c3b.err
fi
