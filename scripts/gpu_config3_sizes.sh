#!/bin/bash
set -x
mkdir -p gpurun_out
# config 3 at every measured size (after scripts/gpu_round.sh): d=9/chi=64, d=7/chi=128, d=9/chi=128, d=9/chi=256 (as stated)
for cfg in "9 64 1184 1" "7 128 148 1" "9 128 148 1" "9 256 148 0"; do
  set -- $cfg
  timeout 900 python bench.py --workload config3 --lattice $1 --chi $2 --batch $3 --steps 1 --warmup $4 > gpurun_out/round_config3_d$1_chi$2.json 2>>gpurun_out/round_bench.err
  cut -c1-170 gpurun_out/round_config3_d$1_chi$2.json
done
